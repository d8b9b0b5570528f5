"""SASS evidence per kernel of libaccflow_b200.so: counts of the tcgen05 / TMA / TMEM mnemonics
(cuobjdump -sass).  python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "accflow_b200", "libaccflow_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTCBAR.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UBLKCP", "LDGSTS", "MUFU.EX2", "MUFU.RCP", "HMMA", "FFMA"]
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter(); sizes[cur] = 0
        continue
    if cur is None or "/*" not in line:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(2); sizes[cur] += 1
    for k in KEYS:
        if "2CTA" in k:
            if op.startswith(k.split(".")[0]) and ".2CTA" in op:
                counts[cur][k] += 1
        elif (op == k or op.startswith(k + ".")) and ".2CTA" not in op:
            counts[cur][k] += 1
print("# cuobjdump -sass accflow_b200/libaccflow_b200.so: instruction counts per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTMASTG = TMA tensor store,")
print("# SYNCS = mbarrier ops, LDGSTS = cp.async")
for k, c in counts.items():
    tags = " ".join(f"{n}={v}" for n, v in c.items() if v)
    print(f"{k[:70]:70s} sass={sizes[k]:6d}  {tags}")

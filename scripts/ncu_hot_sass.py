"""Top stalled SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr = b["rows"][0]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    ns = idx["# Samples"]
    tot = sum(int(r[ns] or 0) for r in data)
    print(f"== {b['name'][:90]}  samples={tot} instr={len(data)}")
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ns] or 0))[:topn]
    for i in sorted(order):
        r = data[i]; n = int(r[ns])
        st = {h[6:]: int(r[idx[h]] or 0) for h in stall}
        st = {k: v for k, v in st.items() if v > 0.15 * n}
        print(f"#{i:5d} {n:6d} {100*n/max(tot,1):5.1f}%  {r[idx['Source']].strip()[:60]:60s} {st}")

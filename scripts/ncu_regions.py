"""Where the warp-state samples of a kernel go: python scripts/ncu_regions.py <csv of `ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1`>"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[0] not in ('Address','Kernel Name')]
# first kernel only: stop when address resets
out=[]; last=None
for r in data:
    a=int(r[0],16)
    if last is not None and a<last: break
    out.append(r); last=a
data=out
ns=idx['# Samples']; stall=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[ns]) for r in data)
agg=collections.Counter()
for r in data:
    for h in stall: agg[h[6:]]+=int(r[idx[h]] or 0)
print(rows[0][1][:60],'instr',len(data),'samples',tot,'stalls',agg.most_common(8))
# waits on mbarriers: BRA following SYNCS within 3 instrs
w=0
for i,r in enumerate(data):
    src=r[idx['Source']]
    if 'SYNCS' in src or ('BRA' in src and i>0 and any('SYNCS' in data[j][idx['Source']] for j in range(max(0,i-4),i))):
        w+=int(r[ns])
print('samples in mbarrier wait loops', w, f"{100*w/tot:.1f}%")
top=sorted(range(len(data)),key=lambda i:-int(data[i][ns]))[:14]
for i in sorted(top):
    r=data[i]; n=int(r[ns]); st={h[6:]:int(r[idx[h]] or 0) for h in stall}; st={k:v for k,v in st.items() if v>0.2*n}
    print(f"#{i:5d} {n:5d} {100*n/tot:4.1f}% {r[idx['Source']].strip()[:56]:56s} {st}")

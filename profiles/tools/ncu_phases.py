"""Stall samples of an `ncu --set full --import-source on` report, aggregated between barriers. Usage: ncu_phases.py source.csv"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
h=rows[hi]
iS,iN,iE=h.index('Source'),h.index('# Samples'),h.index('Instructions Executed')
stall_cols=[i for i,c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
cum=0; last=0; marks=[]; data=[]; tot_inst=0; linst=0
agg={}
for k,r in enumerate(rows[hi+1:]):
    try: n=int(r[iN])
    except: continue
    cum+=n; data.append((n,k,r)); tot_inst+=int(r[iE] or 0)
    for i in stall_cols: agg[h[i]]=agg.get(h[i],0)+int(r[i] or 0)
    s=r[iS]
    if 'BAR.SYNC' in s or 'EXIT' in s:
        marks.append((k,s.strip()[:40],cum-last,tot_inst-linst)); last=cum; linst=tot_inst
for m in marks: print(m)
print('total samples',cum,'warp insts',tot_inst)
print(sorted(agg.items(),key=lambda t:-t[1])[:8])
for n,k,r in sorted(data,key=lambda t:-t[0])[:int(sys.argv[2]) if len(sys.argv)>2 else 20]:
    st=sorted([(int(r[i] or 0),h[i]) for i in stall_cols],reverse=True)[:2]
    print(f"{k:5d} {n:6d} {r[iE]:>9s} {r[iS].strip()[:70]:70s} {st}")

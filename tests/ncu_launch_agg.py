"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name and grid:

    python tests/ncu_launch_agg.py profiles/r2_launches_bench_steps1.csv

Runs on the CPU box."""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None
agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        k=d['Kernel Name'][:48]+' g'+d.get('Grid Size','')
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        u=d['Metric Unit']
        if u=='ns': v/=1000
        elif u=='ms': v*=1000
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print(f"{k:75s} n={n:4d} total={t:9.1f}us avg={t/n:7.1f}")

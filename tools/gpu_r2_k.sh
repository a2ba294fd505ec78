#!/bin/bash
# launch list of the synthetic workload's step (where does the non-ODE time go at B = 1,024?)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_synthetic.csv \
  python bench.py --workload synthetic_dr_constant --steps 3 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02_launches_syn.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_launches_synthetic.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[rows.index(hdr)+1:]:
    if len(r)<=vi: continue
    k=r[ki].split('<')[0].split('(')[0]
    agg.setdefault(k,[]).append(float(r[vi].replace(',',''))/1e3)
for k,v in agg.items(): print('%-50s n=%3d  median %.1f us  last %.1f' % (k[:50], len(v), sorted(v)[len(v)//2], v[-1]))
PY

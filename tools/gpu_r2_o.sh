#!/bin/bash
# round 2, pass O: encoder kernels on an instruction diet (run-based staging, quad conv weight gradient, vector hidden layer,
# vector weight-gradient + Adam): tests, bench, launch list
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 --no-extra-workloads 2>gpurun_out/r02o_bench.err | tail -1 > gpurun_out/r02o_bench_1gpu.json
tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print({k:d['e2e'][k] for k in ('value','ms_per_step','latency_ms')})
PY
for v in "VIHDS_PDL=0" "VIHDS_ONE_GRAPH=0" "VIHDS_PDL=0 VIHDS_ONE_GRAPH=0"; do
  env $v timeout 300 python bench.py --steps 100 --warmup 5 --no-extra-workloads --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02o_ab.json
  echo "$v: $(python -c "import json;d=json.load(open('gpurun_out/r02o_ab.json'));print(round(d['ms_per_step'],5), round(d['e2e']['ms_per_step'],5), round(d['e2e']['latency_ms'],5))")"
done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02o_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02o_launches.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02o_launches_bench.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
body=rows[rows.index(hdr)+1:]
seq=[(r[ki].split('<')[0].replace('void vh::','')[:34], r[mi], float(r[vi].replace(',',''))) for r in body if len(r)>vi]
idx=[i for i,(k,m,_) in enumerate(seq) if k.startswith('conditioner') and m.startswith('gpu__time')]
for k,m,v in seq[idx[-2]:idx[-1]]:
    print('%-36s %-28s %12.1f'%(k,m,v/1e3 if m.startswith('gpu__time') else v))
PY

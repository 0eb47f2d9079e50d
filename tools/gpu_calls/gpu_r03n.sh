#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r03n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03n_pytest_gpu.log
tail -2 gpurun_out/r03n_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r03n_bench.csv python bench.py --steps 2 --warmup 1 --seg 40 --no-cpu --no-extra > gpurun_out/r03n_ncu_list.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_r03n_bench.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i; break
col={n:i for i,n in enumerate(rows[h])}
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[h+1:]:
    if len(r)<len(col) or r[col['Metric Name']]!='gpu__time_duration.sum': continue
    v=float(r[col['Metric Value']].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}.get(r[col['Metric Unit']],1)
    k=r[col['Kernel Name']].split('(')[0][:50]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:4]: print('%-50s n=%4d total %.1f us (%.1f %%) avg %.2f us'%(k,v[0],v[1],100*v[1]/tot,v[1]/v[0]))
PY

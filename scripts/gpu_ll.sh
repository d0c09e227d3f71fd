#!/bin/bash
# launch list only (per-kernel serialised durations under ncu; never a bench number)
mkdir -p gpurun_out
R=${1:-tmp}
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"ani_|seg" -s 9 -c 9 --csv --log-file gpurun_out/ll_$R.csv python scripts/profile_app.py 50000 3 > gpurun_out/ll_$R.log 2>&1; echo "rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/ll_$R.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
cur={}
for r in rows[hi+1:]:
    cur.setdefault((r[0], r[ki].split('(')[0][-45:]), {})[r[mi]]=r[vi]
for k,v in cur.items():
    print(k[1], ' | '.join('%s=%s'%(m.split('.')[0][-22:],x) for m,x in v.items()))
PY

#!/bin/bash
# launch list of ALL kernels of the third evaluation (serialised durations under ncu; never a bench number)
mkdir -p gpurun_out
R=${1:-all}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/lla_$R.csv python scripts/profile_app.py 50000 3 > gpurun_out/lla_$R.log 2>&1; echo "rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/lla_$R.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
data=[(r[ki].split('(')[0][-50:], float(r[vi].replace(',',''))/1000.0) for r in rows[hi+1:] if len(r)>vi]
# last evaluation = after the last geom_kernel
idx=[i for i,d in enumerate(data) if 'geom_kernel' in d[0]]
ev=data[idx[-1]:]
tot=0
for n,t in ev:
    print('%-52s %8.1f us'%(n,t)); tot+=t
print('sum %.1f us over %d launches'%(tot,len(ev)))
PY

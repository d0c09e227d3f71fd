#!/bin/bash
# Round-end evidence: ncu launch list + full-set captures (chain kernel, AEV kernels, fused PME direct kernel), then the default bench line.
R=${1:-r08}
bash scripts/gpu_profile2.sh $R
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pme_direct_fused -s 2 -c 1 -o gpurun_out/prof_pme_$R -f python scripts/profile_pme_app.py > gpurun_out/prof_pme_$R.log 2>&1; echo "pme rc=$?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['value'], d['e2e']['value'], d['stage_ms'], d['roofline']['frac'], d['roofline']['frac_executed'], d.get('sustained',{}).get('value'), d.get('md',{}).get('value'))
print(d.get('pme'))"

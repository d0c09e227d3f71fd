#!/bin/bash
# ncu evidence (round 2 kernels): (1) launch list with per-launch device time, (2) full-set capture of the chain kernel and AEV kernels.
mkdir -p gpurun_out
R=${1:-r06}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python scripts/profile_app.py 50000 3 > gpurun_out/launches_$R.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 1 -c 1 -o gpurun_out/prof_gemm_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_gemm_$R.log 2>&1; echo "chain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ani_" -s 8 -c 8 -o gpurun_out/prof_aev_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_aev_$R.log 2>&1; echo "aev rc=$?"
ls -la gpurun_out/ | grep $R

#!/bin/bash
# ncu evidence: (1) launch list with per-launch device time, (2) full-set capture of the top kernels.  Never a bench number.
mkdir -p gpurun_out
R=${1:-r02}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python scripts/profile_app.py 50000 3 > gpurun_out/launches_$R.log 2>&1; echo "launch list rc=$?"
# one complete evaluation's 12 GEMM launches (the first 12 belong to evaluation 1), with DRAM bytes and tensor-pipe activity
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 12 -c 12 -o gpurun_out/prof_gemm_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_gemm_$R.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ani_" -s 5 -c 5 -o gpurun_out/prof_aev_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_aev_$R.log 2>&1; echo "aev rc=$?"
ls -la gpurun_out/ | grep $R

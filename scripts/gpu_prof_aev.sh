#!/bin/bash
# full-set ncu capture of the AEV kernels of the second evaluation
mkdir -p gpurun_out
R=${1:-tmp}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ani_" -s 9 -c 9 -o gpurun_out/prof_aev_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_aev_$R.log 2>&1; echo "aev rc=$?"
ls -la gpurun_out/ | grep prof_aev_$R

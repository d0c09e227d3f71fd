#!/bin/bash
# ncu evidence: (1) launch list with per-launch device time, (2) full-set capture of the top kernels.  Never a bench number.
mkdir -p gpurun_out
R=${1:-r01}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python scripts/profile_app.py 50000 3 > gpurun_out/launches_$R.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 18 -c 3 -o gpurun_out/prof_gemm_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_gemm_$R.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ani_angular -s 2 -c 2 -o gpurun_out/prof_angular_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_angular_$R.log 2>&1; echo "angular rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ani_radial|ani_rows" -s 3 -c 3 -o gpurun_out/prof_radial_$R -f python scripts/profile_app.py 50000 2 > gpurun_out/prof_radial_$R.log 2>&1; echo "radial rc=$?"
ls -la gpurun_out/

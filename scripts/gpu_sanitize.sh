#!/bin/bash
mkdir -p gpurun_out
export NNPOPS_MLP=tcgen05
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 gpurun_out/sanitize_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_cfconv_gpu.py tests/test_neighbors_pme_gpu.py -m gpu -q -x -k "golden or doctest or periodic or pme_random or too_many" > gpurun_out/sanitize_paths.log 2>&1; echo "memcheck paths rc=$?"; tail -4 gpurun_out/sanitize_paths.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 gpurun_out/racecheck_smoke.log

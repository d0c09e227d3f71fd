#!/bin/bash
# compute-sanitizer memcheck over the kernels changed in round 2's last stretch: the chain kernel (smoke), the fused PME direct kernel
# and getNeighborPairs (small-system tests)
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/sanitize_smoke.log
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_neighbors_pme_gpu.py -m gpu -q -x -k "fused_matches or doctest or neighbor_periodic or sharded_partials" > gpurun_out/sanitize_paths.log 2>&1; echo "memcheck paths rc=$?"; tail -3 gpurun_out/sanitize_paths.log
timeout 110 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -2 gpurun_out/racecheck_smoke.log
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_neighbors_pme_gpu.py -m gpu -q -x -k "fused_matches or doctest or neighbor_periodic or sharded_partials" > gpurun_out/racecheck_paths.log 2>&1; echo "racecheck paths rc=$?"; tail -2 gpurun_out/racecheck_paths.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_ani_gpu.py -m gpu -q -x -k "fused_chain_kernel" > gpurun_out/sanitize_chain.log 2>&1; echo "memcheck chain tests rc=$?"; tail -2 gpurun_out/sanitize_chain.log

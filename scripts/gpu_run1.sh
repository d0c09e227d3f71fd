#!/bin/bash
# first GPU pass: smoke, parity tests, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -s -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
echo "== bench simt"; timeout 900 python bench.py --mlp simt --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "bench rc=$?"; cat gpurun_out/bench_simt.json; tail -5 gpurun_out/bench_simt.err

#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05 parity"; timeout 300 python -m pytest tests -m gpu -q -s --timeout 200 -k "fused_energy_forces or host_entry" > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error|error|\{|rel" gpurun_out/pytest_tc.log | tail -30
echo "== bench tcgen05"; timeout 600 python bench.py --mlp tcgen05 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; cat gpurun_out/bench_tc.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','stage_ms')}); print(d['roofline'])
except Exception as e: print('no json', e)
"; tail -5 gpurun_out/bench_tc.err

#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_ani_gpu.py tests/test_optimized_torchani_gpu.py tests/test_torch_ops.py 2>&1 | tail -14
for G in 8 16 32; do
NNPOPS_ANGULAR_GROUP=$G timeout 300 python bench.py --no-cpu-baseline --steps 20 2> gpurun_out/aev_$G.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('G=$G', d['value'], d['stage_ms'])"
done

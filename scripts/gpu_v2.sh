#!/bin/bash
# angular v2 check: parity tests of the ANI paths, then A/B bench (v1 kernels via NNPOPS_ANGULAR_V1=1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ani_gpu.py tests/test_optimized_torchani_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -s -x --timeout 900 -k "not config_4 and not config_5 and not cfconv and not pme" > gpurun_out/pytest_v2.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|FAILED|ERROR|rel|errs|\{" gpurun_out/pytest_v2.log | tail -40
for V in 0 1; do
  if [ "$V" = "1" ]; then export NNPOPS_ANGULAR_V1=1; else unset NNPOPS_ANGULAR_V1; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --sustain 0 --md-steps 0 2> gpurun_out/v2_$V.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('V1=$V', d['value'], d['e2e']['value'], d['stage_ms'], d.get('forces_rel'))"
  tail -2 gpurun_out/v2_$V.err
done

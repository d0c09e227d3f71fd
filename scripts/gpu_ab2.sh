#!/bin/bash
# usage: gpu_ab2.sh <variant.so|base> ...  -- per-kernel ncu durations and one bench line per library variant
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset NNPOPS_LIB_PATH; else export NNPOPS_LIB_PATH=$PWD/nnpops_b200/variants/$v.so; fi
  echo "=== $v"
  bash scripts/gpu_ll.sh ab_$v | grep -v "^rc=" | head -9
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --sustain 0 --md-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], d['stage_ms'])"
done

#!/bin/bash
# usage: gpu_ab.sh "<ENV=VAL ...>" ...   -- one bench line per environment setting
mkdir -p gpurun_out
for cfg in "$@"; do
env $cfg timeout 300 python bench.py --no-cpu-baseline --steps 20 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['stage_ms'])"
done

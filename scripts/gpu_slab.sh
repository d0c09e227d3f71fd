#!/bin/bash
mkdir -p gpurun_out
for slab in 1152 2304 4736 9472; do
  NNPOPS_MLP_GRAPH=1 NNPOPS_MLP_SLAB=$slab timeout 300 python bench.py --no-cpu-baseline --steps 10 2> gpurun_out/slab_$slab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('slab $slab', d['value'], d['stage_ms'], d['gpu_launches'])"
done

#!/bin/bash
timeout 200 python -m pytest tests/test_neighbors_pme_gpu.py -x -q -k "energy_and_derivatives" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --sustain 0 --md-steps 0 --no-cpu-baseline 2> gpurun_out/b1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value']); print(json.dumps(d['pme'], indent=1))"

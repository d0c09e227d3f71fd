#!/bin/bash
# N = 8 bench line (conformers + the one-box strong-scaling key + the PME key)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/bench_n8.json'))
print(d['value'], d['e2e']['value']); print(d['box']['value'], d['box']['ms_per_step']); print(d['pme'])"

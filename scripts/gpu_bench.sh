#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600; tail -3 gpurun_out/bench_ref.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-1500; tail -5 gpurun_out/bench_n$N.err
fi

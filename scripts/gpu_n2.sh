#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "conformers rc=$?"; cut -c1-330 gpurun_out/bench_n$N.json; tail -2 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --mode box > gpurun_out/bench_box_n$N.json 2> gpurun_out/bench_box_n$N.err; echo "box rc=$?"; cut -c1-330 gpurun_out/bench_box_n$N.json; tail -2 gpurun_out/bench_box_n$N.err

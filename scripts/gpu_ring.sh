#!/bin/bash
# Development aid: the chain kernel with X in the activation buffer -- correctness against the per-layer path, then MLP stage time for
# several ring depths next to the previous build (nnpops_b200/variants/base.so).
mkdir -p gpurun_out
timeout 300 python scripts/chain_check.py 300,5000,50000 2>&1 | tail -12
for r in 8 6 4; do
  NNPOPS_CHAIN_RING=$r timeout 120 python scripts/chain_time.py 2>&1 | tail -1 | sed "s/^/ring=$r /"
done
NNPOPS_LIB_PATH=$PWD/nnpops_b200/variants/base.so timeout 120 python scripts/chain_time.py 2>&1 | tail -1

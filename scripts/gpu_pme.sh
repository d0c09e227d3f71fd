#!/bin/bash
# Development aid: PME tests + config-5 timing of the fused direct kernel for library variants
timeout 300 python -m pytest tests/test_neighbors_pme_gpu.py -x -q -k "pme" 2>&1 | tail -2
for v in default "$@"; do
  if [ "$v" = "default" ]; then unset NNPOPS_LIB_PATH; else export NNPOPS_LIB_PATH=$PWD/nnpops_b200/variants/$v.so; fi
  echo "== $v"; timeout 120 python scripts/bench_paths.py --pme-only 2>&1 | grep fused | cut -c1-140
done

#!/bin/bash
# Development aid: MLP stage time of library variants (nnpops_b200/variants/<name>.so; "default" = the in-tree build)
for v in "$@"; do
  if [ "$v" = "default" ]; then unset NNPOPS_LIB_PATH; else export NNPOPS_LIB_PATH=$PWD/nnpops_b200/variants/$v.so; fi
  timeout 120 python scripts/chain_time.py 2>&1 | tail -1
done

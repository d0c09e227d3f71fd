#!/bin/bash
# Development aid: the reference-CUDA comparators (tests + the bench key)
timeout 600 python -m pytest tests/test_reference_cuda_gpu.py tests/test_neighbors_pme_gpu.py -x -q -s -k "cfconv or pme_random or fused" 2>&1 | grep -E "passed|failed|Error|error|CFConv|PME random" | tail -12
timeout 300 python - <<'PY'
import sys, os, json
sys.path.insert(0, "tests")
import torch, bench
print(json.dumps(bench.cfconv_comparator(torch.device("cuda", 0)), indent=1))
PY

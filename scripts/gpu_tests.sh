#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest ${@:-tests} -m gpu -q -s --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|FAILED|ERROR|rel|errs|\{" gpurun_out/pytest.log | tail -60

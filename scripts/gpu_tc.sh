#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest_tc rc=$?" | tee -a gpurun_out/pytest_tc.log
grep -E "layer|embedding max|score_tc|passed|failed|Error|error|timeout" gpurun_out/pytest_tc.log | head -80

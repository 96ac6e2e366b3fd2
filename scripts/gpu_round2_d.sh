#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -k "layers_vs_oracle or scores_vs_fp32 or incremental_vs_window" > gpurun_out/d_tc.log 2>&1
echo "exit $?" ; grep -n "embedding max\|max |\|layer 1[0-9]\|FAILED\|passed\|failed" gpurun_out/d_tc.log | tail -40

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -k "incremental_vs_window or bench_configs_vs_oracle or partial_reset or gain_sweep" > gpurun_out/f_tc.log 2>&1
echo "exit $?"; grep -n "max |\|FAILED\|one-of\|B=\|passed\|failed\|Error" gpurun_out/f_tc.log | tail -30

#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --split-from 15 > gpurun_out/y_bench15.json 2> gpurun_out/y_bench15.err
OWW_SPLIT_FROM=15 timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s -k "gain_sweep or bench_configs or partial_reset" > gpurun_out/y_tc15.log 2>&1; echo "tc15 exit $?" > gpurun_out/y_status.txt
OWW_SPLIT_FROM=15 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/y_parity15.log 2>&1; echo "parity15 exit $?" >> gpurun_out/y_status.txt
python scripts/clocks_tree.py . 8192 15 > gpurun_out/y_clocks15.log 2>&1
cat gpurun_out/y_status.txt; grep -i "max\|worst\|gain" gpurun_out/y_tc15.log | head -20; tail -3 gpurun_out/y_tc15.log; tail -3 gpurun_out/y_parity15.log; head -3 gpurun_out/y_clocks15.log

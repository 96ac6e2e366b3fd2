#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "grouped_heads" -s > gpurun_out/n_grp.log 2>&1; echo "grp exit $?" > gpurun_out/n_status.txt
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/n_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/n_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/n_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/n_status.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/n_bench11.json 2> gpurun_out/n_bench11.err
python bench.py --steps 30 --warmup 5 --split-from 20 > gpurun_out/n_bench20.json 2> gpurun_out/n_bench20.err
cat gpurun_out/n_status.txt; grep -n "grouped vs" gpurun_out/n_grp.log; tail -5 gpurun_out/n_grp.log; tail -3 gpurun_out/n_tc.log; tail -3 gpurun_out/n_parity.log

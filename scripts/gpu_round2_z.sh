#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s > gpurun_out/z_tc.log 2>&1; echo "tc exit $?" > gpurun_out/z_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/z_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/z_status.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/z_bench11.json 2> gpurun_out/z_bench11.err
python bench.py --steps 30 --warmup 5 --split-from 15 > gpurun_out/z_bench15.json 2> gpurun_out/z_bench15.err
cat gpurun_out/z_status.txt; grep "max |score" gpurun_out/z_tc.log | head -12; tail -4 gpurun_out/z_tc.log; tail -4 gpurun_out/z_parity.log

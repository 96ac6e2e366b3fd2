#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s > gpurun_out/ab_tc.log 2>&1; echo "tc exit $?" > gpurun_out/ab_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/ab_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/ab_status.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/ab_bench11.json 2> gpurun_out/ab_bench11.err
python bench.py --steps 30 --warmup 5 --split-from 15 > gpurun_out/ab_bench15.json 2> gpurun_out/ab_bench15.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/ab_launches_c3.csv python bench.py --steps 3 --warmup 2 > gpurun_out/ab_ncu_bench.log 2>&1
cat gpurun_out/ab_status.txt; grep "max |score" gpurun_out/ab_tc.log | head -12; tail -4 gpurun_out/ab_tc.log; tail -4 gpurun_out/ab_parity.log

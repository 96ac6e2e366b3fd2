#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/clocks_tree.py . 8192 11 > gpurun_out/af_clocks11.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/af_tc.log 2>&1; echo "tc exit $?" > gpurun_out/af_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/af_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/af_status.txt
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/af_bench.json 2> gpurun_out/af_bench.err
cat gpurun_out/af_status.txt; head -3 gpurun_out/af_clocks11.log; tail -2 gpurun_out/af_tc.log; tail -2 gpurun_out/af_parity.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/w_multi.log 2>&1; echo "multi exit $?" > gpurun_out/w_status.txt
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s -k "split_from_variants" > gpurun_out/w_split.log 2>&1; echo "split exit $?" >> gpurun_out/w_status.txt
cat gpurun_out/w_status.txt; tail -3 gpurun_out/w_multi.log; grep "split_from=" gpurun_out/w_split.log; tail -2 gpurun_out/w_split.log

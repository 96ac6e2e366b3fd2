#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/h_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/h_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/h_status.txt
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -k "awkward or partial_reset" > gpurun_out/h_tc.log 2>&1
echo "tc exit $?" >> gpurun_out/h_status.txt
timeout 900 python scripts/bench_configs.py c5 c1 > gpurun_out/h_configs.json 2> gpurun_out/h_configs.err
echo "configs exit $?" >> gpurun_out/h_status.txt
cat gpurun_out/h_status.txt
grep -n "FAILED\|passed\|failed\|bulk_predict\|embed_clips" gpurun_out/h_parity.log | tail -12
grep -n "FAILED\|passed\|failed\|one-of\|awkward" gpurun_out/h_tc.log | tail -6
cat gpurun_out/h_configs.json; tail -5 gpurun_out/h_configs.err

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/ad_tc.log 2>&1; echo "tc exit $?" > gpurun_out/ad_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_metrics.py tests/test_ort_reference.py -x -q -m gpu > gpurun_out/ad_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/ad_status.txt
python bench.py --steps 40 --warmup 5 > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/ad_status.txt
cat gpurun_out/ad_status.txt; tail -3 gpurun_out/ad_tc.log; tail -3 gpurun_out/ad_parity.log; tail -2 gpurun_out/ad_smoke.log

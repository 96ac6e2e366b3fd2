#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/v_tc.log 2>&1; echo "tc exit $?" > gpurun_out/v_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_metrics.py tests/test_ort_reference.py -x -q -m gpu > gpurun_out/v_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/v_status.txt
python scripts/heads_clocks.py 1024 c2 > gpurun_out/v_clocks.txt 2>&1; python scripts/heads_clocks.py 8192 c3 >> gpurun_out/v_clocks.txt 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/v_bench11.json 2> gpurun_out/v_bench11.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/v_launches_c3.csv python bench.py --steps 3 --warmup 2 > gpurun_out/v_ncu_bench.log 2>&1
cat gpurun_out/v_status.txt; tail -3 gpurun_out/v_tc.log; tail -3 gpurun_out/v_parity.log; grep "^rep 2\|team 0" gpurun_out/v_clocks.txt

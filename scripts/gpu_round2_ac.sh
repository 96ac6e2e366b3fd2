#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/ac_tc.log 2>&1; echo "tc exit $?" > gpurun_out/ac_status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/ac_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/ac_status.txt
python bench.py --steps 40 --warmup 5 > gpurun_out/ac_bench_pdl.json 2> gpurun_out/ac_bench_pdl.err
OWW_FLAGS=32 python bench.py --steps 40 --warmup 5 > gpurun_out/ac_bench_nopdl.json 2> gpurun_out/ac_bench_nopdl.err
python bench.py --steps 40 --warmup 5 > gpurun_out/ac_bench_pdl2.json 2> gpurun_out/ac_bench_pdl2.err
cat gpurun_out/ac_status.txt; tail -3 gpurun_out/ac_tc.log; tail -3 gpurun_out/ac_parity.log

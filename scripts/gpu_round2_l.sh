#!/bin/bash
set -u
mkdir -p gpurun_out
( cd build_variants/r1_tree && python scripts/gpu_clocks.py 8192 > ../../gpurun_out/l_r1_clocks.log 2>&1; python scripts/bench_configs.py c3 > ../../gpurun_out/l_r1_c3.json 2>&1 )
python scripts/gpu_clocks.py 8192 20 > gpurun_out/l_now_clocks.log 2>&1
head -6 gpurun_out/l_r1_clocks.log; cat gpurun_out/l_r1_c3.json | tail -2; head -4 gpurun_out/l_now_clocks.log

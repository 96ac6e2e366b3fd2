#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/gpu_clocks.py 8192 11 > gpurun_out/i_clocks.log 2>&1
python scripts/gpu_clocks.py 8192 20 >> gpurun_out/i_clocks.log 2>&1
python scripts/gpu_clocks.py 1024 11 >> gpurun_out/i_clocks.log 2>&1
timeout 900 python scripts/bench_configs.py c5 > gpurun_out/i_c5.json 2> gpurun_out/i_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 60 -c 2 -o gpurun_out/i_prof_late -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/i_ncu.log 2>&1
cat gpurun_out/i_clocks.log; cat gpurun_out/i_c5.json; tail -3 gpurun_out/i_c5.err

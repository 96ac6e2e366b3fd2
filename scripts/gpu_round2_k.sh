#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/gpu_clocks.py 8192 20 > gpurun_out/k_clocks_slots16.log 2>&1
OWW_INC_SLOTS=4 python scripts/gpu_clocks.py 8192 20 > gpurun_out/k_clocks_slots4.log 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --split-from 20 --no-secondary > gpurun_out/k_fast16.json 2>/dev/null
OWW_INC_SLOTS=4 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --split-from 20 --no-secondary > gpurun_out/k_fast4.json 2>/dev/null
OWW_INC_SLOTS=4 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/k_prec4.json 2>/dev/null
head -3 gpurun_out/k_clocks_slots16.log; head -3 gpurun_out/k_clocks_slots4.log
python scripts/show_bench.py gpurun_out/k_fast16.json gpurun_out/k_fast4.json gpurun_out/k_prec4.json

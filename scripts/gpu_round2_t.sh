#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "grouped_heads" -s > gpurun_out/t_grp.log 2>&1; echo "grp exit $?" > gpurun_out/t_status.txt
python scripts/heads_clocks.py 1024 c2 > gpurun_out/t_clocks.txt 2>&1; python scripts/heads_clocks.py 8192 c3 >> gpurun_out/t_clocks.txt 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/t_bench11.json 2> gpurun_out/t_bench11.err
cat gpurun_out/t_status.txt; grep -n "grouped vs" gpurun_out/t_grp.log; tail -2 gpurun_out/t_grp.log; grep "^rep" gpurun_out/t_clocks.txt; grep -A12 "it: producer" gpurun_out/t_clocks.txt | head -14

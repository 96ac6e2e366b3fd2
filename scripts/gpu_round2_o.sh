#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "grouped_heads" -s > gpurun_out/o_grp.log 2>&1; echo "grp exit $?" > gpurun_out/o_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'heads_grp|feat16|gate' -c 60 --csv --log-file gpurun_out/o_launches_c3.csv python bench.py --steps 4 --warmup 2 > gpurun_out/o_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_grp -s 6 -c 1 -o gpurun_out/o_prof_heads_grp python scripts/bench_configs.py c3 > gpurun_out/o_ncu_full.log 2>&1
cat gpurun_out/o_status.txt; grep -n "grouped vs" gpurun_out/o_grp.log; tail -3 gpurun_out/o_grp.log; tail -12 gpurun_out/o_launches_c3.csv

#!/bin/bash
# final evidence of the round: bench line, launch list of the same command, --set full captures of the three kernel families
set -u
mkdir -p gpurun_out
python bench.py --steps 40 --warmup 5 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench exit $?" > gpurun_out/fin_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 3 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/fin_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_inc_kernel -s 5 -c 1 -o gpurun_out/fin_prof_inc python bench.py --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-secondary > gpurun_out/fin_ncu_inc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_blk -s 45 -c 9 -o gpurun_out/fin_prof_late python bench.py --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-secondary > gpurun_out/fin_ncu_late.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_grp -s 5 -c 1 -o gpurun_out/fin_prof_heads python bench.py --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-secondary > gpurun_out/fin_ncu_heads.log 2>&1
cat gpurun_out/fin_status.txt; ls -la gpurun_out/fin_*

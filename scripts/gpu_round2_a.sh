#!/bin/bash
# Round-2 GPU call A: full -m gpu suite (one pytest process per file, so a sticky CUDA error cannot poison the rest),
# smoke, default bench, launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/a_smi.txt 2>&1
for f in test_gpu_tc test_gpu_parity; do
  timeout 900 python -m pytest tests/$f.py -m gpu -x -q -s > gpurun_out/a_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/a_status.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/a_status.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?" >> gpurun_out/a_status.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-secondary --no-tc-heads > gpurun_out/a_bench_cc_heads.json 2> gpurun_out/a_bench_cc_heads.err
echo "bench cc heads exit $?" >> gpurun_out/a_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/a_ncu_bench.log 2>&1
echo "ncu exit $?" >> gpurun_out/a_status.txt
cat gpurun_out/a_status.txt
tail -3 gpurun_out/a_test_gpu_tc.log gpurun_out/a_test_gpu_parity.log gpurun_out/a_smoke.log
head -c 1500 gpurun_out/a_bench.json

#!/bin/bash
# Round-2 GPU call B: full -m gpu suite (every test runs, one process per file), then ncu --set full of the two step kernels at C3.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/b_status.txt
for f in test_gpu_tc test_gpu_parity test_metrics; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/b_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/b_status.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_tc_kernel -s 12 -c 2 -o gpurun_out/b_prof_heads_tc -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_heads.log 2>&1
echo "ncu heads exit $?" >> gpurun_out/b_status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_inc_kernel -s 12 -c 1 -o gpurun_out/b_prof_inc_c3 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_inc.log 2>&1
echo "ncu inc exit $?" >> gpurun_out/b_status.txt
cat gpurun_out/b_status.txt
for f in test_gpu_tc test_gpu_parity test_metrics; do tail -4 gpurun_out/b_$f.log; done

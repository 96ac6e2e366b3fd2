#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/j_status.txt
for f in test_gpu_tc test_gpu_parity; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/j_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/j_status.txt
done
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench exit $?" >> gpurun_out/j_status.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --split-from 20 > gpurun_out/j_bench_fast.json 2> gpurun_out/j_bench_fast.err
echo "bench fast exit $?" >> gpurun_out/j_status.txt
timeout 900 python scripts/bench_configs.py c5 > gpurun_out/j_c5.json 2> gpurun_out/j_c5.err
cat gpurun_out/j_status.txt
grep -n "FAILED\|passed\|failed" gpurun_out/j_test_gpu_tc.log gpurun_out/j_test_gpu_parity.log | tail -12
tail -3 gpurun_out/j_bench.err; cat gpurun_out/j_c5.json

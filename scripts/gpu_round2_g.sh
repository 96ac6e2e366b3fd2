#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/g_status.txt
for f in test_gpu_tc test_gpu_parity test_metrics; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/g_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/g_status.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/g_status.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?" >> gpurun_out/g_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 40 --csv --log-file gpurun_out/g_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/g_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/g_status.txt
cat gpurun_out/g_status.txt
grep -n "FAILED\|passed\|failed" gpurun_out/g_test_gpu_tc.log gpurun_out/g_test_gpu_parity.log | tail -12
tail -2 gpurun_out/g_smoke.log; tail -3 gpurun_out/g_bench.err

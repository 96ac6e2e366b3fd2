#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/e_status.txt
for f in test_gpu_tc test_gpu_parity; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/e_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/e_status.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/e_status.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench exit $?" >> gpurun_out/e_status.txt
cat gpurun_out/e_status.txt
grep -n "max |\|FAILED\|one-of\|B=\|passed\|failed\|Error" gpurun_out/e_test_gpu_tc.log | tail -40
grep -n "FAILED\|passed\|failed\|gated" gpurun_out/e_test_gpu_parity.log | tail -12
tail -2 gpurun_out/e_smoke.log; tail -3 gpurun_out/e_bench.err

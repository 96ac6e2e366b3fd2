#!/bin/bash
# Round-2 GPU call C: full -m gpu suite (one process per file), smoke, bench.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/c_status.txt
for f in test_gpu_tc test_gpu_parity test_metrics; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/c_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/c_status.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/c_status.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench exit $?" >> gpurun_out/c_status.txt
cat gpurun_out/c_status.txt
for f in test_gpu_tc test_gpu_parity test_metrics; do tail -5 gpurun_out/c_$f.log; done
tail -2 gpurun_out/c_smoke.log

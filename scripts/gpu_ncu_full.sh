#!/bin/bash
# one ncu --set full capture of the kernels of one step (skip warm-up steps)
set -u
mkdir -p gpurun_out
SKIP=${SKIP:-58}   # 2 reset + 2 full steps x 28 launches
COUNT=${COUNT:-28}
timeout 1200 ncu --set full --clock-control none --import-source on -s $SKIP -c $COUNT -f -o gpurun_out/prof_step \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep

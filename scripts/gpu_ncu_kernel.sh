#!/bin/bash
# ncu --set full on launches of one kernel (regex $KERNEL), skipping $SKIP matches
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_${TAG:-k} \
   python bench.py --steps 2 --warmup 4 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_${TAG:-k}.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_${TAG:-k}.ncu-rep

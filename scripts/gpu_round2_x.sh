#!/bin/bash
set -u
mkdir -p gpurun_out
for sp in 12 13 14 15 17; do
  python bench.py --steps 20 --warmup 4 --split-from $sp > gpurun_out/x_bench$sp.json 2> gpurun_out/x_bench$sp.err
done

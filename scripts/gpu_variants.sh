#!/bin/bash
# Try every library build under build_variants/ (see scripts/README.md): phase clocks + a short bench for each.
set -u
mkdir -p gpurun_out
cp openwakeword_b200/csrc/libowwb200.so /tmp/lib_keep.so
for f in build_variants/*.so; do
  cp "$f" openwakeword_b200/csrc/libowwb200.so
  echo "== $f"
  timeout 120 python scripts/gpu_clocks.py 2>&1 | tail -1
  timeout 200 python bench.py --no-cpu-baseline --steps 1000 > gpurun_out/b_v.json 2>gpurun_out/b_v.err && python scripts/show_bench.py gpurun_out/b_v.json | cut -c1-80
  timeout 200 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -1
done
cp /tmp/lib_keep.so openwakeword_b200/csrc/libowwb200.so

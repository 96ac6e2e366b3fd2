#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/clocks_tree.py build_variants/a0c5b51 8192 > gpurun_out/m_a0.log 2>&1
python scripts/clocks_tree.py . 8192 20 > gpurun_out/m_now20.log 2>&1
python scripts/clocks_tree.py . 8192 11 > gpurun_out/m_now11.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/m_tc.log 2>&1; echo "tc exit $?" > gpurun_out/m_status.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/m_bench11.json 2> gpurun_out/m_bench11.err
python bench.py --steps 30 --warmup 5 --split-from 20 > gpurun_out/m_bench20.json 2> gpurun_out/m_bench20.err
head -3 gpurun_out/m_a0.log gpurun_out/m_now20.log gpurun_out/m_now11.log; cat gpurun_out/m_status.txt; tail -3 gpurun_out/m_tc.log

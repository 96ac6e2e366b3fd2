#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py c1 c5 > gpurun_out/cfg_c1_c5.json 2> gpurun_out/cfg_c1_c5.err; echo "cfg exit $?" > gpurun_out/cfg_status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/cfg_ref_arm.json 2> gpurun_out/cfg_ref_arm.err; echo "ref exit $?" >> gpurun_out/cfg_status.txt
cat gpurun_out/cfg_status.txt; cat gpurun_out/cfg_c1_c5.json | cut -c1-900; tail -3 gpurun_out/cfg_c1_c5.err; cut -c1-600 gpurun_out/cfg_ref_arm.json

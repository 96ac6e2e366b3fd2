#!/bin/bash
# 2-GPU validation: peer-memory gather test, bench with the three gather modes
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/w_gpus.txt 2>&1
OWW_TEST_MULTI_GPU=1 timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/w_multi.log 2>&1; echo "multi exit $?" > gpurun_out/w_status.txt
for g in nccl-overlap nccl peer; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --gather $g > gpurun_out/w_bench2_$g.json 2> gpurun_out/w_bench2_$g.err; echo "bench2 $g exit $?" >> gpurun_out/w_status.txt
done
cat gpurun_out/w_status.txt; tail -5 gpurun_out/w_multi.log

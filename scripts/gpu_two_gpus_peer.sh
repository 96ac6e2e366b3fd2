#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/w_multi.log 2>&1; echo "multi exit $?" > gpurun_out/w_status.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --gather peer > gpurun_out/w_bench2_peer.json 2> gpurun_out/w_bench2_peer.err; echo "bench2 peer exit $?" >> gpurun_out/w_status.txt
cat gpurun_out/w_status.txt; tail -2 gpurun_out/w_multi.log

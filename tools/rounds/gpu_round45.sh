#!/bin/bash
# BASELINE configs[3]: CLIP ViT-L/14, K = 1000, synthetic 224x224 stream sharded over 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus 8 --model ViT-L/14 --batch 256 --steps 10 --warmup 3 --pool 2 --e2e-pool 2 > gpurun_out/bench_l14_n8.log 2>&1; echo "bench L/14 n8 exit $?"; tail -1 gpurun_out/bench_l14_n8.log | cut -c1-1400

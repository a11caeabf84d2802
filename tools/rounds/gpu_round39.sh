#!/bin/bash
# 8-GPU check of the contract launch (weak scaling, one process per GPU, one all-gather of the scores per stream)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "bench n8 exit $?"; tail -1 gpurun_out/bench_n8.log | cut -c1-1200
timeout 600 $TR tools/eval_synthetic.py --model ViT-L/14 --K 1000 --n-id 20000 --ood 10000,10000,10000,5640 > gpurun_out/eval_l14_n8.log 2>&1; echo "eval L/14 n8 exit $?"; tail -1 gpurun_out/eval_l14_n8.log | cut -c1-900

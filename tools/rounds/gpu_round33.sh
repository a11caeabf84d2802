#!/bin/bash
# 2-GPU check of the contract launch: bench.py (ours + reference arm) and the sharded synthetic evaluation
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?"; tail -1 gpurun_out/bench_n2.log | cut -c1-900
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "reference arm n2 exit $?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-400
timeout 600 $TR tools/eval_synthetic.py --model ViT-B/16 --K 1000 --n-id 5000 --ood 10000 > gpurun_out/eval_n2.log 2>&1; echo "eval n2 exit $?"; tail -1 gpurun_out/eval_n2.log | cut -c1-600
timeout 600 python tools/eval_synthetic.py --model ViT-B/16 --K 1000 --n-id 5000 --ood 10000 > gpurun_out/eval_n1.log 2>&1; echo "eval n1 exit $?"; tail -1 gpurun_out/eval_n1.log | cut -c1-600

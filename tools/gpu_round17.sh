#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_gpus2.log 2>&1; echo "bench 2gpu exit $?"; tail -1 gpurun_out/bench_gpus2.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_gpus2.log 2>&1; echo "ref 2gpu exit $?"; tail -1 gpurun_out/bench_ref_gpus2.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/eval_synthetic.py --model ViT-B/16 --K 1000 --n-id 5000 --ood 10000,5640 > gpurun_out/eval_gpus2.log 2>&1; echo "eval 2gpu exit $?"; tail -1 gpurun_out/eval_gpus2.log | cut -c1-700
timeout 600 python tools/eval_synthetic.py --model ViT-B/16 --K 1000 --n-id 5000 --ood 10000,5640 > gpurun_out/eval_gpus1.log 2>&1; echo "eval 1gpu exit $?"; tail -1 gpurun_out/eval_gpus1.log | cut -c1-700

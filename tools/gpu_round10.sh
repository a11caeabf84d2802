#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
python tools/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_atcpipe.so python tools/attn_sweep.py >> gpurun_out/attn_sweep.log 2>&1
MCM_ATTN_MMA=1 python tools/attn_sweep.py >> gpurun_out/attn_sweep.log 2>&1
cat gpurun_out/attn_sweep.log | grep '"S": 197' | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl_$i.log 2>&1; echo "bench pdl $i: $(tail -1 gpurun_out/bench_pdl_$i.log | cut -c60-120)"
MCM_NO_PDL=1 timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl_$i.log 2>&1; echo "bench nopdl $i: $(tail -1 gpurun_out/bench_nopdl_$i.log | cut -c60-120)"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "features" > gpurun_out/test_gpu_parity_features.log 2>&1; echo "features exit $?"; tail -4 gpurun_out/test_gpu_parity_features.log | cut -c1-300
timeout 600 python tools/eval_synthetic.py --model ViT-B/16 --K 1000 --n-id 5000 --ood 10000 > gpurun_out/eval_b16.log 2>&1; echo "eval exit $?"; tail -1 gpurun_out/eval_b16.log | cut -c1-600
timeout 600 python tools/eval_synthetic.py --model ViT-L/14 --K 1000 --n-id 2048 --ood 2048 --batch 128 > gpurun_out/eval_l14.log 2>&1; echo "eval L14 exit $?"; tail -1 gpurun_out/eval_l14.log | cut -c1-600
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --model ViT-L/14 --batch 128 > gpurun_out/bench_l14.log 2>&1; echo "bench L14: $(tail -1 gpurun_out/bench_l14.log | cut -c1-200)"

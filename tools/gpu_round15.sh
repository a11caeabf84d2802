#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention.log 2>&1; echo "attention exit $?"; tail -3 gpurun_out/test_gpu_attention.log | cut -c1-300
MCM_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention_split.log 2>&1; echo "attention split exit $?"; tail -3 gpurun_out/test_gpu_attention_split.log | cut -c1-300
echo "--- stagger"; python tools/attn_sweep.py 2>&1 | grep '"S"' | cut -c1-200
echo "--- stagger + split"; MCM_ATTN_SPLIT=1 python tools/attn_sweep.py 2>&1 | grep '"S"' | cut -c1-200
echo "--- no stagger"; MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_nostagger.so python tools/attn_sweep.py 2>&1 | grep '"S"' | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stagger_$i.log 2>&1; echo "bench stagger $i: $(tail -1 gpurun_out/bench_stagger_$i.log | cut -c60-100)"
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_nostagger.so timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nostagger_$i.log 2>&1; echo "bench nostagger $i: $(tail -1 gpurun_out/bench_nostagger_$i.log | cut -c60-100)"
done

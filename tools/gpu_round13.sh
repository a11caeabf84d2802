#!/bin/bash
mkdir -p gpurun_out
MCM_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention_split.log 2>&1; rc=$?; echo "attention split exit $rc"; tail -8 gpurun_out/test_gpu_attention_split.log | cut -c1-300
python tools/attn_sweep.py > gpurun_out/attn_sweep_ab.log 2>&1
if [ $rc -eq 0 ]; then
MCM_ATTN_SPLIT=1 python tools/attn_sweep.py >> gpurun_out/attn_sweep_ab.log 2>&1
fi
grep '"S"' gpurun_out/attn_sweep_ab.log | cut -c1-200
if [ $rc -eq 0 ]; then
for i in 1 2; do
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosplit_$i.log 2>&1; echo "bench nosplit $i: $(tail -1 gpurun_out/bench_nosplit_$i.log | cut -c60-100)"
MCM_ATTN_SPLIT=1 timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split_$i.log 2>&1; echo "bench split $i: $(tail -1 gpurun_out/bench_split_$i.log | cut -c60-100)"
done
MCM_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "golden or features" > gpurun_out/test_gpu_parity_split.log 2>&1; echo "parity split exit $?"; tail -3 gpurun_out/test_gpu_parity_split.log | cut -c1-300
fi

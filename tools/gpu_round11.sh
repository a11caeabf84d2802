#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16_$i.log 2>&1; echo "bench fp16 $i: $(tail -1 gpurun_out/bench_fp16_$i.log | cut -c60-100) $(tail -1 gpurun_out/bench_fp16_$i.log | grep -o '"clocks": {[^}]*}')"
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_bf16.so timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_$i.log 2>&1; echo "bench bf16 $i: $(tail -1 gpurun_out/bench_bf16_$i.log | cut -c60-100) $(tail -1 gpurun_out/bench_bf16_$i.log | grep -o '"clocks": {[^}]*}')"
done

#!/bin/bash
mkdir -p gpurun_out
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_trace.so python tools/attn_sweep.py > gpurun_out/attn_trace.log 2>&1
grep ATC_TRACE gpurun_out/attn_trace.log | head -48
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --batch 512 --e2e-pool 6 > gpurun_out/bench_b512.log 2>&1; echo "bench b512: $(tail -1 gpurun_out/bench_b512.log | cut -c60-100)"
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b256.log 2>&1; echo "bench b256: $(tail -1 gpurun_out/bench_b256.log | cut -c60-100) $(tail -1 gpurun_out/bench_b256.log | grep -o '"e2e": {[^}]*}')"
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --batch 512 --e2e-pool 6 > gpurun_out/bench_b512_2.log 2>&1; echo "bench b512: $(tail -1 gpurun_out/bench_b512_2.log | cut -c60-100) $(tail -1 gpurun_out/bench_b512_2.log | grep -o '"e2e": {[^}]*}')"

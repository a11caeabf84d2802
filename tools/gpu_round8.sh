#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_sweep.jsonl
export MCM_GEMM_PAIRS=2
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_gemm_pairs2.log 2>&1; rc=$?; echo "gemm pairs2 exit $rc"; tail -5 gpurun_out/test_gpu_gemm_pairs2.log | cut -c1-300
if [ $rc -eq 0 ]; then
timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_pairs2.log 2>&1; echo "sweep exit $?"; mv gpurun_out/gemm_sweep.jsonl gpurun_out/gemm_sweep_pairs2.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "golden" > gpurun_out/test_gpu_parity_pairs2.log 2>&1; echo "parity pairs2 exit $?"; tail -3 gpurun_out/test_gpu_parity_pairs2.log
timeout 900 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pairs2_b256.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_pairs2_b256.log | cut -c1-400
timeout 900 python bench.py --steps 80 --warmup 3 --no-cpu-baseline --batch 192 > gpurun_out/bench_pairs2_b192.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_pairs2_b192.log | cut -c1-400
fi
unset MCM_GEMM_PAIRS
timeout 900 python bench.py --steps 80 --warmup 3 --no-cpu-baseline --batch 192 > gpurun_out/bench_pairs1_b192.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_pairs1_b192.log | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/gemm_sweep.jsonl
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x -k "layernorm_fold" > gpurun_out/test_gpu_gemm.log 2>&1; rc=$?; echo "gemm exit $rc"; tail -3 gpurun_out/test_gpu_gemm.log | cut -c1-300
export SWEEP_CASES="768,768,6;768,3072,6"
MCM_GEMM_RESID_TMA=1 SWEEP_TAG=tma timeout 300 python tools/gemm_sweep.py 2>&1 | cut -c1-100
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_gtrace.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/ncu_step.py --steps 1 --batch 512 2>&1 | grep GEMM_TRACE > gpurun_out/gemm_trace.log
sed -n 100,104p gpurun_out/gemm_trace.log
for v in 1 2; do
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$v.log 2>&1; echo "bench $v: $(tail -1 gpurun_out/bench_$v.log | cut -c60-100)"
tail -1 gpurun_out/bench_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print(d['clocks'], d['e2e']['value'], d['config'].get('batch_per_gpu'))"
done

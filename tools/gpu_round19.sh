#!/bin/bash
# LN-folded GEMM epilogues: unit tests, full parity, sweep, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_gpu_gemm.log 2>&1; rc=$?; echo "gemm exit $rc"; tail -15 gpurun_out/test_gpu_gemm.log | cut -c1-300
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_rowwise.py tests/test_gpu_tail.py tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_rest.log 2>&1; echo "rest exit $?"; tail -15 gpurun_out/test_gpu_rest.log | cut -c1-300
rm -f gpurun_out/gemm_sweep.jsonl; timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; cut -c1-120 gpurun_out/gemm_sweep.jsonl
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$i.log 2>&1; echo "bench $i: $(tail -1 gpurun_out/bench_$i.log | cut -c60-100)"; done
tail -1 gpurun_out/bench_2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print(d['clocks'], d['e2e']['value'], d['config'].get('batch_per_gpu'))"
fi

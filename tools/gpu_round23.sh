#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -x -k "not fullsize" > gpurun_out/test_gpu_main.log 2>&1; rc=$?; echo "tests exit $rc"; tail -5 gpurun_out/test_gpu_main.log | cut -c1-300
cp gpurun_out/parity_report.jsonl gpurun_out/parity_report_tanh.jsonl; rm -f gpurun_out/parity_report.jsonl
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_geluexact.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -x -k "not fullsize" > gpurun_out/test_gpu_exact.log 2>&1; echo "exact-gelu parity exit $?"
cp gpurun_out/parity_report.jsonl gpurun_out/parity_report_exact.jsonl
echo "--- tanh"; cat gpurun_out/parity_report_tanh.jsonl | cut -c1-200
echo "--- exact"; cat gpurun_out/parity_report_exact.jsonl | cut -c1-200
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_gtrace.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/ncu_step.py --steps 1 --batch 512 2>&1 | grep GEMM_TRACE > gpurun_out/gemm_trace.log
sed -n 100,104p gpurun_out/gemm_trace.log
if [ $rc -eq 0 ]; then
for v in default geluexact; do
lib=$PWD/mcm_b200/_C/libmcm_b200_$v.so; [ $v = default ] && lib=$PWD/mcm_b200/_C/libmcm_b200.so
MCM_B200_LIB=$lib timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$v.log 2>&1; echo "bench $v: $(tail -1 gpurun_out/bench_$v.log | cut -c60-100)"
tail -1 gpurun_out/bench_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print(d['clocks'], d['e2e']['value'], d['config'].get('batch_per_gpu'))"
done
fi

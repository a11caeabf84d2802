#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention.log 2>&1; rc=$?; echo "attention exit $rc"; tail -3 gpurun_out/test_gpu_attention.log | cut -c1-300
python tools/attn_sweep.py 2>&1 | grep '"S"' | cut -c1-200
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_trace.so python tools/attn_sweep.py > gpurun_out/attn_trace.log 2>&1
grep ATC_TRACE gpurun_out/attn_trace.log | grep -v -- "-1       -1       -1" | head -30
if [ $rc -eq 0 ]; then
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_rest.log 2>&1; echo "api+parity exit $?"; tail -3 gpurun_out/test_gpu_rest.log | cut -c1-300
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$i.log 2>&1; echo "bench $i: $(tail -1 gpurun_out/bench_$i.log | cut -c60-100)"; done
tail -1 gpurun_out/bench_2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print(d['clocks'], d['e2e']['value'])"
fi

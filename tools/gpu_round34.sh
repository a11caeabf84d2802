#!/bin/bash
# fp16 epilogue stores through dedicated store warps (variant sw) vs TMA bulk stores (default)
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
MCM_B200_LIB=$L/libmcm_b200_sw.so timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_gpu_gemm_sw.log 2>&1; echo "gemm(sw) exit $?"; tail -4 gpurun_out/test_gpu_gemm_sw.log | cut -c1-300
MCM_B200_LIB=$L/libmcm_b200_sw.so timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_main_sw.log 2>&1; echo "api+parity(sw) exit $?"; tail -4 gpurun_out/test_gpu_main_sw.log | cut -c1-300
for v in gtrace swtrace; do
for c in "3072,768,1" "2304,768,4" "3072,768,5"; do
  echo -n "$v: "; SWEEP_CASES="$c" MCM_B200_LIB=$L/libmcm_b200_$v.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/gemm_sweep.py 2>&1 | grep -E "GEMM_TRACE" | tail -1 | cut -c1-220
done
done
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['e2e_uint8']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for v in "" _sw "" _sw; do
  MCM_B200_LIB=$L/libmcm_b200$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench$v.log 2>&1; echo "B/16 lib '$v': $(summ gpurun_out/bench$v.log)"
done

#!/bin/bash
# fp16 epilogues with direct 32-byte stores (default) vs staging + TMA bulk stores (variant tmast)
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_gpu_gemm.log 2>&1; echo "gemm exit $?"; tail -4 gpurun_out/test_gpu_gemm.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_main.log 2>&1; echo "api+parity exit $?"; tail -4 gpurun_out/test_gpu_main.log | cut -c1-300
rm -f gpurun_out/gemm_skip.log
for c in "3072,768,1" "2304,768,0" "3072,768,5" "2304,768,4"; do
  SWEEP_CASES="$c" MCM_B200_LIB=$L/libmcm_b200_gtrace.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/gemm_sweep.py 2>&1 | grep -E "GEMM_TRACE" | tail -1 | cut -c1-220
done
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for v in "" _tmast "" _tmast; do
  MCM_B200_LIB=$L/libmcm_b200$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench$v.log 2>&1; echo "B/16 lib '$v': $(summ gpurun_out/bench$v.log)"
done
timeout 900 python bench.py --model ViT-L/14 --batch 256 --steps 8 --pool 2 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_l14.log 2>&1; echo "L/14: $(summ gpurun_out/bench_l14.log)"

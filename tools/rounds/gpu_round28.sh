#!/bin/bash
# tail-row warp (ViT-L/14), single-pass softmax (variant sp), 3 residual buffers in the out_proj epilogue (variant rb3)
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
for v in "" _sp; do
  MCM_B200_LIB=$L/libmcm_b200$v.so timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention$v.log 2>&1
  echo "attention tests lib '$v' exit $?"; tail -4 gpurun_out/test_gpu_attention$v.log | cut -c1-300
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "features" > gpurun_out/test_gpu_parity.log 2>&1
echo "parity(features) exit $?"; tail -4 gpurun_out/test_gpu_parity.log | cut -c1-300
MCM_B200_LIB=$L/libmcm_b200_sp.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_parity_sp.log 2>&1
echo "parity+api sp exit $?"; tail -4 gpurun_out/test_gpu_parity_sp.log | cut -c1-300
rm -f gpurun_out/attn_sweep.log
for v in "" _sp; do
  MCM_B200_LIB=$L/libmcm_b200$v.so timeout 300 python tools/attn_sweep.py 2>&1 | cut -c1-160 >> gpurun_out/attn_sweep.log
done
cat gpurun_out/attn_sweep.log
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for v in "" _sp _rb3 _sprb3 "" _sprb3; do
  MCM_B200_LIB=$L/libmcm_b200$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench$v.log 2>&1; echo "B/16 lib '$v': $(summ gpurun_out/bench$v.log)"
done
for v in "" _sp; do
MCM_B200_LIB=$L/libmcm_b200$v.so timeout 900 python bench.py --model ViT-L/14 --batch 256 --steps 8 --pool 2 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_l14$v.log 2>&1; echo "L/14 lib '$v': $(summ gpurun_out/bench_l14$v.log)"
done

#!/bin/bash
# attention: score chunks kept in registers between the two softmax passes (TMEM read port = 64 B/clk is the bound)
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention.log 2>&1; echo "attention tests exit $?"; tail -3 gpurun_out/test_gpu_attention.log | cut -c1-300
MCM_B200_LIB=$L/libmcm_b200_keep3.so timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention3.log 2>&1; echo "attention tests keep3 exit $?"; tail -3 gpurun_out/test_gpu_attention3.log | cut -c1-300
for v in _keep0 _keep1 "" _keep3 _keep0 ""; do
SWEEP_SHAPES="256,197,12;128,257,16;256,50,12" MCM_B200_LIB=$L/libmcm_b200$v.so timeout 300 python tools/attn_sweep.py 2>&1 | cut -c40-160
done
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for v in _keep0 "" _keep3; do
  MCM_B200_LIB=$L/libmcm_b200$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench$v.log 2>&1; echo "B/16 lib '$v': $(summ gpurun_out/bench$v.log)"
done

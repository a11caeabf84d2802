#!/bin/bash
# 257-token attention with the extra token's q / k / v in shared memory (x box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention.log 2>&1
echo "attention tests exit $?"; tail -4 gpurun_out/test_gpu_attention.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "features or uint8" > gpurun_out/test_gpu_parity.log 2>&1
echo "parity(features) exit $?"; tail -4 gpurun_out/test_gpu_parity.log | cut -c1-300
SWEEP_SHAPES="256,197,12;128,256,16;128,257,16" timeout 300 python tools/attn_sweep.py 2>&1 | cut -c1-160
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
timeout 900 python bench.py --model ViT-L/14 --batch 256 --steps 8 --pool 2 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_l14.log 2>&1; echo "L/14: $(summ gpurun_out/bench_l14.log)"
timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "B/16: $(summ gpurun_out/bench.log)"

#!/bin/bash
# attention pair mode (two items per unit for S <= 64, ViT-B/32)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention.log 2>&1; echo "attention tests exit $?"; tail -6 gpurun_out/test_gpu_attention.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "features" > gpurun_out/test_gpu_parity.log 2>&1; echo "parity(features) exit $?"; tail -4 gpurun_out/test_gpu_parity.log | cut -c1-300
SWEEP_SHAPES="256,50,12;1024,50,12;256,197,12" timeout 300 python tools/attn_sweep.py 2>&1 | cut -c1-160
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
timeout 600 python bench.py --model ViT-B/32 --batch 2048 --steps 20 --pool 1 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_b32.log 2>&1; echo "B/32: $(summ gpurun_out/bench_b32.log)"

#!/bin/bash
# out_proj residual epilogue A/B (TMA vs LSU) under the final code, and a ViT-B/32 datapoint
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for v in 1 0 1 0; do
  MCM_GEMM_RESID_TMA=$v timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_rt$v.log 2>&1; echo "B/16 resid_tma=$v: $(summ gpurun_out/bench_rt$v.log)"
done
timeout 600 python bench.py --model ViT-B/32 --batch 2048 --steps 20 --pool 1 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_b32.log 2>&1; echo "B/32: $(summ gpurun_out/bench_b32.log)"

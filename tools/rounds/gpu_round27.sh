#!/bin/bash
# attention for 257 tokens, uint8 ingest, smem-staged patchify: tests, then batch sweep + ViT-L/14 bench + GEMM cycle trace
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
for f in tests/test_gpu_attention.py tests/test_gpu_rowwise.py tests/test_gpu_api.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$f exit $?"; tail -4 gpurun_out/$n.log | cut -c1-300
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_parity.log 2>&1
echo "parity exit $?"; tail -4 gpurun_out/test_gpu_parity.log | cut -c1-300
timeout 300 python tools/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1
SWEEP_SHAPES="128,257,16" MCM_ATTN_MMA=1 timeout 300 python tools/attn_sweep.py >> gpurun_out/attn_sweep.log 2>&1
cat gpurun_out/attn_sweep.log
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"; }
for b in 256 512 768 1024; do
  timeout 600 python bench.py --batch $b --steps 20 --no-cpu-baseline > gpurun_out/bench_b$b.log 2>&1; echo "batch $b: $(summ gpurun_out/bench_b$b.log)"
done
timeout 900 python bench.py --model ViT-L/14 --batch 256 --steps 8 --pool 2 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_l14.log 2>&1; echo "L/14: $(summ gpurun_out/bench_l14.log)"
MCM_ATTN_MMA=1 timeout 900 python bench.py --model ViT-L/14 --batch 256 --steps 8 --pool 2 --e2e-pool 2 --no-cpu-baseline > gpurun_out/bench_l14_mma.log 2>&1; echo "L/14 mma.sync attention: $(summ gpurun_out/bench_l14_mma.log)"
if [ -f mcm_b200/_C/libmcm_b200_gtrace.so ]; then
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_gtrace.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/ncu_step.py --steps 1 --batch 512 2>&1 | grep GEMM_TRACE > gpurun_out/gemm_trace.log
sed -n 100,104p gpurun_out/gemm_trace.log
fi

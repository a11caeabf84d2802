#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/gemm_sweep.jsonl
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_gpu_gemm.log 2>&1; rc=$?; echo "gemm exit $rc"; tail -5 gpurun_out/test_gpu_gemm.log | cut -c1-300
MCM_GEMM_TMA_STORE=1 timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_gpu_gemm_tma.log 2>&1; rc2=$?; echo "gemm tma-store exit $rc2"; tail -5 gpurun_out/test_gpu_gemm_tma.log | cut -c1-300
export SWEEP_CASES="3072,768,0;3072,768,4;3072,768,5;2304,768,4;768,768,6;768,3072,6;3072,64,0;3072,64,5"
SWEEP_TAG=lsu timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_lsu.log 2>&1
MCM_GEMM_TMA_STORE=1 SWEEP_TAG=tma timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_tma.log 2>&1
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/gemm_sweep.jsonl')]
cases=[]
for r in rows:
    k=(r['N'],r['K'],r['epi'])
    if k not in cases: cases.append(k)
tags=[]
for r in rows:
    if r['tag'] not in tags: tags.append(r['tag'])
print('case'.ljust(18)+''.join(t.rjust(9) for t in tags))
for c in cases:
    print(str(c).ljust(18)+''.join(('%.1f'%[r['us'] for r in rows if (r['N'],r['K'],r['epi'])==c and r['tag']==t][0]).rjust(9) for t in tags))
PY
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "not fullsize" > gpurun_out/test_gpu_rest.log 2>&1; echo "rest exit $?"; tail -4 gpurun_out/test_gpu_rest.log | cut -c1-300
for v in 0 1; do
MCM_GEMM_TMA_STORE=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_tma$v.log 2>&1; echo "bench tma_store=$v: $(tail -1 gpurun_out/bench_tma$v.log | cut -c60-100)"
tail -1 gpurun_out/bench_tma$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print(d['clocks'], d['e2e']['value'], d['config'].get('batch_per_gpu'))"
done
fi

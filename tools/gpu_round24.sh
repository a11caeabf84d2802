#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/gemm_sweep.jsonl
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_gpu_gemm.log 2>&1; rc=$?; echo "gemm exit $rc"; tail -12 gpurun_out/test_gpu_gemm.log | cut -c1-300
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -x -k "not fullsize" > gpurun_out/test_gpu_main.log 2>&1; rc=$?; echo "tests exit $rc"; tail -5 gpurun_out/test_gpu_main.log | cut -c1-300
grep features gpurun_out/parity_report.jsonl | cut -c1-120
export SWEEP_CASES="3072,768,4;3072,768,5;2304,768,4;768,768,6;768,3072,6"
MCM_GEMM_RESID_TMA=0 SWEEP_TAG=lsu timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_lsu.log 2>&1
MCM_GEMM_RESID_TMA=1 SWEEP_TAG=tma timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_tma.log 2>&1
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
MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_gtrace.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/ncu_step.py --steps 1 --batch 512 2>&1 | grep GEMM_TRACE > gpurun_out/gemm_trace.log
sed -n 100,104p gpurun_out/gemm_trace.log
for v in 0 1 x; do
[ $v = x ] && unset MCM_GEMM_RESID_TMA || export MCM_GEMM_RESID_TMA=$v
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_rt$v.log 2>&1; echo "bench resid_tma=$v: $(tail -1 gpurun_out/bench_rt$v.log | cut -c60-100)"
tail -1 gpurun_out/bench_rt$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print(d['clocks'], d['e2e']['value'], d['config'].get('batch_per_gpu'))"
done
fi

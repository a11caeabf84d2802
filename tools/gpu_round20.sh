#!/bin/bash
# where does the GEMM epilogue time go: sweep with parts of the epilogue switched off + ncu of one layer's GEMMs
mkdir -p gpurun_out
rm -f gpurun_out/gemm_sweep.jsonl
export SWEEP_CASES="3072,768,0;3072,768,4;3072,768,5;2304,768,4;768,768,2;768,768,6;768,3072,6;3072,1536,0"
for skip in 0 1 2 4 8 9 12; do
  MCM_GEMM_DBG_SKIP=$skip SWEEP_TAG=skip$skip timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_skip$skip.log 2>&1
done
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
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16 -s 13 -c 4 -f -o gpurun_out/prof_gemm python tools/ncu_step.py --steps 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"

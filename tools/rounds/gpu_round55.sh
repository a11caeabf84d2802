#!/bin/bash
# programmatic dependent launch A/B under the final kernels
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['step_frac'],4))"; }
for v in 0 1 0 1; do
  MCM_PDL=$v timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_pdl$v.log 2>&1; echo "B/16 MCM_PDL=$v: $(summ gpurun_out/bench_pdl$v.log)"
done

#!/bin/bash
# why is the 257-token attention unit 2x slower than the 197-token one?  phase trace + ncu source-level capture
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
SWEEP_SHAPES="128,257,16" MCM_B200_LIB=$L/libmcm_b200_atrace.so timeout 300 python tools/attn_sweep.py > gpurun_out/atrace_l14.log 2>&1
SWEEP_SHAPES="256,197,12" MCM_B200_LIB=$L/libmcm_b200_atrace.so timeout 300 python tools/attn_sweep.py > gpurun_out/atrace_b16.log 2>&1
SWEEP_SHAPES="128,256,16" timeout 300 python tools/attn_sweep.py 2>&1 | cut -c1-200
SWEEP_SHAPES="128,257,16" timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 5 -c 1 -f -o gpurun_out/prof_attn_l14 python tools/attn_sweep.py > gpurun_out/ncu_attn_l14.log 2>&1; echo "ncu exit $?"
grep -c ATC_TRACE gpurun_out/atrace_l14.log gpurun_out/atrace_b16.log

#!/bin/bash
mkdir -p gpurun_out
export MCM_B200_LIB=$PWD/mcm_b200/_C/libmcm_b200_gtrace.so
MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/ncu_step.py --steps 1 --batch 512 2>&1 | grep GEMM_TRACE > gpurun_out/gemm_trace.log
sed -n 100,112p gpurun_out/gemm_trace.log

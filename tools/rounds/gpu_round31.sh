#!/bin/bash
# is the GEMM main loop bound by shared-memory bandwidth (operand writes + reads + epilogue staging)?
# real SM cycles per tile of the MMA issuer (clock64 trace build) with the epilogue progressively removed
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
rm -f gpurun_out/gemm_skip.log
for skip in 0 1 4; do
  for c in "3072,768,1" "2304,768,0" "768,3072,2"; do
    echo "skip=$skip case=$c" >> gpurun_out/gemm_skip.log
    SWEEP_CASES="$c" SWEEP_TAG=skip$skip MCM_GEMM_DBG_SKIP=$skip MCM_B200_LIB=$L/libmcm_b200_gtrace.so MCM_GEMM_TRACE_PRINT=1 timeout 300 python tools/gemm_sweep.py 2>&1 | grep -E "GEMM_TRACE|tflops" | tail -2 | cut -c1-220 >> gpurun_out/gemm_skip.log
  done
done
cat gpurun_out/gemm_skip.log

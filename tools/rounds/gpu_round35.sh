#!/bin/bash
# attention: max tree in pass 1 (default) and software-pipelined pass 2 (variant pipe)
mkdir -p gpurun_out
L=$PWD/mcm_b200/_C
for v in "" _pipe; do
MCM_B200_LIB=$L/libmcm_b200$v.so timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attention$v.log 2>&1
echo "attention tests '$v' exit $?"; tail -3 gpurun_out/test_gpu_attention$v.log | cut -c1-300
done
for v in "" _pipe "" _pipe; do
SWEEP_SHAPES="256,197,12;128,257,16" MCM_B200_LIB=$L/libmcm_b200$v.so timeout 300 python tools/attn_sweep.py 2>&1 | cut -c40-160
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tail.py tests/test_gpu_maha.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_tail.log 2>&1; echo "tail exit $?"; tail -12 gpurun_out/test_gpu_tail.log | cut -c1-300

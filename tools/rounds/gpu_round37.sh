#!/bin/bash
# device Resize + CenterCrop: parity tests and a first timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_api.log 2>&1; echo "api exit $?"; tail -15 gpurun_out/test_gpu_api.log | cut -c1-300
timeout 600 python tools/resize_bench.py 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_api.py tests/test_gpu_maha.py -q -m gpu -p no:cacheprovider -k "resize or uint8 or maha" > gpurun_out/sanitizer_api.log 2>&1; echo "memcheck api exit $?"; tail -4 gpurun_out/sanitizer_api.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_api.py -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_api.log 2>&1; echo "api exit $?"; tail -2 gpurun_out/test_gpu_api.log

#!/bin/bash
# compute-sanitizer memcheck over the smoke path, the attention shapes (incl. 257 tokens) and the device preprocess
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "memcheck smoke exit $?"; tail -4 gpurun_out/sanitizer_smoke.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_attention.py -q -m gpu -p no:cacheprovider -k "257 or 197 or 50-2" > gpurun_out/sanitizer_attn.log 2>&1; echo "memcheck attention exit $?"; tail -4 gpurun_out/sanitizer_attn.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_api.py -q -m gpu -p no:cacheprovider -k "resize or uint8" > gpurun_out/sanitizer_api.log 2>&1; echo "memcheck api exit $?"; tail -4 gpurun_out/sanitizer_api.log | cut -c1-200

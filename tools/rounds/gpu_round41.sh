#!/bin/bash
# Mahalanobis baseline on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_maha.py tests/test_gpu_tail.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_maha.log 2>&1; echo "maha exit $?"; tail -15 gpurun_out/test_gpu_maha.log | cut -c1-300
grep maha gpurun_out/parity_report.jsonl | tail -2

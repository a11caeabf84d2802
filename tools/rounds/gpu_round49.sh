#!/bin/bash
# robustness of the LayerNorm fold against outlier ("massive activation") channels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider -k "outlier" > gpurun_out/test_outlier.log 2>&1; echo "exit $?"; tail -12 gpurun_out/test_outlier.log | cut -c1-300
grep outliers gpurun_out/parity_report.jsonl

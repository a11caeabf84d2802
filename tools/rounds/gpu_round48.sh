#!/bin/bash
# last check of round 1: the driver's own sequence (pytest -m gpu, smoke, reference arm, bench)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/test_gpu_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-300

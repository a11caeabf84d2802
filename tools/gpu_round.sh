#!/bin/bash
# One GPU-box session: per-kernel diagnostics, the -m gpu tests (one process per file so a device
# trap cannot cascade), smoke, bench, ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.csv 2>&1
timeout 900 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1
echo "gemm_diag exit $?"
for f in tests/test_gpu_gemm.py tests/test_gpu_rowwise.py tests/test_gpu_attention.py tests/test_gpu_tail.py tests/test_gpu_api.py tests/test_gpu_parity.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/$n.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1500

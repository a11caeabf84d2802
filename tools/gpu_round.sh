#!/bin/bash
# One GPU-box session: the -m gpu tests (one process per file so a device trap cannot cascade),
# smoke, bench, ncu launch list + full captures of the top kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/gemm_sweep.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.csv 2>&1
for f in tests/test_gpu_attention.py tests/test_gpu_gemm.py tests/test_gpu_rowwise.py tests/test_gpu_tail.py tests/test_gpu_api.py tests/test_gpu_parity.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/$n.log | cut -c1-300
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 60 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1800
if [ "$1" != "noncu" ]; then
timeout 600 python tools/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1
timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; echo "sweep exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16 -s 14 -c 4 -f -o gpurun_out/prof_gemm python tools/ncu_step.py --steps 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_ -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/ncu_step.py --steps 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
fi

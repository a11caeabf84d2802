#!/bin/bash
# round-1 final: all -m gpu tests, smoke, bench (+ reference arm), ncu launch list and full captures at the bench batch size
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/nvsmi.csv 2>&1
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/test_gpu_all.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "reference arm exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-2500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py --steps 1 --batch 512 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16 -s 14 -c 4 -f -o gpurun_out/prof_gemm python tools/ncu_step.py --steps 1 --batch 512 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tcgen05 -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/ncu_step.py --steps 1 --batch 512 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"

#!/bin/bash
# round 2, call B: GPU tests after the kernel refinements, bench, ncu --set full of the step's SELL kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py > gpurun_out/bench_fused.log 2>&1; echo "bench fused rc=$?"
timeout 600 python bench.py --coefficient variable --no-cpu-baseline --no-e2e > gpurun_out/bench_variable.log 2>&1; echo "bench variable rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:sell_kernel -c 4 -o gpurun_out/step_sell_full_first4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_full1.log 2>&1; echo "ncu full first rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:sell_kernel -s 44 -c 4 -o gpurun_out/step_sell_full_last4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_full2.log 2>&1; echo "ncu full last rc=$?"
ls -la gpurun_out/
grep -h '^{' gpurun_out/bench_fused.log gpurun_out/bench_variable.log | cut -c1-300

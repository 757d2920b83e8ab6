#!/bin/bash
# round 2, call A: GPU tests, fused vs plain bench, launch list of one step
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1; echo "bench fused rc=$?"
timeout 600 env MGB_CYCLE_FUSION=0 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_unfused.log 2>&1; echo "bench unfused rc=$?"
timeout 600 env MGB_CYCLE_FUSION=0 MGB_IMPLIED_COLUMNS=0 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_plain.log 2>&1; echo "bench plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu rc=$?"
grep -h '^{' gpurun_out/bench_fused.log gpurun_out/bench_unfused.log gpurun_out/bench_plain.log | cut -c1-400

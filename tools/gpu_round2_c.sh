#!/bin/bash
# round 2, call C: GPU tests (PCG on device, short-row kernel, diagonal gather skipped, arrive/sync reduction), bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1; echo "bench fused rc=$?"
for r in 1 4; do
timeout 600 env MGB_SHORT_ROWS=$r python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_short$r.log 2>&1; echo "bench short rows $r rc=$?"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 python tools/bench_setup_pcg.py --sizes 2048,8192 > gpurun_out/setup_pcg.log 2>&1; echo "setup_pcg rc=$?"
grep -h '^{' gpurun_out/bench_fused.log gpurun_out/bench_short1.log gpurun_out/bench_short4.log | cut -c1-200
grep -h '^{' gpurun_out/setup_pcg.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d.pop('galerkin'); print(d)"

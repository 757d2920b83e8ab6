#!/bin/bash
# 1-GPU check: GPU tests, bench.py (constant and variable coefficient), ncu launch list of one step with DRAM bytes.
# Usage: gpurun --timeout 3000 -- 'bash tools/gpu_single.sh'
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_spec.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --coefficient variable --no-cpu-baseline --no-e2e > gpurun_out/bench_variable.log 2>&1; echo "bench variable rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
for f in bench_spec bench_variable; do grep -h '^{' gpurun_out/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']; print('$f', round(d['ms_per_step'], 4), 'sweep', d['roofline']['ms_per_launch'], 'setup', c['setup_s'], c.get('setup_phases_s'), 'frac', round(d['roofline']['cycle']['frac'], 3), 'moved', d['roofline']['cycle'].get('moved_frac'), d['e2e'])"; done

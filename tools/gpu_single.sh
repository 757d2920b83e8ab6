#!/bin/bash
# 1-GPU check: GPU tests, bench.py (constant, variable, symmetrically eliminated variable coefficient), ncu launch list
# of one step with DRAM bytes, ncu --set full of the level-0 kernels.  Usage: gpurun --timeout 3000 -- 'bash tools/gpu_single.sh'
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "bench reference rc=$? $(tail -c 300 gpurun_out/bench_reference.log)"
timeout 600 python bench.py --coefficient variable --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_variable.log 2>&1; echo "bench variable rc=$?"
timeout 600 python bench.py --coefficient variable-symmetric --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_varsym.log 2>&1; echo "bench varsym rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:sell_ -c 4 -o gpurun_out/step_sell_full_first4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_full1.log 2>&1; echo "ncu full first rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:sell_ -s 44 -c 4 -o gpurun_out/step_sell_full_last4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_full2.log 2>&1; echo "ncu full last rc=$?"
for f in bench bench_variable bench_varsym; do grep -h '^{' gpurun_out/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']; print('$f', round(d['ms_per_step'], 4), 'sweep', d['roofline']['ms_per_launch'], 'setup', c['setup_s'], 'frac', round(d['roofline']['cycle']['frac'], 3), 'moved', d['roofline']['cycle'].get('moved_frac'), d['e2e'], d['cpu_baseline'])"; done

#!/bin/bash
# round 2, call D: value dictionary
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_dict.log 2>&1; echo "bench dict rc=$?"
timeout 600 python bench.py --coefficient variable --no-cpu-baseline --no-e2e > gpurun_out/bench_variable.log 2>&1; echo "bench variable rc=$?"
timeout 600 env MGB_VALUE_DICT=0 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_nodict.log 2>&1; echo "bench nodict rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:sell_ -c 4 -o gpurun_out/step_sell_full_first4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_full1.log 2>&1; echo "ncu full first rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:sell_ -s 44 -c 4 -o gpurun_out/step_sell_full_last4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --profile-step > gpurun_out/ncu_full2.log 2>&1; echo "ncu full last rc=$?"
grep -h '^{' gpurun_out/bench_dict.log gpurun_out/bench_variable.log gpurun_out/bench_nodict.log | cut -c1-200

#!/bin/bash
# 1-GPU check of the round's last changes next to its final evidence: GPU tests, bench.py (constant and variable
# coefficient) with setup phases, the same bench with every library under learnmultigrid_b200/_variants/
# (tools/build_variant.sh) swapped in, the sub-steps of the level-0 Galerkin product, ncu launch list of one step with
# DRAM bytes, ncu --set full of the level-0 kernels.
# Usage: gpurun --timeout 1100 -- 'bash tools/gpu_setup_round.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc $(tail -n 1 gpurun_out/pytest_gpu.log)"
if [ $rc -ne 0 ]; then
  # the newest kernel variant (implied values) has a switch: is everything else still green without it?
  tail -n 60 gpurun_out/pytest_gpu.log
  export MGB_IMPLIED_VALUES=0
  timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_without_implied_values.log 2>&1; echo "pytest (MGB_IMPLIED_VALUES=0) rc=$? $(tail -n 1 gpurun_out/pytest_gpu_without_implied_values.log)"
fi
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --coefficient variable --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_variable.log 2>&1; echo "bench variable rc=$?"
MGB_IMPLIED_VALUES=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_variant_no_implied_values.log 2>&1; echo "bench without implied values rc=$?"
cp learnmultigrid_b200/libmgb200.so /tmp/libmgb200_shipped.so
for lib in $(ls learnmultigrid_b200/_variants/ 2>/dev/null | sed 's/libmgb200_//; s/\.so//'); do
  cp learnmultigrid_b200/_variants/libmgb200_$lib.so learnmultigrid_b200/libmgb200.so
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_variant_$lib.log 2>&1; echo "bench variant $lib rc=$?"
done
cp /tmp/libmgb200_shipped.so learnmultigrid_b200/libmgb200.so
for f in bench bench_variable $(ls gpurun_out | grep '^bench_variant_' | sed 's/\.log//'); do grep -h '^{' gpurun_out/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']; print('$f', round(d['ms_per_step'], 4), 'sweep', round(d['roofline']['ms_per_launch'], 4), 'setup', c['setup_s'], c['setup_phases_s'], 'gen', c['generate_s'], 'res', c['residual_after_timed_steps'], 'e2e', (d.get('e2e') or {}).get('value'), 'pin', c.get('pinned_staging_alloc_s'), [(g['level'], g['ms'], g['gb_per_s']) for g in (c.get('setup_galerkin_per_level') or [])])
    x = c.get('extra') or {}
    for k, v in x.items():
        print('  extra', k, {kk: vv for kk, vv in v.items() if kk in ('ms_per_step', 'setup_s', 'setup_phases_s', 'ms_per_iteration', 'iterations', 'hierarchy_setup_s', 'solve_ms')} if isinstance(v, dict) else v)"; done
timeout 200 python tools/time_galerkin.py --n 8192 > gpurun_out/time_galerkin.log 2>&1; echo "time_galerkin rc=$? $(tail -n 1 gpurun_out/time_galerkin.log)"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:sell_ -c 4 -o gpurun_out/step_sell_full_first4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_full1.log 2>&1; echo "ncu full first rc=$?"
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:sell_ -s 44 -c 4 -o gpurun_out/step_sell_full_last4 python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_full2.log 2>&1; echo "ncu full last rc=$?"

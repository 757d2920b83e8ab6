#!/bin/bash
# N-GPU run (N = $1) on one box: multi-process parity check, then bench.py (and with "variable" as $2 the
# variable-coefficient configuration C4).  Usage: gpurun --gpus 8 -- 'bash tools/gpu_multi.sh 8 variable' 
set -u
n=${1:-4}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run; export n
timeout 300 bash -c "run 29511 tools/dist_check.py --size 512 --levels 5 --n-dist 3" > gpurun_out/dist_check_${n}gpu.log 2>&1; echo "dist_check rc=$? $(tail -n 1 gpurun_out/dist_check_${n}gpu.log)"
timeout 900 bash -c "run 29513 bench.py --gpus $n --steps 40 --no-cpu-baseline" > gpurun_out/bench_${n}gpu.log 2>&1; echo "bench ${n}gpu rc=$?"
if [ "${2:-}" = "variable" ]; then
timeout 900 bash -c "run 29515 bench.py --gpus $n --steps 40 --coefficient variable --no-cpu-baseline --no-e2e --no-parity-check" > gpurun_out/bench_${n}gpu_variable.log 2>&1; echo "bench ${n}gpu variable rc=$?"
fi
for f in gpurun_out/bench_${n}gpu*.log; do grep -h '^{' $f | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']; print('$f', round(d['ms_per_step'], 4), 'dry', c.get('ms_per_step_without_exchange_waits'), 'launches', d['gpu_launches'] / d['steps'], 'setup', c['setup_s'], c.get('multi_rank_parity'), c['residual_after_timed_steps'], d['e2e'])"; done

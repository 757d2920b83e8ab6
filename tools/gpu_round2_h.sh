#!/bin/bash
# round 2, 2-GPU call: all GPU tests (incl. the 2-process one), final 2-GPU bench, weak-scaling probe, symmetric-operator probe
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run
timeout 900 bash -c "run 29513 bench.py --gpus 2 --steps 40 --no-cpu-baseline" > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu rc=$?"
timeout 600 bash -c "run 29516 tools/bench_weak.py --n 4096 --steps 20 --colors structured" > gpurun_out/weak_2gpu_structured.log 2>&1; echo "weak structured rc=$? $(tail -c 600 gpurun_out/weak_2gpu_structured.log)"
timeout 600 bash -c "run 29517 tools/bench_weak.py --n 4096 --steps 20" > gpurun_out/weak_2gpu_greedy.log 2>&1; echo "weak greedy rc=$? $(tail -c 600 gpurun_out/weak_2gpu_greedy.log)"
timeout 600 python bench.py --coefficient variable-symmetric --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_varsym.log 2>&1; echo "bench varsym rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_varsym.csv python bench.py --coefficient variable-symmetric --steps 3 --no-cpu-baseline --no-e2e --no-extra --profile-step > gpurun_out/ncu_varsym.log 2>&1; echo "ncu varsym rc=$?"
for f in gpurun_out/bench_2gpu.log gpurun_out/bench_varsym.log; do grep -h '^{' $f | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']; print('$f', round(d['ms_per_step'], 4), 'dry', c.get('ms_per_step_without_exchange_waits'), 'setup', c['setup_s'], c.get('multi_rank_parity'), d['e2e']); print(json.dumps(c.get('extra'))[:1500])"; done

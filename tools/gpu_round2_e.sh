#!/bin/bash
# round 2, call E (2 GPUs): all GPU tests incl. the 2-process one, multi-process parity, 2-GPU bench both exchange modes
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); run 29511 tools/dist_check.py --size 512 --levels 5 --n-dist 3" > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$? $(tail -n 1 gpurun_out/dist_check.log)"
timeout 300 bash -c "$(declare -f run); MGB_PUSH_EXCHANGE=1 run 29512 tools/dist_check.py --size 512 --levels 5 --n-dist 3" > gpurun_out/dist_check_push.log 2>&1; echo "dist_check push rc=$? $(tail -n 1 gpurun_out/dist_check_push.log)"
timeout 900 bash -c "$(declare -f run); run 29513 bench.py --gpus 2 --steps 40 --no-cpu-baseline" > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu rc=$?"
timeout 900 bash -c "$(declare -f run); MGB_PUSH_EXCHANGE=1 run 29514 bench.py --gpus 2 --steps 40 --no-cpu-baseline --no-e2e" > gpurun_out/bench_2gpu_push.log 2>&1; echo "bench 2gpu push rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_1gpu.log 2>&1; echo "bench 1gpu rc=$?"
grep -h '^{' gpurun_out/bench_2gpu.log gpurun_out/bench_2gpu_push.log gpurun_out/bench_1gpu.log | cut -c1-250

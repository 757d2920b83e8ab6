#!/bin/bash
# weak-scaling line on N GPUs ($1) with C cells per side and GPU ($2): strip-local setup, structured and first-fit colours
set -u
n=${1:-2}; c=${2:-4096}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run; export n
timeout 900 bash -c "run 29516 tools/bench_weak.py --cells $c --steps 20 --colors structured" > gpurun_out/weak_${n}gpu_${c}_structured.log 2>&1; echo "weak structured rc=$? $(grep '^{' gpurun_out/weak_${n}gpu_${c}_structured.log | cut -c1-700)"; tail -n 3 gpurun_out/weak_${n}gpu_${c}_structured.log | cut -c1-300
if [ "${3:-}" = "greedy" ]; then
timeout 900 bash -c "run 29517 tools/bench_weak.py --cells $c --steps 20" > gpurun_out/weak_${n}gpu_${c}_greedy.log 2>&1; echo "weak greedy rc=$? $(grep '^{' gpurun_out/weak_${n}gpu_${c}_greedy.log | cut -c1-700)"; tail -n 3 gpurun_out/weak_${n}gpu_${c}_greedy.log | cut -c1-300
fi

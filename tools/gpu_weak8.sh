#!/bin/bash
# weak scaling on 8 GPUs: 4097^2 per GPU always; 8193^2 per GPU (537 M DOF) only if the host has the memory for the
# host-side strip setup (8 ranks x ~25 GB)
set -u
mkdir -p gpurun_out
echo "cores $(nproc)"; free -g | head -2
mem=$(free -g | awk '/^Mem:/{print $7}')
bash tools/gpu_weak.sh 8 4096
if [ "$mem" -ge 400 ]; then
  run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
  export -f run
  timeout 480 bash -c "run 29518 tools/bench_weak.py --cells 8192 --steps 20 --colors structured" > gpurun_out/weak_8gpu_8192_structured.log 2>&1; echo "weak 8192 rc=$? $(grep '^{' gpurun_out/weak_8gpu_8192_structured.log | cut -c1-900)"; tail -n 3 gpurun_out/weak_8gpu_8192_structured.log | cut -c1-300
else
  echo "host memory ${mem} GB: 8193^2 per GPU skipped"
fi

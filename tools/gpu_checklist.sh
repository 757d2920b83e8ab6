#!/bin/bash
# First GPU calls of the next round, in the order of DESIGN 12 ("Work done after the last GPU minute").  Every step
# writes its log under gpurun_out/ and is bounded by its own timeout, so that one failing step costs minutes, not the
# call.  Usage on the GPU box (one GPU unless stated):
#     gpurun --timeout 1500 -- 'bash tools/gpu_checklist.sh single'
#     gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_checklist.sh multi 8'
set -u
mkdir -p gpurun_out
step() {   # step <name> <timeout_s> <command...>
    local name=$1 t=$2; shift 2
    echo "== $name" | tee -a gpurun_out/checklist.log
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
    echo "   rc=$? ($(tail -n 1 "gpurun_out/$name.log" | cut -c1-160))" | tee -a gpurun_out/checklist.log
}
case "${1:-single}" in
single)
    step pytest_gpu 900 python -m pytest tests -m gpu -x -q
    step pytest_unverified 600 env MGB_UNVERIFIED=1 python -m pytest tests/test_gpu_strip.py -m gpu -q
    step bench_default 600 python bench.py
    step bench_implied_columns 600 env MGB_IMPLIED_COLUMNS=1 python bench.py --no-cpu-baseline
    step setup_host_colours 900 python tools/bench_setup_pcg.py
    step setup_device_colours 900 env MGB_DEVICE_COLORS=1 python tools/bench_setup_pcg.py
    ;;
multi)
    n=${2:-8}
    run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 "$@"; }
    step dist_check 300 run tools/dist_check.py --size 512 --levels 5 --n-dist 3
    step dist_check_push 300 env MGB_PUSH_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py --size 512 --levels 5 --n-dist 3
    step bench_consumer 600 run bench.py --gpus "$n" --steps 40 --no-cpu-baseline
    step bench_push 600 env MGB_PUSH_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus "$n" --steps 40 --no-cpu-baseline
    step bench_implied 600 env MGB_IMPLIED_COLUMNS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus "$n" --steps 40 --no-cpu-baseline
    step bench_push_implied 600 env MGB_PUSH_EXCHANGE=1 MGB_IMPLIED_COLUMNS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus "$n" --steps 40 --no-cpu-baseline
    step bench_weak 900 run tools/bench_weak.py --n 8192 --steps 20 --colors structured
    ;;
*)
    echo "usage: $0 single | multi N"; exit 2 ;;
esac
cat gpurun_out/checklist.log

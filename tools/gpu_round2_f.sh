#!/bin/bash
# round 2, call F (1 GPU): tests, device colouring with setup phases, C3 / C2 / C4-style configurations
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
timeout 600 env MGB_DEVICE_COLORS=1 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_devcolors.log 2>&1; echo "bench device colours rc=$?"
timeout 600 python bench.py --n 4096 --transfer quasi --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_quasi.log 2>&1; echo "bench c3 quasi rc=$?"
timeout 600 python bench.py --n 4096 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_linear.log 2>&1; echo "bench c3 linear rc=$?"
timeout 600 python bench.py --mesh irregular --transfer nn --n 1024 --levels 4 --nu 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"
timeout 600 python bench.py --smoother Jacobi --no-cpu-baseline --no-e2e > gpurun_out/bench_jacobi.log 2>&1; echo "bench jacobi rc=$?"
timeout 900 python tools/bench_setup_pcg.py --sizes 8192 > gpurun_out/setup_pcg.log 2>&1; echo "setup_pcg rc=$?"
for f in bench_devcolors bench_c3_quasi bench_c3_linear bench_c2 bench_jacobi; do grep -h '^{' gpurun_out/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']; print('$f', round(d['ms_per_step'], 4), 'setup', c['setup_s'], c.get('setup_phases_s'), 'frac', round(d['roofline']['cycle']['frac'], 3), 'moved', d['roofline']['cycle'].get('moved_frac'))"; done

#!/bin/bash
# Kernel experiment on one GPU: the headline bench (constant and variable coefficient, device-timed part only) with the
# shipped library and with each variant under learnmultigrid_b200/_variants/ (tools/build_variant.sh) swapped in.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_variants.sh'
set -u
mkdir -p gpurun_out
cp learnmultigrid_b200/libmgb200.so /tmp/libmgb200_shipped.so
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/variant_pytest.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/variant_pytest.log)"
for lib in shipped $(ls learnmultigrid_b200/_variants/ 2>/dev/null | sed 's/libmgb200_//; s/\.so//'); do
  if [ $lib = shipped ]; then cp /tmp/libmgb200_shipped.so learnmultigrid_b200/libmgb200.so
  else cp learnmultigrid_b200/_variants/libmgb200_$lib.so learnmultigrid_b200/libmgb200.so; fi
  for coef in constant variable; do
    timeout 300 python bench.py --coefficient $coef --no-cpu-baseline --no-e2e --no-extra > gpurun_out/variant_${lib}_$coef.log 2>&1
    echo "$lib $coef rc=$? $(grep -h '^{' gpurun_out/variant_${lib}_$coef.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['ms_per_step'], 4), 'sweep', d['roofline']['ms_per_launch'], 'res', d['config'].get('residual_after_timed_steps'), d['clocks'])")"
  done
done
cp /tmp/libmgb200_shipped.so learnmultigrid_b200/libmgb200.so

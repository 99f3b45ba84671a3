#!/bin/bash
# time every build/variants/*.so on the cfg-2 step (layer-major path only)
mkdir -p gpurun_out
: > gpurun_out/variants.log
for f in build/variants/*.so; do
  echo "== $f" >> gpurun_out/variants.log
  RENI_ONLY_LBWD=1 RENI_B200_LIB=$PWD/$f python tools/step_phases.py 32 20 >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log

#!/bin/bash
# two-term forward weights: precision sweep + timing, one- vs two-term
mkdir -p gpurun_out
for t in 1 2; do
  echo "== RENI_FWD_TERMS=$t" 
  RENI_FWD_TERMS=$t python tools/precision_sweep.py 2>&1 | grep -E "seed|rror" 
  RENI_FWD_TERMS=$t RENI_ONLY_LBWD=1 RENI_TILE_MAJOR_BWD=1 python tools/step_phases.py 32 20 2>&1 | tail -3
done > gpurun_out/d_terms.log 2>&1
cat gpurun_out/d_terms.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/d_gpu_tests.log
cat gpurun_out/d_gpu_tests.log

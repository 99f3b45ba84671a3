#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "layer_major" 2>&1 | tail -5 > gpurun_out/b_lbwd_test.log
cat gpurun_out/b_lbwd_test.log
python tools/step_phases.py 32 20 > gpurun_out/b_phases.log 2>&1
cat gpurun_out/b_phases.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/b_gpu_tests.log
cat gpurun_out/b_gpu_tests.log

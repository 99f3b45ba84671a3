#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/i_gpu_tests.log
cat gpurun_out/i_gpu_tests.log
python tools/shade_time.py > gpurun_out/r2_shade_time.txt 2>&1; cat gpurun_out/r2_shade_time.txt
python tools/ab_grid.py > gpurun_out/r2_ab_grid.txt 2>&1; tail -8 gpurun_out/r2_ab_grid.txt

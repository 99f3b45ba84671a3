#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/e_gpu_tests.log
cat gpurun_out/e_gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -5 gpurun_out/e_bench.err
cat gpurun_out/e_bench.json

#!/bin/bash
# first GPU call of round 2: layer-major backward A/B + parity suite + bench (old vs new path)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "layer_major" 2>&1 | tail -15 > gpurun_out/a_lbwd_test.log
cat gpurun_out/a_lbwd_test.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/a_gpu_tests.log
cat gpurun_out/a_gpu_tests.log
python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_lbwd.json 2> gpurun_out/a_bench_lbwd.err
cat gpurun_out/a_bench_lbwd.json
RENI_TILE_MAJOR_BWD=1 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_tile.json 2> gpurun_out/a_bench_tile.err
cat gpurun_out/a_bench_tile.json

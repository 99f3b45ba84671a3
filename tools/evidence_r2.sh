#!/bin/bash
# Round-2 evidence set (run under gpurun; outputs land in gpurun_out/ and are summarised into profiles/ here)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.json
timeout 300 python tools/config_sweep.py > gpurun_out/r2_sweep.log 2>&1; tail -8 gpurun_out/r2_sweep.log
timeout 300 python tools/precision_sweep.py > gpurun_out/r2_precision_sweep.txt 2>&1; tail -4 gpurun_out/r2_precision_sweep.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"reni_(fwd|bwd|dw)_kernel" -s 9 -c 3 -o gpurun_out/r2_prof -f python tools/profile_step.py 5 > gpurun_out/p.log 2>&1; tail -2 gpurun_out/p.log

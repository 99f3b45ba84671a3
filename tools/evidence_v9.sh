#!/bin/bash
# Round-1 v9 evidence set (run under gpurun; outputs land in gpurun_out/ and are copied to profiles/ by hand)
set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r1_v9_bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/r1_v9_bench.json
timeout 200 python tools/config_sweep.py > gpurun_out/sweep.log 2>&1; tail -8 gpurun_out/sweep.log
timeout 100 python tools/step_time.py > gpurun_out/r1_v9_step_time.txt 2>&1; cat gpurun_out/r1_v9_step_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_v9_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"reni_(fwd|bwd|dw)_kernel" -s 9 -c 3 -o gpurun_out/r1_v9_prof -f python tools/profile_step.py 5 > gpurun_out/p.log 2>&1; tail -2 gpurun_out/p.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"reni_bwd_kernel" -s 3 -c 1 -o gpurun_out/r1_v9_prof_latent -f python tools/profile_step.py 5 latent > gpurun_out/p3.log 2>&1; tail -2 gpurun_out/p3.log

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reni_lbwd -s 18 -c 6 -f -o gpurun_out/r2_lbwd_prof python tools/step_phases.py 32 2 > gpurun_out/ncu_lbwd.log 2>&1
tail -5 gpurun_out/ncu_lbwd.log
ls -la gpurun_out/*.ncu-rep

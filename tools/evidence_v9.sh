set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py > gpurun_out/r1_v9_bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/r1_v9_bench.json
timeout 200 python tools/film_time.py > gpurun_out/r1_v9_film_time.json 2> gpurun_out/film_time.err
timeout 200 python tools/precision_sweep.py > gpurun_out/r1_v9_precision_sweep.txt 2>&1; tail -5 gpurun_out/r1_v9_precision_sweep.txt
timeout 100 python tools/decode_latency.py > gpurun_out/decode_latency.log 2>&1; tail -8 gpurun_out/decode_latency.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_v9_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"reni_(fwd|bwd|dw)_kernel|reni_film" -s 24 -c 8 -o gpurun_out/r1_v9_prof_film -f python tools/profile_step.py 5 film > gpurun_out/p4.log 2>&1; tail -3 gpurun_out/p4.log

#!/bin/bash
# multi-GPU checks: 2-rank numerics test, then bench at N GPUs with the in-graph exchange and with NCCL
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/mg_topo.log
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_parallel_gpu.py -x -q -m gpu -s 2>&1 | tail -30 > gpurun_out/mg_tests.log
cat gpurun_out/mg_tests.log
for ex in auto nccl; do
  if [ $ex = nccl ]; then export RENI_EXCHANGE=nccl; else unset RENI_EXCHANGE; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/mg_bench_${N}gpu_$ex.json 2> gpurun_out/mg_bench_${N}gpu_$ex.err
  tail -3 gpurun_out/mg_bench_${N}gpu_$ex.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/mg_bench_${N}gpu_$ex.json'))
print('$ex', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'exchange', d.get('exchange'), d['config']['step'])
"
done

#!/bin/bash
# same-box scaling check: 1-GPU bench, then N-GPU bench with in-graph exchange and NCCL (short runs)
N=${1:-2}
CFG=${2:-cfg2}
STEPS=${3:-50}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu-baseline --config $CFG > gpurun_out/sb_${CFG}_1gpu.json 2> gpurun_out/sb_${CFG}_1gpu.err
for ex in auto nccl; do
  if [ $ex = nccl ]; then export RENI_EXCHANGE=nccl; else unset RENI_EXCHANGE; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 --config $CFG > gpurun_out/sb_${CFG}_${N}gpu_$ex.json 2> gpurun_out/sb_${CFG}_${N}gpu_$ex.err
done
python - <<PY
import json
for name in ["1gpu", "${N}gpu_auto", "${N}gpu_nccl"]:
    try:
        txt = open(f"gpurun_out/sb_${CFG}_{name}.json").read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(name, "ms/step", round(d["ms_per_step"], 4), "M dir/s", round(d["value"] / 1e6, 1), "sustained", d["sustained_ms_per_step"], "e2e ms", round(d["e2e"]["ms_per_step"], 4), d.get("exchange"))
    except Exception as e:
        print(name, "ERR", e)
PY

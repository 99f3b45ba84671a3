#!/bin/bash
# same-box scaling: bench at 1 GPU and at N GPUs.   tools/gpu_scale.sh N cfg steps "auto nccl"
N=${1:-2}; CFG=${2:-cfg2}; STEPS=${3:-50}; EXS=${4:-auto}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu-baseline --config $CFG > gpurun_out/sc_${CFG}_1of${N}.json 2> gpurun_out/sc_${CFG}_1of${N}.err
for ex in $EXS; do
  if [ $ex = nccl ]; then export RENI_EXCHANGE=nccl; else unset RENI_EXCHANGE; fi
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 --config $CFG > gpurun_out/sc_${CFG}_${N}gpu_$ex.json 2> gpurun_out/sc_${CFG}_${N}gpu_$ex.err
done
python - <<PY
import json
base = None
for name in ["1of${N}"] + ["${N}gpu_" + e for e in "$EXS".split()]:
    try:
        txt = open(f"gpurun_out/sc_${CFG}_{name}.json").read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        if base is None: base = d["value"]
        print("${CFG}", name, "ms/step", round(d["ms_per_step"], 4), "M dir/s", round(d["value"] / 1e6, 1), "x", round(d["value"]/base, 3), "sustained", d["sustained_ms_per_step"], "e2e ms", round(d["e2e"]["ms_per_step"], 4), d.get("exchange"))
    except Exception as e:
        print(name, "ERR", e)
PY

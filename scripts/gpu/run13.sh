#!/bin/bash
# 2-GPU: why is the device-resident DP loop slower than the e2e loop?  + CUPTI timelines (1 GPU)
mkdir -p gpurun_out
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29540 \
    bench.py --gpus 2 --no-cpu --no-extra $BARGS > gpurun_out/dpx_$name.json 2> gpurun_out/dpx_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dpx_$name.json").read().strip().splitlines()[-1])
    print("$name", "resident", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("$name ERR", e)
PY
}
BARGS="--steps 20 --warmup 5" run default A=1
BARGS="--steps 20 --warmup 40" run warm40 A=1
BARGS="--steps 20 --warmup 5" run nograph SMD_TRAIN_GRAPH=0
BARGS="--steps 20 --warmup 5" run nooverlap SMD_DP_OVERLAP=0
BARGS="--steps 20 --warmup 5" run nograph_nooverlap SMD_TRAIN_GRAPH=0 SMD_DP_OVERLAP=0
timeout 200 python scripts/timeline.py train gpurun_out/r02_timeline_train.json 2>&1 | tail -2
SMD_TRAIN_GRAPH=0 timeout 200 python scripts/timeline.py train gpurun_out/r02_timeline_train_eager.json 2>&1 | tail -2
timeout 200 python scripts/timeline.py sample gpurun_out/r02_timeline_sample.json 2>&1 | tail -2

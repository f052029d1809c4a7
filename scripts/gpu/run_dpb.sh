#!/bin/bash
# usage (under gpurun --gpus N): bash scripts/gpu/run_dpb.sh N   -- the N-GPU bench line exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_train_${N}gpu.json 2> gpurun_out/bench_dpb${N}.err ) 2>&1 | grep real
tail -c 400 gpurun_out/bench_dpb${N}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_train_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=$N ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "value", d["value"], "div", d.get("dp_rank_divergence"), "rel", d.get("dp_vs_single_rel_l2"), d["config"].get("untimed_settle_steps"))
    for e in d.get("extra", []):
        print(" ", e["name"], round(e["ms_per_step"], 4), round(e["step_frac_of_sustained_peak"], 3))
except Exception as e:
    print("ERR", e)
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29545 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_${N}gpu.json 2>> gpurun_out/bench_dpb${N}.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02_bench_reference_arm_${N}gpu.json

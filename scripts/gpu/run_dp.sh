#!/bin/bash
# usage (under gpurun --gpus N): bash scripts/gpu/run_dp.sh N
# data-parallel equivalence worker + the N-GPU bench line (with the dp_* proof keys and the extra configs)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/dp${N}_gpus.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29533 \
  tests/dp_equivalence_worker.py > gpurun_out/r02_dp_equivalence_${N}gpu.log 2>&1
echo "dp worker exit=$?"; grep -E "world|graph step|dp-ok|Error|error" gpurun_out/r02_dp_equivalence_${N}gpu.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_train_${N}gpu.json 2> gpurun_out/bench_dp${N}.err
echo "bench exit=$?"; tail -c 500 gpurun_out/bench_dp${N}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_train_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=$N ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "div", d.get("dp_rank_divergence"), "rel", d.get("dp_vs_single_rel_l2"), "dloss", d.get("dp_vs_single_dloss_rel"))
    for e in d.get("extra", []):
        print(" ", e["name"], round(e["ms_per_step"], 4), round(e["step_frac_of_sustained_peak"], 3))
except Exception as e:
    print("ERR", e)
PY

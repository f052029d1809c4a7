#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_forward.py -m gpu -q -x --timeout=600 -p no:cacheprovider -k "ln_fused or parity" 2>&1 | tail -3
for v in 1 0; do
  SMD_LNF=$v timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_ab2_sample_lnf$v.json 2>> gpurun_out/bench20.err
  SMD_LNF=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_ab2_train_lnf$v.json 2>> gpurun_out/bench20.err
done
python - <<'PY'
import json
for n in ["sample_lnf1", "sample_lnf0", "train_lnf1", "train_lnf0"]:
    try:
        d = json.loads(open("gpurun_out/r02_ab2_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
SMD_LNF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/r02_launches_sample_lnf1b.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_sample_lnf1b.csv 2>&1 | head -12

#!/bin/bash
# trimmed round-end refresh at HEAD (the full suite ran piecewise in run22 / run23 after the last kernel changes)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "not opt_in" 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke-ok')" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r02_bench_train.json 2> gpurun_out/final2.err
timeout 200 python bench.py --workload sample --no-extra > gpurun_out/r02_bench_sample.json 2>> gpurun_out/final2.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/r02_launches_train.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
python - <<'PY'
import json
for n in ["train", "sample"]:
    d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["gpu_launches"])
    for e in d.get("extra", []):
        print("   ", e["name"], round(e["ms_per_step"], 4), round(e["step_frac_of_sustained_peak"], 3))
PY

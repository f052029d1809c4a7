#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_sharding.py tests/test_gpu_forward.py tests/test_gpu_bench_shapes.py tests/test_gpu_strict.py tests/test_gpu_dropin.py -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -3
timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_emb.json 2>> gpurun_out/bench23.err
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_train_emb.json 2>> gpurun_out/bench23.err
python - <<'PY'
import json
for n in ["sample_emb", "train_emb"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(n, "ERR", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 45 --csv --log-file gpurun_out/r02_launches_sample_emb.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_sample_emb.csv 2>&1 | tail -5

#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_train.py tests/test_gpu_forward.py tests/test_gpu_bench_shapes.py tests/test_gpu_sampler.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/r02_launches_sample_v4.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
for v in 1 0; do
  SMD_LNF=$v timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_lnf$v.json 2>> gpurun_out/bench8.err
  SMD_LNF=$v timeout 200 python bench.py --steps 30 --warmup 8 --no-cpu --no-extra > gpurun_out/r02_bench_train_lnf$v.json 2>> gpurun_out/bench8.err
done
for v in 1 0; do
  SMD_LNF=$v timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_lnf${v}_b.json 2>> gpurun_out/bench8.err
done
tail -c 300 gpurun_out/bench8.err
python - <<'PY'
import json
for n in ["sample_lnf1", "sample_lnf0", "sample_lnf1_b", "sample_lnf0_b", "train_lnf1", "train_lnf0"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY

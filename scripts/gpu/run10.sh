#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ncsn.py tests/test_gpu_sharding.py tests/test_gpu_sampler.py tests/test_gpu_forward.py -m gpu -q --timeout=280 -p no:cacheprovider 2>&1 | tail -8
for cfg in "12 8" "12 12" "8 8"; do
  set -- $cfg
  tag="a$1_b$2"
  SMD_LNF_WARPS_A=$1 SMD_LNF_WARPS_B=$2 timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_$tag.json 2>> gpurun_out/bench10.err
  SMD_LNF_WARPS_A=$1 SMD_LNF_WARPS_B=$2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 45 --csv --log-file gpurun_out/r02_launches_sample_$tag.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
done
SMD_LNF=0 timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_lnf0.json 2>> gpurun_out/bench10.err
SMD_LNF=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 45 --csv --log-file gpurun_out/r02_launches_sample_lnf0.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
for cfg in "12 8" "12 12"; do
  set -- $cfg
  SMD_LNF_WARPS_A=$1 SMD_LNF_WARPS_B=$2 timeout 200 python bench.py --steps 30 --warmup 8 --no-cpu --no-extra > gpurun_out/r02_bench_train_a$1_b$2.json 2>> gpurun_out/bench10.err
done
tail -c 300 gpurun_out/bench10.err
python - <<'PY'
import json
for n in ["sample_a12_b8", "sample_a12_b12", "sample_a8_b8", "sample_lnf0", "train_a12_b8", "train_a12_b12"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY

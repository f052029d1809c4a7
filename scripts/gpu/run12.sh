#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sampler.py tests/test_gpu_bench_shapes.py tests/test_gpu_sharding.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -12
for v in 1 0; do
  SMD_ATTN_BLOCK=$v timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_attn$v.json 2>> gpurun_out/bench12.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 45 --csv --log-file gpurun_out/r02_launches_sample_attn1.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
tail -c 300 gpurun_out/bench12.err
python - <<'PY'
import json
for n in ["sample_attn1", "sample_attn0"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_forward.py::test_transformer_forward_parity[tiny-2]" "tests/test_gpu_forward.py::test_transformer_forward_parity[large_c42-2]" -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_memcheck_attn.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_memcheck_attn.log | tail -3

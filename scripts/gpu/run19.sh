#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sampler.py "tests/test_gpu_train.py::test_opt_in_trunk_paths" -m gpu -q -x --timeout=400 -p no:cacheprovider 2>&1 | tail -4
timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_ffn5.json 2>> gpurun_out/bench19.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_sample_ffn5.json").read().strip().splitlines()[-1])
print("sample", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"ffn_fused|attn_block" -s 12 -c 6 --csv --log-file gpurun_out/r02_ffn_sample.csv python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
awk -F'","' '{print substr($5,1,40), $NF}' gpurun_out/r02_ffn_sample.csv | tail -4
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:ffn_fused -s 6 -c 1 -o gpurun_out/r02_ffn_fused_full -f python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1

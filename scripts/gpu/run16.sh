#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sampler.py tests/test_gpu_bench_shapes.py tests/test_gpu_sharding.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -3
timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_ffn4.json 2>> gpurun_out/bench16.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_sample_ffn4.json").read().strip().splitlines()[-1])
print("sample", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"ffn_fused|attn_block" -s 12 -c 8 --csv --log-file gpurun_out/r02_ffn_sample.csv python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
awk -F'","' '{print substr($5,1,40), $NF}' gpurun_out/r02_ffn_sample.csv | tail -6
SMD_FFN_FUSED=2 SMD_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ffn_fused -c 4 --csv --log-file gpurun_out/r02_ffn_train_4096.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
awk -F'","' '{print substr($5,1,40), $NF}' gpurun_out/r02_ffn_train_4096.csv | tail -3
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:ffn_fused -s 6 -c 1 -o gpurun_out/r02_ffn_fused_full -f python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_forward.py::test_transformer_forward_parity[tiny-2]" tests/test_gpu_forward.py::test_fused_ffn_kernel -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_memcheck_ffn.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_memcheck_ffn.log | tail -3

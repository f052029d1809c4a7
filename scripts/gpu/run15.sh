#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_forward.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "ffn or parity" 2>&1 | tail -3
timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_ffn3.json 2>> gpurun_out/bench15.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_sample_ffn3.json").read().strip().splitlines()[-1])
print("sample", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ffn_fused -s 6 -c 6 --csv --log-file gpurun_out/r02_ffn_sample.csv python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
awk -F'","' '{print $5, $NF}' gpurun_out/r02_ffn_sample.csv | tail -4
SMD_FFN_FUSED=2 SMD_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ffn_fused -c 8 --csv --log-file gpurun_out/r02_ffn_train_4096.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
awk -F'","' '{print $5, $NF}' gpurun_out/r02_ffn_train_4096.csv | tail -4
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:ffn_fused -s 6 -c 1 -o gpurun_out/r02_ffn_fused_full -f python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1

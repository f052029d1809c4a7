#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_strict.py tests/test_gpu_forward.py tests/test_gpu_train.py -m gpu -q --timeout=280 -p no:cacheprovider 2>&1 | tail -4
for v in 1 0; do
  SMD_LNF=$v timeout 200 python bench.py --workload sample --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_sample_lnf$v.json 2>> gpurun_out/bench11.err
  SMD_LNF=$v timeout 200 python bench.py --steps 30 --warmup 8 --no-cpu --no-extra > gpurun_out/r02_bench_train_lnf$v.json 2>> gpurun_out/bench11.err
  SMD_LNF=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 45 --csv --log-file gpurun_out/r02_launches_sample_lnf$v.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
done
SMD_TRAIN_GRAPH=0 timeout 200 python bench.py --steps 30 --warmup 8 --no-cpu --no-extra > gpurun_out/r02_bench_train_nograph.json 2>> gpurun_out/bench11.err
tail -c 300 gpurun_out/bench11.err
python - <<'PY'
import json
for n in ["sample_lnf1", "sample_lnf0", "train_lnf1", "train_lnf0", "train_nograph"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
# memory checker on the smallest forward / sampler / train cases (slow: one case each)
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_forward.py::test_transformer_forward_parity[tiny-2]" "tests/test_gpu_sampler.py::test_reverse_step_supplied_noise[500]" "tests/test_gpu_train.py::test_gradients_match_autograd[tiny-2]" -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_memcheck.log | tail -4

#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_train.py tests/test_gpu_bench_shapes.py tests/test_gpu_forward.py tests/test_gpu_sampler.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -4
for v in 1 0; do
  SMD_FFN_SPLITK=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_train_splitk$v.json 2>> gpurun_out/bench18.err
done
SMD_ATTN_BLOCK_TRAIN=1 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_train_splitk1_attn1.json 2>> gpurun_out/bench18.err
python - <<'PY'
import json
for n in ["train_splitk1", "train_splitk0", "train_splitk1_attn1"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
SMD_TRAIN_GRAPH=0 timeout 200 python scripts/timeline.py train gpurun_out/r02_timeline_train_eager.json 2>&1 | tail -1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_train.py" -m gpu -q -p no:cacheprovider -x -k "grad or train_step" > gpurun_out/r02_memcheck_train.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_memcheck_train.log | tail -3

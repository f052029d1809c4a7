#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_bench_shapes.py tests/test_gpu_forward.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -4
for v in 1 0; do
  SMD_ATTN_BLOCK_TRAIN=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/r02_bench_train_attn$v.json 2>> gpurun_out/bench17.err
done
python - <<'PY'
import json
for n in ["train_attn1", "train_attn0"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
SMD_TRAIN_GRAPH=0 timeout 200 python scripts/timeline.py train gpurun_out/r02_timeline_train_eager.json 2>&1 | tail -1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_train.py" -m gpu -q -p no:cacheprovider -x -k "grad or train_step" > gpurun_out/r02_memcheck_train.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_memcheck_train.log | tail -3

#!/bin/bash
# Runs the GPU parity suites one process per file (a device fault in one cannot poison the next).
# Usage (under gpurun): bash scripts/gpu_tests.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in tests/test_gpu_gemm.py tests/test_gpu_forward.py tests/test_gpu_sampler.py "$@"; do
  n=$(basename $f .py)
  echo "=== $f"
  timeout 420 python -m pytest $f -m gpu -q -x --timeout=240 -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "exit=$?" >> gpurun_out/$n.log
  tail -n 25 gpurun_out/$n.log
done

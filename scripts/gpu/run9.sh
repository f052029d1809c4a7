#!/bin/bash
# 2-GPU box: NCSN-family parity tests, the pytest DP test, the DP worker log and the 2-GPU bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ncsn.py tests/test_gpu_strict.py tests/test_gpu_dp.py -m gpu -q --timeout=280 -p no:cacheprovider 2>&1 | tail -12
bash scripts/gpu/run_dp.sh 2

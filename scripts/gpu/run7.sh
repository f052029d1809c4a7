#!/bin/bash
# LN-fused epilogue v3 diagnosis: tests that failed, launch lists with / without the wait, full ncu capture
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_train.py tests/test_gpu_forward.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -15
for w in 0 1; do
  SMD_LNF_NOWAIT=$w timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/r02_launches_sample_nowait$w.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:4153 -s 4 -c 1 -o gpurun_out/r02_prof_lnf_a python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"
ls -la gpurun_out/*.ncu-rep

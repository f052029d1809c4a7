#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -s 6 -c 6 -o gpurun_out/r02_gemm_step_full -f python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/run21.log 2>&1
tail -3 gpurun_out/run21.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
SMD_TRAIN_GRAPH=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:ln_film_bwd_fast -s 10 -c 3 -o gpurun_out/r02_lnfilm_bwd_full -f python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/run24.log 2>&1
tail -2 gpurun_out/run24.log | cut -c1-200
ls -la gpurun_out/r02_lnfilm_bwd_full.ncu-rep

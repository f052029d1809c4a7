#!/bin/bash
# round-end validation on one B200: full GPU suite, smoke(), both bench arms, sample bench, launch lists, ncu --set full
# captures of the dominant GEMM and the two fused trunk kernels, CUPTI timelines.  Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke-ok')" 2>&1 | tail -2
echo "== bench (default)"; ( time timeout 900 python bench.py > gpurun_out/r02_bench_train.json 2> gpurun_out/final_bench.err ) 2>&1 | grep real
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/final_bench.err ) 2>&1 | grep real
echo "== bench sample"; timeout 400 python bench.py --workload sample --no-extra > gpurun_out/r02_bench_sample.json 2>> gpurun_out/final_bench.err
python - <<'PY'
import json
for n in ["train", "reference_arm", "sample"]:
    try:
        d = json.loads(open("gpurun_out/r02_bench_" + n + ".json").read().strip().splitlines()[-1])
        print(n, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "impl")}, (d.get("e2e") or {}).get("value"),
              (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
        for e in d.get("extra", []):
            print("   ", e["name"], round(e["ms_per_step"], 4), round(e["step_frac_of_sustained_peak"], 3))
    except Exception as e:
        print(n, "ERR", e)
PY
echo "== launch lists"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/r02_launches_train.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 45 --csv --log-file gpurun_out/r02_launches_sample.csv python bench.py --workload sample --steps 3 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_train.csv 2>&1 | head -30
python scripts/summarize_launches.py gpurun_out/r02_launches_sample.csv 2>&1 | head -14
echo "== ncu --set full"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/r02_dominant_gemm_4096 -f python scripts/dominant_gemm.py 4096 2 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/r02_dominant_gemm_32000 -f python scripts/dominant_gemm.py 32000 2 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:ffn_fused -s 6 -c 1 -o gpurun_out/r02_ffn_fused_full -f python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:attn_block -s 6 -c 1 -o gpurun_out/r02_attn_block_full -f python bench.py --workload sample --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
for f in r02_dominant_gemm_4096 r02_dominant_gemm_32000 r02_ffn_fused_full r02_attn_block_full; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv --metrics $M > gpurun_out/$f.csv 2>/dev/null
done
echo "== timelines"
timeout 200 python scripts/timeline.py train gpurun_out/r02_timeline_train.json 2>&1 | tail -1
SMD_TRAIN_GRAPH=0 timeout 200 python scripts/timeline.py train gpurun_out/r02_timeline_train_eager.json 2>&1 | tail -1
timeout 200 python scripts/timeline.py sample gpurun_out/r02_timeline_sample.json 2>&1 | tail -1
ls -la gpurun_out | tail -30

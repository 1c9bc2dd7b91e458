#!/bin/bash
# round-2d A/B on one box: warp-per-query attention + few-row GEMM against the round-2c kernels (env switches), short-chunk launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 60 python tools/ub_gemm_fewrows.py > gpurun_out/r2d_ub_fewrows_new.log 2>&1; cat gpurun_out/r2d_ub_fewrows_new.log
B="python bench.py --no-cpu-baseline --no-incumbent --sequence-frames 0"
timeout 150 $B --workload short > gpurun_out/r2d_bench_short_new2.json 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches_short.csv $B --workload short --steps 1 --warmup 1 > gpurun_out/r2d_ncu_short.log 2>&1
python tools/summarize_launches.py gpurun_out/r2d_launches_short.csv > gpurun_out/r2d_launches_short_summary.txt 2>&1; head -40 gpurun_out/r2d_launches_short_summary.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_bench_*3.json"))+sorted(glob.glob("gpurun_out/r2d_bench_short_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d.get("clocks",{}).get("sm_mhz"), {k:round(v["ms_per_step"],2) for k,v in d.get("kernel_classes",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY

#!/bin/bash
# End-of-session validation on one B200 box (run through gpurun): GPU suite, smoke, the driver's default bench line, the reference
# arm, the short-chunk and config-5 lines, and the ncu launch list of the bench command.  Outputs under gpurun_out/<tag>_*.
#   gpurun --timeout 1800 -- 'bash tools/validate_round.sh r2d'
tag=${1:-r}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/${tag}_gpu_suite.log 2>&1; tail -4 gpurun_out/${tag}_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
B="python bench.py --no-cpu-baseline --no-incumbent --sequence-frames 0"
timeout 200 $B --workload short > gpurun_out/${tag}_bench_short_chunk.json 2>/dev/null
timeout 300 $B --workload config5 > gpurun_out/${tag}_bench_config5.json 2>/dev/null
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${tag}_launches_all.csv $B --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_launch.log 2>&1
python - <<PY
import json
for f in ["${tag}_bench_n1","${tag}_bench_reference_arm","${tag}_bench_short_chunk","${tag}_bench_config5"]:
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1]); print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("clocks",{}).get("sm_mhz"), (d.get("roofline") or {}).get("frac"), {k:round(v["ms_per_step"],2) for k,v in d.get("kernel_classes",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY

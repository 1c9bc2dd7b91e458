#!/bin/bash
# round-2d A/B on one box: dispatch switches on the 5-frame-chunk workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-incumbent --sequence-frames 0 --workload short"
run() { name=$1; shift; env "$@" timeout 150 $B > gpurun_out/r2d_short_$name.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r2d_short_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), round(d["ms_per_step"],3), d.get("eager",{}).get("ms_per_step"))
PY
}
run base A=1
run narrow1.0 LSVS_GEMM_NARROW_PENALTY=1.0
run narrow2.0 LSVS_GEMM_NARROW_PENALTY=2.0
run nopersist LSVS_ATTN_PERSIST=0
run slices4 LSVS_GEMM_FEWROWS_SLICES=4
run base2 A=1

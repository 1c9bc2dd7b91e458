#!/bin/bash
# round-2d A/B on one box: narrow-tile penalty of the GEMM dispatcher at 4 / 5 / 8-frame chunks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 1.25 2.0 1.6 1.25 2.0; do echo "penalty $p"; LSVS_GEMM_NARROW_PENALTY=$p timeout 120 python tools/ub_short_chunk.py 2>/dev/null | cut -c1-120; done | tee gpurun_out/r2d_narrow_penalty.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dpt_gpu.py tests/test_gemm_gpu.py tests/test_precision_gpu.py -q -m gpu > gpurun_out/r2d_gpu_dpt.log 2>&1; tail -6 gpurun_out/r2d_gpu_dpt.log

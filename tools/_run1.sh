timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
LSVS_B200_LIB=variants/base/liblsvs_b200.so python tools/ub_gemm_epi4.py
python tools/ub_gemm_epi4.py
done 2>&1 | tee gpurun_out/r2c_ub_gemm_epi4.log

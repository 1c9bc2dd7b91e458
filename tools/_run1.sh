timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
LSVS_ATTN_POLY=0 python tools/ub_attn.py 2>&1 | grep "global S=32\|cfg5" | sed "s/^/p0.000 /"
LSVS_ATTN_POLY=0 LSVS_B200_LIB=variants/polyx/liblsvs_b200.so python tools/ub_attn.py 2>&1 | grep "global S=32\|cfg5" | sed "s/^/p0.125 /"
LSVS_ATTN_POLY=1 python tools/ub_attn.py 2>&1 | grep "global S=32\|cfg5" | sed "s/^/p0.250 /"
LSVS_ATTN_POLY=1 LSVS_B200_LIB=variants/polyx/liblsvs_b200.so python tools/ub_attn.py 2>&1 | grep "global S=32\|cfg5" | sed "s/^/p0.375 /"
done 2>&1 | cut -c1-230 | sed 's/"rel_l2_vs_sdpa"/err/; s/"B": 1, "H": 16, "hd": 64, //; s/"poly": "[01]", //' | tee gpurun_out/r2b_ab_poly_share.log
python tools/ub_attn.py 2>&1 | grep "frame S" | cut -c1-200

timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -m gpu > gpurun_out/r2b_test_attn_split.log 2>&1; tail -5 gpurun_out/r2b_test_attn_split.log
for rep in 1 2; do
python tools/ub_attn.py 2>&1 | grep "global\|frame S" | sed "s/^/split rep$rep /"
LSVS_ATTN_SPLIT_MAX=1 python tools/ub_attn.py 2>&1 | grep "global\|frame S" | sed "s/^/nosplit rep$rep /"
done > gpurun_out/r2b_ab_split.log 2>&1
LSVS_ATTN_SPLIT_MAX=3 python tools/ub_attn.py 2>&1 | grep "global" | sed "s/^/split3 /" >> gpurun_out/r2b_ab_split.log
LSVS_ATTN_SPLIT_MAX=6 python tools/ub_attn.py 2>&1 | grep "global" | sed "s/^/split6 /" >> gpurun_out/r2b_ab_split.log
cut -c1-200 gpurun_out/r2b_ab_split.log | sed 's/"poly": "default", "rel_l2_vs_sdpa"/err/; s/"B": 1, "H": 16, "hd": 64, //'

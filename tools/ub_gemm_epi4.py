"""qkv GEMM with the per-head LayerNorm + 2-D RoPE epilogue at M = 13 184: what the epilogue costs (ablations through the run-time
arguments: no RoPE, no LayerNorm columns = plain bias through the same kernel), back-to-back launches, 3 rotating buffer sets."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
M, N, K = 13184, 3072, 1024
def bench(fn_list, reps=30):
    for f in fn_list: f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn_list[i % len(fn_list)]()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
nw = torch.ones(64, device="cuda"); nb = torch.zeros(64, device="cuda")
tab = ops.rope_table(64, 16)
for name, kw0 in [("qkv + LayerNorm + RoPE", dict(n_q_cols=1024, n_k_cols=1024, rope_mode=ops.ROPE_2D, rope_tab=tab, tokens_per_frame=412, n_special=5, grid_w=37)),
                  ("qkv + LayerNorm (no RoPE)", dict(n_q_cols=1024, n_k_cols=1024, rope_mode=ops.ROPE_NONE)),
                  ("bias only through the same kernel", dict(n_q_cols=0, n_k_cols=0, rope_mode=ops.ROPE_NONE)),
                  ("plain qkv (EPI_BIAS_BF16)", None)]:
    fns = []
    for r in range(3):
        a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        if kw0 is None:
            fns.append(lambda a=a, w=w, bias=bias, out=out: ops.gemm(a, w, ops.EPI_BIAS_BF16, bias=bias, out=out))
        else:
            kw = dict(bias=bias, out=out, qn=(nw, nb), kn=(nw, nb), **kw0)
            fns.append(lambda a=a, w=w, kw=kw: ops.gemm(a, w, ops.EPI_HEADNORM64_BF16, **kw))
    ms = bench(fns)
    print(json.dumps({"gemm": name, "lib": os.environ.get("LSVS_B200_LIB", "product"), "us": round(ms * 1e3, 1), "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}))

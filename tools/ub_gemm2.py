"""GEMM microbenchmark, back-to-back launches (CPU launch overhead hidden), 3 rotating buffer sets (> L2)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
M = 13184
if os.environ.get("LSVS_GEMM_MODE"):
    from lsvs_b200 import native
    native.lib().lsvs_debug_gemm_mode(int(os.environ["LSVS_GEMM_MODE"]))
def bench(fn_list, reps=30):
    for f in fn_list: f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn_list[i % len(fn_list)]()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for (N, K, kind, name) in [(3072, 1024, ops.EPI_BIAS_BF16, "qkv"), (3072, 1024, ops.EPI_HEADNORM64_BF16, "qkv+norm+rope"),
                           (1024, 1024, ops.EPI_RESID_F32, "proj+resid"), (4096, 1024, ops.EPI_BIAS_GELU_BF16, "fc1+gelu"),
                           (1024, 4096, ops.EPI_RESID_F32, "fc2+resid"), (1024, 1024, ops.EPI_BIAS_BF16, "proj plain bf16")]:
    fns, fns_t = [], []
    for r in range(3):
        a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        bias = torch.randn(N, device="cuda"); gamma = torch.full((N,), 0.01, device="cuda")
        resid = torch.randn(M, N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(bias=bias)
        if kind == ops.EPI_RESID_F32: kw.update(gamma=gamma, resid=resid)
        else: kw.update(out=out)
        if kind == ops.EPI_HEADNORM64_BF16:
            nw = torch.ones(64, device="cuda"); nb = torch.zeros(64, device="cuda")
            kw.update(qn=(nw, nb), kn=(nw, nb), n_q_cols=1024, n_k_cols=1024, rope_mode=ops.ROPE_2D, rope_tab=ops.rope_table(64, 16),
                      tokens_per_frame=412, n_special=5, grid_w=37)
        fns.append(lambda a=a, w=w, kw=kw: ops.gemm(a, w, kind, **kw))
        fns_t.append(lambda a=a, w=w: torch.matmul(a, w.T))
    ms, ms_t = bench(fns), bench(fns_t)
    fl = 2.0 * M * N * K
    print(json.dumps({"gemm": name, "N": N, "K": K, "us": round(ms * 1e3, 1), "TFLOPs": round(fl / ms / 1e9, 1), "cublas_us": round(ms_t * 1e3, 1),
                      "cublas_TFLOPs": round(fl / ms_t / 1e9, 1)}))

"""Microbenchmark: tcgen05 GEMM shapes of one Aggregator block at S=32 (M=13184) vs torch.matmul (cuBLAS)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops

def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

M = int(sys.argv[1]) if len(sys.argv) > 1 else 13184
from lsvs_b200 import native
for mode in ((1, 0) if hasattr(native.lib(), 'lsvs_debug_gemm_mode') else (0,)):  # modes need a -DLSVS_MEASURE build
  if mode: native.lib().lsvs_debug_gemm_mode(mode)
  print("--- gemm mode", mode, "(1 = single-CTA 128x256, 0 = CTA pairs 256x256)")
  for (N, K, kind, name) in [(3072, 1024, ops.EPI_BIAS_BF16, "qkv"), (3072, 1024, ops.EPI_HEADNORM64_BF16, "qkv+norm+rope"),
                             (1024, 1024, ops.EPI_RESID_F32, "proj+resid"), (4096, 1024, ops.EPI_BIAS_GELU_BF16, "fc1+gelu"),
                             (1024, 4096, ops.EPI_RESID_F32, "fc2+resid")]:
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
      ms = timeit(lambda: ops.gemm(a, w, kind, **kw))
      ms_t = timeit(lambda: torch.matmul(a, w.T))
      fl = 2.0 * M * N * K
      print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "TFLOPs": round(fl / ms / 1e9, 1),
                        "cublas_ms": round(ms_t, 4), "cublas_TFLOPs": round(fl / ms_t / 1e9, 1)}))

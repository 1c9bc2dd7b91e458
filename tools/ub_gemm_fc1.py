"""fc1 GEMM (M = 13 184, N = 4096, K = 1024): GELU epilogue vs plain bias through the same kernel family; back-to-back launches."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
M, N, K = 13184, 4096, 1024
def bench(fn_list, reps=30):
    for f in fn_list: f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn_list[i % len(fn_list)]()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for name, kind in [("fc1 + GELU", ops.EPI_BIAS_GELU_BF16), ("fc1 plain bias", ops.EPI_BIAS_BF16)]:
    fns = []
    for r in range(3):
        a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fns.append(lambda a=a, w=w, bias=bias, out=out, kind=kind: ops.gemm(a, w, kind, bias=bias, out=out))
    ms = bench(fns)
    print(json.dumps({"gemm": name, "lib": os.environ.get("LSVS_B200_LIB", "product"), "us": round(ms * 1e3, 1), "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}))

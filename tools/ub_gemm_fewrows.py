"""Few-row GEMMs of the camera-head trunk (M = 32 frames, D = 2048) and of the camera head's vector path: weight streaming.
Six weight sets per shape are cycled (> 126 MB of L2 for the large shapes), launches back to back on one stream; reports us per launch
and the weight bytes per second against the measured HBM copy peak.  LSVS_GEMM_FEWROWS=0 selects the round-2 128 x 64-tile kernel.

    python tools/ub_gemm_fewrows.py            # gemm_fewrows_tcgen05 (product default)
    LSVS_GEMM_FEWROWS=0 python tools/ub_gemm_fewrows.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops

M = int(os.environ.get("UB_M", 32))
SHAPES = [("trunk qkv", 6144, 2048, ops.EPI_BIAS_BF16), ("trunk proj+resid", 2048, 2048, ops.EPI_RESID_F32),
          ("trunk fc1+gelu", 8192, 2048, ops.EPI_BIAS_GELU_BF16), ("trunk fc2+resid", 2048, 8192, ops.EPI_RESID_F32),
          ("adaLN modulation (split weights)", 6144, 6144, ops.EPI_BIAS_F32), ("pose fc1 (split weights)", 1024, 6144, ops.EPI_BIAS_F32)]
peak = 6546.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def bench(fns, reps=10):
    """the launches go into one CUDA graph (6 weight sets x 4), replayed: the Python / ctypes launch cost (~15 us) is not in the number"""
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns: f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(4):
                for f in fns: f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * 4 * len(fns))


total = 0.0
for name, N, K, kind in SHAPES:
    fns = []
    for r in range(6):
        a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        bias = torch.randn(N, device="cuda")
        if kind == ops.EPI_RESID_F32:
            resid = torch.randn(M, N, device="cuda"); g = torch.full((N,), 0.01, device="cuda")
            fns.append(lambda a=a, w=w, bias=bias, resid=resid, g=g: ops.gemm(a, w, kind, bias=bias, gamma=g, resid=resid))
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.float32 if kind == ops.EPI_BIAS_F32 else torch.bfloat16)
            fns.append(lambda a=a, w=w, bias=bias, out=out, kind=kind: ops.gemm(a, w, kind, bias=bias, out=out))
    ms = bench(fns)
    total += ms
    gbps = N * K * 2 / ms / 1e6
    print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "fewrows_kernel": os.environ.get("LSVS_GEMM_FEWROWS", "1") != "0",
                      "us": round(ms * 1e3, 2), "weight_GBps": round(gbps, 0), "frac_of_hbm_peak": round(gbps / peak, 3)}))
print(json.dumps({"sum_us": round(total * 1e3, 1), "hbm_peak_GBps": peak}))

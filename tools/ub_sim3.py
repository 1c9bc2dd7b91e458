"""Microbenchmark: Sim(3) apply over point maps, achieved HBM GB/s (24 B/point) at BASELINE sizes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from aligned_vggt.utils import alignment as A

def timeit(fn, iters=20, warm=5):
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

for (S, H, W) in [(32, 154, 518), (64, 518, 518)]:
    pts = torch.randn(1, S, H, W, 3, device="cuda")
    T = torch.eye(4, device="cuda")[None]; s = torch.full((1,), 1.3, device="cuda")
    ms = timeit(lambda: A.apply_sim3_alignment_on_point_maps(pts, T, s))
    n = S * H * W
    print(json.dumps({"kernel": "sim3_points", "points": n, "ms": ms, "GBps": 24 * n / ms / 1e6}))

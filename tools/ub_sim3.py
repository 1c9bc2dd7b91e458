"""Microbenchmark: Sim(3) apply over point maps through the C ABI, achieved HBM GB/s (24 B/point) at BASELINE sizes.
Back-to-back launches on rotating buffers (> L2) so that neither Python overhead nor cache hits enter the number."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import native

lib = native.lib()
for (S, H, W) in [(32, 154, 518), (64, 518, 518)]:
    n = S * H * W
    nbuf = max(3, int(400e6 // (n * 12)) + 1)
    pts = [torch.randn(n, 3, device="cuda") for _ in range(nbuf)]
    outs = [torch.empty(n, 3, device="cuda") for _ in range(nbuf)]
    T = torch.eye(4, device="cuda")[None].contiguous(); s = torch.full((1,), 1.3, device="cuda")
    def call(i):
        native.check(lib.lsvs_sim3_apply_points(native.ptr(pts[i % nbuf]), native.ptr(T), native.ptr(s), native.ptr(outs[i % nbuf]), ctypes.c_int(1),
                                                ctypes.c_longlong(n), native.stream_ptr()), "sim3")
    for i in range(5): call(i)
    torch.cuda.synchronize()
    reps = 60
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): call(i)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"kernel": "sim3_points", "points": n, "us": round(ms * 1e3, 2), "GBps": round(24 * n / ms / 1e6, 1)}))

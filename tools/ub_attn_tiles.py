"""Probe: per-key-block cycle cost of the attention kernel with one vs two active query tiles per CTA (Lq = 128 / 256)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

B, H, hd = 37, 16, 64
D = H * hd
res = {}
for Lk in (3328, 13184):
    kv = torch.randn(B * Lk, 2 * D, device="cuda").bfloat16()
    for Lq in (128, 256):
        q = torch.randn(B * Lq, D, device="cuda").bfloat16()
        out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(q, kv[:, :D], kv[:, D:], B, H, hd, Lq, Lk, out=out))
        res[(Lq, Lk)] = ms
        print(json.dumps({"Lq": Lq, "Lk": Lk, "ms": round(ms, 4)}))
waves = -(-B * H // 148)
for Lq in (128, 256):
    d_ms = res[(Lq, 13184)] - res[(Lq, 3328)]
    d_blocks = (13184 - 3328) // 128
    print(json.dumps({"Lq": Lq, "cycles_per_kv_block@1.9GHz (slope)": round(d_ms * 1e-3 * 1.9e9 / waves / d_blocks, 1)}))

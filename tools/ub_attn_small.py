"""Short-sequence attention shapes of the alignment / camera heads, 50 back-to-back launches between two events (a single
launch is shorter than the host-side launch path).  LSVS_ATTN_WARP=0 routes them to the tcgen05 kernel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
for (name, B, H, hd, Lq, Lk) in [("head temporal 32x9", 413, 8, 128, 32, 9), ("head temporal first chunk 32x32", 413, 8, 128, 32, 32),
                                 ("camera trunk 32x32", 1, 16, 128, 32, 32), ("short chunk temporal 5x2", 413, 8, 128, 5, 2),
                                 ("config 5 temporal 64x17", 1370, 8, 128, 64, 17)]:
    D = H * hd
    q = torch.randn(B * Lq, D, device="cuda").bfloat16(); kv = torch.randn(B * Lk, 2 * D, device="cuda").bfloat16()
    out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.attention(q, kv[:, :D], kv[:, D:], B, H, hd, Lq, Lk, out=out)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): fn()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1000 / 50
    mb = (q.numel() + kv.numel() + out.numel()) * 2 / 1e6
    print(json.dumps({"attn": name, "warp_path": os.environ.get("LSVS_ATTN_WARP", "1"), "us": round(us, 2), "MB": round(mb, 1), "GBps": round(mb / us * 1e3, 0)}))

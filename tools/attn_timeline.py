"""Debug probe (build with -DLSVS_ATTN_PHASES): globaltimer timeline of the first items of CTA 0's softmax warp 0 in the
persistent 412-token kernel: item start, after each key block, rows stored (ns relative to the first item's start)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops, native
lib = native.lib()
B, H, hd, L = 32, 16, 64, 412
D = H * hd
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
out = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, hd, L, L, out=out)
buf = (ctypes.c_ulonglong * 64)()
assert lib.lsvs_debug_attn_timeline(buf) == 0
t0 = buf[0]
for n in range(7):
    row = [int(buf[n * 8 + k]) - int(t0) if buf[n * 8 + k] else None for k in range(8)]
    print(json.dumps({"item": n, "start": row[0], "after blocks": row[1:5], "rows stored": row[7]}))

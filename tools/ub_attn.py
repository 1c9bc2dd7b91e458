"""Microbenchmark: tcgen05 attention at BASELINE shapes vs torch SDPA (library flash kernel)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops

def timeit(fn, iters=5, warm=2):
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

for (name, B, H, hd, L) in [("global S=32", 1, 16, 64, 13184), ("frame S=32", 32, 16, 64, 412), ("global S=4", 1, 16, 64, 1648),
                            ("head frame", 32, 8, 128, 413), ("global cfg5 1/4", 1, 16, 64, 21984)]:
    D = H * hd
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
    out = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, hd, L, L, out=out))
    q4 = qkv.view(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2]))
    fl = 4.0 * B * H * L * L * hd
    ref = torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2]).transpose(1, 2).reshape(B * L, D).float()
    err = float((out.float() - ref).norm() / ref.norm())
    print(json.dumps({"attn": name, "poly": os.environ.get("LSVS_ATTN_POLY", "default"), "rel_l2_vs_sdpa": round(err, 5), "B": B, "H": H, "hd": hd, "L": L, "ms": round(ms, 4), "TFLOPs": round(fl / ms / 1e9, 1),
                      "sdpa_ms": round(ms_t, 4), "sdpa_TFLOPs": round(fl / ms_t / 1e9, 1)}))

# temporal cross attention of the alignment head: 413 groups x 8 heads, 32 queries x 9 keys, head dim 128
B, H, hd, Lq, Lk = 413, 8, 128, 32, 9
D = H * hd
q = torch.randn(B * Lq, D, device="cuda").bfloat16(); kv = torch.randn(B * Lk, 2 * D, device="cuda").bfloat16()
out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: ops.attention(q, kv[:, :D], kv[:, D:], B, H, hd, Lq, Lk, out=out))
sh = lambda t, L: t.reshape(B, L, H, hd).transpose(1, 2).float()
ref = torch.nn.functional.scaled_dot_product_attention(sh(q, Lq), sh(kv[:, :D], Lk), sh(kv[:, D:], Lk)).transpose(1, 2).reshape(B * Lq, D)
print(json.dumps({"attn": "head temporal 32x9", "ms": round(ms, 4), "rel_l2_vs_fp32": round(float((out.float() - ref).norm() / ref.norm()), 5),
                  "bytes_MB": round((q.numel() + kv.numel() + out.numel()) * 2 / 1e6, 1)}))

"""Debug probe: per-phase cycle breakdown of the attention softmax warps (needs a build with -DLSVS_ATTN_PHASES:
   LSVS_NVCC_DEFINES=-DLSVS_ATTN_PHASES python large-scale-vit-slam_b200/lsvs_b200/build.py)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops, native

lib = native.lib()
names = ["wait s_full", "tmem ld", "arrive+mask+wait pv(i-2)", "max+exp2+P store", "fence+arrive p_ready"]
for (B, H, hd, Lq, Lk) in [(1, 16, 64, 13184, 13184), (37, 16, 64, 128, 13184), (32, 16, 64, 412, 412)]:
    D = H * hd
    q = torch.randn(B * Lq, D, device="cuda").bfloat16()
    kv = torch.randn(B * Lk, 2 * D, device="cuda").bfloat16()
    out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.attention(q, kv[:, :D], kv[:, D:], B, H, hd, Lq, Lk, out=out)
    buf = (ctypes.c_ulonglong * 64)()
    assert lib.lsvs_debug_attn_phases(buf) == 0
    n_kv = -(-Lk // 128)
    print(json.dumps({"shape": [B, H, Lq, Lk], "milestones_ns (setup, first S, loop done, epilogue, sync, dealloc)": [int(buf[48 + k]) for k in range(6)]}))
    for w in (0, 4):
        ph = [buf[w * 8 + k] / n_kv for k in range(5)]
        print(json.dumps({"shape": [B, H, Lq, Lk], "warp": w, "cycles/iter": round(sum(ph), 1),
                          **{names[k]: round(ph[k], 1) for k in range(5)}}))

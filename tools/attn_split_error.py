"""Error of the attention output against fp32 math for a head whose units are split into key ranges (head 15) and one whose
units are not (head 0), 1 x 13 184 tokens x 16 heads; LSVS_ATTN_SPLIT_MAX=1 turns the split off."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
H, hd, L = 16, 64, 13184
D = H * hd
g = torch.Generator("cuda").manual_seed(0)
qkv = torch.randn(L, 3 * D, device="cuda", generator=g).bfloat16()
out = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 1, H, hd, L, L).float()
q4 = qkv.view(1, L, 3, H, hd).permute(2, 0, 3, 1, 4)
sd = torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2]).transpose(1, 2).reshape(L, D).float()
res = {"split_max": os.environ.get("LSVS_ATTN_SPLIT_MAX", "default")}
for h in (0, 15):
    q, k, v = (qkv[:, i * D + h * hd:i * D + (h + 1) * hd].double() for i in range(3))
    ref = torch.zeros(L, hd, device="cuda", dtype=torch.float64)
    for r0 in range(0, L, 2048):
        p = torch.softmax(q[r0:r0 + 2048] @ k.T / 8.0, dim=-1)
        ref[r0:r0 + 2048] = p @ v
    e = lambda a: float((a[:, h * hd:(h + 1) * hd].double() - ref).norm() / ref.norm())
    res[f"head{h}"] = {"ours_vs_fp64": round(e(out), 6), "sdpa_vs_fp64": round(e(sd), 6)}
print(json.dumps(res))

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
B, H, hd, L = (int(x) for x in sys.argv[1:5])
D = H * hd
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
out = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, hd, L, L)
torch.cuda.synchronize()
q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(B, L, H, hd).transpose(1, 2) for i in range(3))
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * L, D)
print("ok", B, H, hd, L, float((out.float() - ref).norm() / ref.norm()))

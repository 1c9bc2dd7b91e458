"""Tiny driver for ncu --set full captures: one global-attention launch and the GEMM epilogue variants at S=32 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "all"
M, D = 13184, 1024
if which in ("all", "attn"):
    qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
    out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(int(os.environ.get("PROF_ITERS", 3))):
        ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 1, 16, 64, M, M, out=out)
if which in ("all", "gemm"):
    a = torch.randn(M, D, device="cuda").bfloat16(); w = (torch.randn(3 * D, D, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(3 * D, device="cuda"); o = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
    nw = torch.ones(64, device="cuda"); nb = torch.zeros(64, device="cuda")
    for _ in range(int(os.environ.get("PROF_ITERS", 3))):
        ops.gemm(a, w, ops.EPI_HEADNORM64_BF16, bias=bias, out=o, qn=(nw, nb), kn=(nw, nb), n_q_cols=D, n_k_cols=D, rope_mode=ops.ROPE_2D,
                 rope_tab=ops.rope_table(64, 16), tokens_per_frame=412, n_special=5, grid_w=37)
    w2 = (torch.randn(D, D, device="cuda") * 0.03).bfloat16(); resid = torch.randn(M, D, device="cuda"); g = torch.full((D,), 0.01, device="cuda")
    for _ in range(int(os.environ.get("PROF_ITERS", 3))):
        ops.gemm(a, w2, ops.EPI_RESID_F32, bias=bias[:D].contiguous(), gamma=g, resid=resid)
    w3 = (torch.randn(4 * D, D, device="cuda") * 0.03).bfloat16(); b3 = torch.randn(4 * D, device="cuda"); o3 = torch.empty(M, 4 * D, device="cuda", dtype=torch.bfloat16)
    for _ in range(int(os.environ.get("PROF_ITERS", 3))):
        ops.gemm(a, w3, ops.EPI_BIAS_GELU_BF16, bias=b3, out=o3)
if which in ("all", "frame"):
    L = 412
    qkv = torch.randn(32 * L, 3 * D, device="cuda").bfloat16()
    out = torch.empty(32 * L, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(int(os.environ.get("PROF_ITERS", 3))):
        ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 32, 16, 64, L, L, out=out)
torch.cuda.synchronize()
print("ok")

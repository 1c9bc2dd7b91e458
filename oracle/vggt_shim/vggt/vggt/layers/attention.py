import torch.nn as nn
from oracle import functional as OF
from .._p import params_of


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True, proj_bias=True, attn_drop=0.0, proj_drop=0.0,
                 norm_layer=nn.LayerNorm, qk_norm=False, fused_attn=True, rope=None):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.fused_attn = fused_attn
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)
        self.rope = rope

    def forward(self, x, pos=None):
        base = self.rope.base_frequency if self.rope is not None else None
        return OF.attention(params_of(self), "", x, self.num_heads, pos, base)


class MemEffAttention(Attention):
    """Without xformers upstream falls back to plain Attention.forward(x) (no RoPE)."""

    def forward(self, x, attn_bias=None, pos=None):
        assert attn_bias is None
        return super().forward(x, pos=None if self.rope is None else pos)

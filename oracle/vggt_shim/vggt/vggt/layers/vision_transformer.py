"""DINOv2 ViT with register tokens (A.2) as a parameter container; math in oracle.functional."""
from functools import partial

import torch
import torch.nn as nn
from oracle import functional as OF
from .._p import params_of
from .attention import MemEffAttention
from .block import Block
from .patch_embed import PatchEmbed


class DinoVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, ffn_bias=True, proj_bias=True, init_values=None, num_register_tokens=0,
                 interpolate_antialias=False, interpolate_offset=0.1, block_chunks=0, qk_norm=False, **_unused):
        super().__init__()
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.embed_dim = embed_dim
        self.depth = depth
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.num_register_tokens = num_register_tokens
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        n_patches = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patches + 1, embed_dim))
        self.register_tokens = nn.Parameter(torch.zeros(1, num_register_tokens, embed_dim)) if num_register_tokens else None
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, proj_bias=proj_bias,
                  ffn_bias=ffn_bias, norm_layer=norm_layer, init_values=init_values, attn_class=MemEffAttention,
                  qk_norm=qk_norm) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.init_weights()

    def init_weights(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        if self.register_tokens is not None:
            nn.init.normal_(self.register_tokens, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x, is_training=True):
        tok = OF.dinov2_patch_tokens(params_of(self), "", x, self.depth, self.num_heads, self.patch_size,
                                     self.num_register_tokens)
        return {"x_norm_patchtokens": tok}


def vit_large(patch_size=16, num_register_tokens=0, depth=24, **kwargs):
    return DinoVisionTransformer(patch_size=patch_size, embed_dim=1024, depth=depth, num_heads=16, mlp_ratio=4,
                                 num_register_tokens=num_register_tokens, **kwargs)

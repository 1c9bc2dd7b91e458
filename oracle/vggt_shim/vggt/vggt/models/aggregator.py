"""Aggregator (A.1) as a parameter container; math in oracle.functional.aggregator_forward."""
import os

import torch
import torch.nn as nn
from oracle import functional as OF
from .._p import params_of
from ..layers.block import Block
from ..layers.rope import RotaryPositionEmbedding2D, PositionGetter
from ..layers.vision_transformer import vit_large


class Aggregator(nn.Module):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                 num_register_tokens=4, block_fn=Block, qkv_bias=True, proj_bias=True, ffn_bias=True,
                 patch_embed="dinov2_vitl14_reg", aa_order=("frame", "global"), aa_block_size=1, qk_norm=True,
                 rope_freq=100, init_values=0.01, patch_embed_depth=24):
        super().__init__()
        # test hook: the reference builds Aggregator(img_size, patch_size, embed_dim) with no depth argument;
        # VGGT_SHIM_DEPTH="<aa_depth>,<dino_depth>" shrinks the stack for fast CPU fixtures.
        if os.environ.get("VGGT_SHIM_DEPTH"):
            depth, patch_embed_depth = (int(v) for v in os.environ["VGGT_SHIM_DEPTH"].split(","))
        assert patch_embed == "dinov2_vitl14_reg" and embed_dim == 1024 and aa_block_size == 1
        self.patch_embed = vit_large(img_size=img_size, patch_size=patch_size, num_register_tokens=num_register_tokens,
                                     interpolate_antialias=True, interpolate_offset=0.0, block_chunks=0,
                                     init_values=1.0, depth=patch_embed_depth)
        self.patch_embed.mask_token.requires_grad_(False)
        self.rope = RotaryPositionEmbedding2D(frequency=rope_freq) if rope_freq > 0 else None
        self.position_getter = PositionGetter() if self.rope is not None else None
        mk = lambda: block_fn(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                              proj_bias=proj_bias, ffn_bias=ffn_bias, init_values=init_values, qk_norm=qk_norm,
                              rope=self.rope)
        self.frame_blocks = nn.ModuleList([mk() for _ in range(depth)])
        self.global_blocks = nn.ModuleList([mk() for _ in range(depth)])
        self.depth = depth
        self.dino_depth = patch_embed_depth
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.num_register_tokens = num_register_tokens
        self.patch_start_idx = 1 + num_register_tokens
        self.rope_freq = rope_freq
        self.camera_token = nn.Parameter(torch.randn(1, 2, 1, embed_dim))
        self.register_token = nn.Parameter(torch.randn(1, 2, num_register_tokens, embed_dim))
        nn.init.normal_(self.camera_token, std=1e-6)
        nn.init.normal_(self.register_token, std=1e-6)
        for name, value in (("_resnet_mean", OF.RESNET_MEAN), ("_resnet_std", OF.RESNET_STD)):
            self.register_buffer(name, torch.FloatTensor(value).view(1, 1, 3, 1, 1), persistent=False)

    def forward(self, images):
        return OF.aggregator_forward(params_of(self), "", images, self.depth, self.dino_depth, self.num_heads,
                                     self.patch_size, self.num_register_tokens, float(self.rope_freq))

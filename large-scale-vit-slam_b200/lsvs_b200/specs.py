"""state_dict specifications (name -> shape) of the modules on the hot path, plus default initialisation.

The key names are the reference's checkpoint contract: `aggregator.*`, `camera_head.*` as in
facebook/VGGT-1B, `alignment_head.*` as created by aligned_vggt/heads/alignment_head.py:100-221.
tests/golden/state_dict_spec.json (generated from the reference classes) pins these lists.
"""
import math
from typing import List, Tuple

import torch
import torch.nn as nn

Spec = List[Tuple[str, Tuple[int, ...]]]


def _lin(n, o, i):
    return [(n + ".weight", (o, i)), (n + ".bias", (o,))]


def _norm(n, d):
    return [(n + ".weight", (d,)), (n + ".bias", (d,))]


def block_spec(pre: str, dim: int, heads: int, qk_norm: bool, layer_scale: bool = True) -> Spec:
    hd = dim // heads
    s = _norm(pre + "norm1", dim) + _lin(pre + "attn.qkv", 3 * dim, dim)
    if qk_norm:
        s += _norm(pre + "attn.q_norm", hd) + _norm(pre + "attn.k_norm", hd)
    s += _lin(pre + "attn.proj", dim, dim)
    if layer_scale:
        s += [(pre + "ls1.gamma", (dim,))]
    s += _norm(pre + "norm2", dim) + _lin(pre + "mlp.fc1", 4 * dim, dim) + _lin(pre + "mlp.fc2", dim, 4 * dim)
    if layer_scale:
        s += [(pre + "ls2.gamma", (dim,))]
    return s


def cross_block_spec(pre: str, dim: int, heads: int) -> Spec:
    hd = dim // heads
    s = _norm(pre + "norm1", dim) + _lin(pre + "attn.q", dim, dim) + _lin(pre + "attn.k", dim, dim) + _lin(pre + "attn.v", dim, dim)
    s += _norm(pre + "attn.q_norm", hd) + _norm(pre + "attn.k_norm", hd) + _lin(pre + "attn.proj", dim, dim)
    s += [(pre + "ls1.gamma", (dim,))] + _norm(pre + "norm2", dim) + _lin(pre + "mlp.fc1", 4 * dim, dim)
    s += _lin(pre + "mlp.fc2", dim, 4 * dim) + [(pre + "ls2.gamma", (dim,))] + _norm(pre + "norm3", dim)
    return s


def aggregator_spec(depth: int = 24, dino_depth: int = 24, img_size: int = 518, patch: int = 14, dim: int = 1024) -> Spec:
    n_pos = (img_size // patch) ** 2 + 1
    s = [("camera_token", (1, 2, 1, dim)), ("register_token", (1, 2, 4, dim)),
         ("patch_embed.cls_token", (1, 1, dim)), ("patch_embed.pos_embed", (1, n_pos, dim)),
         ("patch_embed.register_tokens", (1, 4, dim)), ("patch_embed.mask_token", (1, dim)),
         ("patch_embed.patch_embed.proj.weight", (dim, 3, patch, patch)), ("patch_embed.patch_embed.proj.bias", (dim,))]
    for i in range(dino_depth):
        s += block_spec(f"patch_embed.blocks.{i}.", dim, 16, qk_norm=False)
    s += _norm("patch_embed.norm", dim)
    for grp in ("frame_blocks", "global_blocks"):
        for i in range(depth):
            s += block_spec(f"{grp}.{i}.", dim, 16, qk_norm=True)
    return s


def camera_head_spec(dim_in: int = 2048, trunk_depth: int = 4) -> Spec:
    s = []
    for i in range(trunk_depth):
        s += block_spec(f"trunk.{i}.", dim_in, 16, qk_norm=False)
    s += _norm("token_norm", dim_in) + _norm("trunk_norm", dim_in) + [("empty_pose_tokens", (1, 1, 9))]
    s += _lin("embed_pose", dim_in, 9) + _lin("poseLN_modulation.1", 3 * dim_in, dim_in)
    s += _lin("pose_branch.fc1", dim_in // 2, dim_in) + _lin("pose_branch.fc2", 9, dim_in // 2)
    return s


def alignment_head_spec(in_dim: int = 2048, dim: int = 1024, dec: int = 512, depth_aa: int = 4, depth_dec: int = 2,
                        heads: int = 8, n_mem: int = 8) -> Spec:
    s = [("per_frame_alignment_token", (1, 2, 1, dim))]
    if n_mem > 0:
        s += [("memory_token", (1, n_mem, dec)), ("alpha", ())]
    s += _lin("project_in", dim, in_dim) + _lin("project_dec", dec, dim)
    for i in range(depth_aa):
        s += block_spec(f"frame_blocks.{i}.", dim, heads, qk_norm=True)
    for i in range(depth_aa):
        s += cross_block_spec(f"temporal_blocks.{i}.", dim, heads)
    for grp in ("chunk_cross_blocks", "frame_cross_blocks"):
        for i in range(depth_dec):
            s += cross_block_spec(f"{grp}.{i}.", dec, heads)
    s += _lin("chunk_sim3_decoder.fc1", dec // 2, dec) + _lin("chunk_sim3_decoder.fc2", 8, dec // 2)
    s += _lin("frame_se3_decoder.fc1", dec // 2, dec) + _lin("frame_se3_decoder.fc2", 7, dec // 2)
    s += _norm("token_norm", dim) + _norm("dec_norm", dec) + _norm("chunk_norm", dec) + _norm("frame_norm", dec)
    if n_mem > 0:
        s += _lin("frame_proj", n_mem * dec, dec)
        for i in range(n_mem):
            s += _lin(f"gated_update.delta_mlps.{i}.0", dec, 3 * dec) + _lin(f"gated_update.delta_mlps.{i}.2", dec, dec)
        s += _lin("gated_update.gate_mlp.0", dec, 2 * dec) + _lin("gated_update.gate_mlp.2", 1, dec)
    return s


DPT_OUT_CHANNELS = (256, 512, 1024, 1024)


def dpt_head_spec(dim_in: int = 2048, output_dim: int = 4, features: int = 256, out_channels=DPT_OUT_CHANNELS) -> Spec:
    """UPSTREAM vggt DPTHead (facebook/VGGT-1B keys `depth_head.*` / `point_head.*`); conv weights in torch layout."""
    def conv(n, o, i, k, bias=True):
        return [(n + ".weight", (o, i, k, k))] + ([(n + ".bias", (o,))] if bias else [])
    s = _norm("norm", dim_in)
    for i, c in enumerate(out_channels):
        s += conv(f"projects.{i}", c, dim_in, 1)
    s += [("resize_layers.0.weight", (out_channels[0], out_channels[0], 4, 4)), ("resize_layers.0.bias", (out_channels[0],)),
          ("resize_layers.1.weight", (out_channels[1], out_channels[1], 2, 2)), ("resize_layers.1.bias", (out_channels[1],))]
    s += conv("resize_layers.3", out_channels[3], out_channels[3], 3)
    for i, c in enumerate(out_channels):
        s += conv(f"scratch.layer{i + 1}_rn", features, c, 3, bias=False)
    for r in (1, 2, 3, 4):
        s += conv(f"scratch.refinenet{r}.out_conv", features, features, 1)
        for u in ((1, 2) if r != 4 else (2,)):  # refinenet4 has no residual input (has_residual=False)
            s += conv(f"scratch.refinenet{r}.resConfUnit{u}.conv1", features, features, 3)
            s += conv(f"scratch.refinenet{r}.resConfUnit{u}.conv2", features, features, 3)
    s += conv("scratch.output_conv1", features // 2, features, 3)
    s += conv("scratch.output_conv2.0", 32, features // 2, 3) + conv("scratch.output_conv2.2", output_dim, 32, 1)
    return s


class ParamTree(nn.Module):
    """Nested parameter container whose state_dict keys are exactly the dotted names of a spec."""

    def __init__(self, spec: Spec = ()):
        super().__init__()
        for name, shape in spec:
            parts = name.split(".")
            m = self
            for p in parts[:-1]:
                if p not in m._modules:
                    m.add_module(p, ParamTree())
                m = m._modules[p]
            m.register_parameter(parts[-1], nn.Parameter(torch.empty(shape)))


@torch.no_grad()
def init_default_(root: nn.Module, seed: int = None) -> None:
    """Random init in the spirit of the reference modules' constructors: LayerNorm 1/0, LayerScale 0.01
    (1.0 in the DINOv2 ViT), special tokens std 1e-6, Linear kaiming-uniform (trunc-normal 0.02 in DINOv2),
    orthonormal memory tokens, alpha 0.1, gate bias 0 / weight std 0.1 (alignment_head.py:208-218,
    gated_update.py:38-40)."""
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    all_params = dict(root.named_parameters())
    for name, p in all_params.items():
        leaf = name.rsplit(".", 1)[-1]
        dino = "patch_embed." in name
        if leaf == "gamma":
            p.fill_(1.0 if "patch_embed.blocks." in name else 0.01)
        elif leaf == "memory_token":
            a = torch.empty(p.shape[-1], p.shape[-2]).normal_(generator=gen)
            q, _ = torch.linalg.qr(a)
            p.copy_(q.T.reshape(p.shape))
        elif leaf == "alpha":
            p.fill_(0.1)
        elif leaf in ("mask_token", "empty_pose_tokens"):
            p.zero_()
        elif leaf == "pos_embed":
            p.normal_(0, 0.02, generator=gen).clamp_(-0.04, 0.04)
        elif leaf in ("camera_token", "register_token", "cls_token", "register_tokens", "per_frame_alignment_token"):
            p.normal_(0, 1e-6, generator=gen)
        elif leaf == "weight" and p.dim() == 1:
            p.fill_(1.0)
        elif leaf == "weight":
            if "gate_mlp.2" in name:
                p.normal_(0, 0.1, generator=gen)
            elif dino:
                p.normal_(0, 0.02, generator=gen).clamp_(-0.04, 0.04)
            else:
                fan_in = p[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                p.uniform_(-bound, bound, generator=gen)
        elif leaf == "bias":
            if dino or "gate_mlp.2" in name or "norm" in name.rsplit(".", 2)[-2]:
                p.zero_()
            else:
                fan_in = None
                wname = name[:-4] + "weight"
                w = all_params.get(wname)
                fan_in = w[0].numel() if w is not None and w.dim() > 1 else p.numel()
                bound = 1.0 / math.sqrt(fan_in)
                p.uniform_(-bound, bound, generator=gen)
        else:
            raise KeyError(name)

"""fp32 restatement of the UPSTREAM facebookresearch/vggt math used on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  All functions take ``p`` — a flat mapping
``name -> tensor`` (a state_dict or dict(named_parameters())) — plus a key prefix, so the same
code serves the nn.Module shim (oracle/vggt_shim) and state-dict driven parity tests.

Upstream is not vendored by the reference; behaviour follows SURVEY.md Appendix A, cross-checked
against the in-repo sibling code that mirrors it:
  * RoPE math            – /root/reference/aligned_vggt/layers/rope.py:30-96
  * attention structure  – /root/reference/aligned_vggt/layers/cross_attention.py:47-78
  * block structure      – /root/reference/aligned_vggt/layers/cross_attention.py:126-131
  * special-token expand – /root/reference/aligned_vggt/heads/alignment_head.py:543-568
  * position ids         – /root/reference/aligned_vggt/heads/alignment_head.py:301-310
"""
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)


# ----------------------------------------------------------------------------- basic layers
def linear(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, p[name + ".weight"], p.get(name + ".bias"))


def layer_norm(p: Params, name: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    w = p.get(name + ".weight")
    b = p.get(name + ".bias")
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def mlp(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Mlp: fc2(GELU_erf(fc1(x)))  (A.3)."""
    return linear(p, prefix + "fc2", F.gelu(linear(p, prefix + "fc1", x)))


def layer_scale(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    g = p.get(name + ".gamma")
    return x if g is None else x * g


# ----------------------------------------------------------------------------- rotary embedding
def rope_angles(dim: int, n_pos: int, base: float, device, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin tables of shape (n_pos, dim); angles are cast to ``dtype`` BEFORE cos/sin and the
    half table is duplicated (rope.py:46-58)."""
    expo = torch.arange(0, dim, 2, device=device).float() / dim
    inv_freq = 1.0 / (base ** expo)
    pos = torch.arange(n_pos, device=device, dtype=inv_freq.dtype)
    ang = torch.outer(pos, inv_freq).to(dtype)
    ang = torch.cat((ang, ang), dim=-1)
    return ang.cos().to(dtype), ang.sin().to(dtype)


def _rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def rope_apply_1d(tokens: torch.Tensor, positions: torch.Tensor, base: float = 100.0) -> torch.Tensor:
    """tokens (B, heads, N, d), positions (B, N) integer.  rope.py:78-96."""
    d = tokens.shape[-1]
    n_pos = int(positions.max()) + 1
    cos_t, sin_t = rope_angles(d, n_pos, base, tokens.device, tokens.dtype)
    cos = cos_t[positions][:, None]
    sin = sin_t[positions][:, None]
    return tokens * cos + _rotate_half(tokens) * sin


def rope_apply_2d(tokens: torch.Tensor, positions: torch.Tensor, base: float = 100.0) -> torch.Tensor:
    """tokens (B, heads, N, d), positions (B, N, 2) = (y, x).  First half of d rotates with y,
    second half with x (A.4)."""
    half = tokens.shape[-1] // 2
    n_pos = int(positions.max()) + 1
    cos_t, sin_t = rope_angles(half, n_pos, base, tokens.device, tokens.dtype)
    out = []
    for part, idx in ((tokens[..., :half], positions[..., 0]), (tokens[..., half:], positions[..., 1])):
        cos = cos_t[idx][:, None]
        sin = sin_t[idx][:, None]
        out.append(part * cos + _rotate_half(part) * sin)
    return torch.cat(out, dim=-1)


def grid_positions(batch: int, h: int, w: int, device) -> torch.Tensor:
    """PositionGetter: cartesian_prod(arange(h), arange(w)) -> (batch, h*w, 2) int64 (y, x)."""
    ys = torch.arange(h, device=device)
    xs = torch.arange(w, device=device)
    pos = torch.cartesian_prod(ys, xs)
    return pos.view(1, h * w, 2).expand(batch, -1, -1).clone()


def token_positions(batch: int, h: int, w: int, n_special: int, device) -> torch.Tensor:
    """patch positions +1, ``n_special`` leading special tokens at (0,0)
    (alignment_head.py:301-310; same construction in the upstream Aggregator)."""
    pos = grid_positions(batch, h, w, device) + 1
    special = torch.zeros(batch, n_special, 2, dtype=pos.dtype, device=device)
    return torch.cat([special, pos], dim=1)


# ----------------------------------------------------------------------------- attention / block
def attention(p: Params, prefix: str, x: torch.Tensor, num_heads: int,
              pos: Optional[torch.Tensor] = None, rope_base: Optional[float] = None) -> torch.Tensor:
    """Attention.forward (A.3): fused qkv -> optional per-head LayerNorm on q,k -> optional 2-D RoPE ->
    softmax(q k^T / sqrt(d)) v -> proj."""
    B, N, C = x.shape
    d = C // num_heads
    qkv = linear(p, prefix + "qkv", x).reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if prefix + "q_norm.weight" in p:
        q = layer_norm(p, prefix + "q_norm", q)
        k = layer_norm(p, prefix + "k_norm", k)
    if rope_base is not None and pos is not None:
        q = rope_apply_2d(q, pos, rope_base)
        k = rope_apply_2d(k, pos, rope_base)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, N, C)
    return linear(p, prefix + "proj", o)


def block(p: Params, prefix: str, x: torch.Tensor, num_heads: int, pos: Optional[torch.Tensor] = None,
          rope_base: Optional[float] = None, ln_eps: float = 1e-5) -> torch.Tensor:
    """Pre-LN transformer block with LayerScale (A.3)."""
    a = attention(p, prefix + "attn.", layer_norm(p, prefix + "norm1", x, ln_eps), num_heads, pos, rope_base)
    x = x + layer_scale(p, prefix + "ls1", a)
    m = mlp(p, prefix + "mlp.", layer_norm(p, prefix + "norm2", x, ln_eps))
    return x + layer_scale(p, prefix + "ls2", m)


# ----------------------------------------------------------------------------- DINOv2 ViT (A.2)
def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int, antialias: bool = True) -> torch.Tensor:
    """(1, 1+M*M, C) learned table -> (1, 1+gh*gw, C).  Identity when the grid is already M x M."""
    n = pos_embed.shape[1] - 1
    m = int(math.sqrt(n))
    assert m * m == n
    if gh == m and gw == m:
        return pos_embed
    pe = pos_embed.float()
    cls_pe, patch_pe = pe[:, :1], pe[:, 1:]
    c = pe.shape[-1]
    patch_pe = F.interpolate(patch_pe.reshape(1, m, m, c).permute(0, 3, 1, 2), size=(gh, gw),
                             mode="bicubic", antialias=antialias)
    patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, gh * gw, c)
    return torch.cat([cls_pe, patch_pe], dim=1).to(pos_embed.dtype)


def dinov2_patch_tokens(p: Params, prefix: str, images: torch.Tensor, depth: int = 24, num_heads: int = 16,
                        patch: int = 14, n_reg: int = 4) -> torch.Tensor:
    """images (N,3,H,W) already ImageNet-normalised -> x_norm_patchtokens (N, Pp, C)."""
    n, _, H, W = images.shape
    gh, gw = H // patch, W // patch
    x = F.conv2d(images, p[prefix + "patch_embed.proj.weight"], p[prefix + "patch_embed.proj.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([p[prefix + "cls_token"].expand(n, -1, -1), x], dim=1)
    x = x + interpolate_pos_embed(p[prefix + "pos_embed"], gh, gw)
    if n_reg:
        x = torch.cat([x[:, :1], p[prefix + "register_tokens"].expand(n, -1, -1), x[:, 1:]], dim=1)
    for i in range(depth):
        x = block(p, f"{prefix}blocks.{i}.", x, num_heads, ln_eps=1e-6)
    x = layer_norm(p, prefix + "norm", x, 1e-6)
    return x[:, 1 + n_reg:]


# ----------------------------------------------------------------------------- Aggregator (A.1)
def expand_special(tok: torch.Tensor, B: int, S: int) -> torch.Tensor:
    """(1,2,X,C) -> (B,S,X,C): index 0 for frame 0, index 1 for frames 1..S-1."""
    first = tok[:, 0:1].expand(B, 1, *tok.shape[2:])
    rest = tok[:, 1:2].expand(B, S - 1, *tok.shape[2:])
    return torch.cat([first, rest], dim=1)


def aggregator_prepare(p: Params, prefix: str, images: torch.Tensor, dino_depth: int = 24, num_heads: int = 16,
                       patch: int = 14, n_reg: int = 4):
    """Normalise, DINO patch embed, prepend camera+register tokens.  Returns (tokens (B*S,P,C), pos (B*S,P,2))."""
    B, S, _, H, W = images.shape
    mean = torch.tensor(RESNET_MEAN, dtype=images.dtype, device=images.device).view(1, 1, 3, 1, 1)
    std = torch.tensor(RESNET_STD, dtype=images.dtype, device=images.device).view(1, 1, 3, 1, 1)
    x = ((images - mean) / std).view(B * S, 3, H, W)
    patch_tok = dinov2_patch_tokens(p, prefix + "patch_embed.", x, dino_depth, num_heads, patch, n_reg)
    C = patch_tok.shape[-1]
    cam = expand_special(p[prefix + "camera_token"], B, S).reshape(B * S, 1, C)
    reg = expand_special(p[prefix + "register_token"], B, S).reshape(B * S, n_reg, C)
    tokens = torch.cat([cam, reg, patch_tok], dim=1)
    pos = token_positions(B * S, H // patch, W // patch, 1 + n_reg, images.device)
    return tokens, pos


def aggregator_forward(p: Params, prefix: str, images: torch.Tensor, depth: int = 24, dino_depth: int = 24,
                       num_heads: int = 16, patch: int = 14, n_reg: int = 4, rope_base: float = 100.0,
                       keep: Optional[Tuple[int, ...]] = None):
    """Returns (list of (B,S,P,2C) for each alternating-attention layer, patch_start_idx).
    ``keep`` limits the returned list to those layer ids (others are None) to save memory."""
    B, S = images.shape[:2]
    tokens, pos = aggregator_prepare(p, prefix, images, dino_depth, num_heads, patch, n_reg)
    _, P, C = tokens.shape
    out = []
    for i in range(depth):
        tokens = block(p, f"{prefix}frame_blocks.{i}.", tokens.view(B * S, P, C), num_heads, pos, rope_base)
        frame_out = tokens.view(B, S, P, C)
        tokens = block(p, f"{prefix}global_blocks.{i}.", tokens.view(B, S * P, C), num_heads,
                       pos.view(B, S * P, 2), rope_base)
        global_out = tokens.view(B, S, P, C)
        if keep is None or i in keep:
            out.append(torch.cat([frame_out, global_out], dim=-1))
        else:
            out.append(None)
    return out, 1 + n_reg


# ----------------------------------------------------------------------------- rotation / pose_enc (A.7)
def quat_to_mat(q: torch.Tensor) -> torch.Tensor:
    """Scalar-last (x,y,z,w) quaternion -> rotation matrix; not required to be unit norm."""
    i, j, k, r = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_pos(x: torch.Tensor) -> torch.Tensor:
    r = torch.zeros_like(x)
    m = x > 0
    r[m] = torch.sqrt(x[m])
    return r


def mat_to_quat(m: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> (x,y,z,w), w >= 0 (PyTorch3D matrix_to_quaternion, reordered)."""
    bd = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(bd + (9,)), dim=-1)
    q_abs = _sqrt_pos(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                   1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(flr))
    out = cand[F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(bd + (4,))
    out = out[..., [1, 2, 3, 0]]
    return torch.where(out[..., 3:4] < 0, -out, out)


def closed_form_inverse_se3(se3, R=None, T=None):
    """[R t; 0 1]^-1 = [R^T, -R^T t; 0 1] for (N,4,4) or (N,3,4); torch or numpy."""
    import numpy as np
    is_np = isinstance(se3, np.ndarray)
    if se3.shape[-2:] not in ((4, 4), (3, 4)):
        raise ValueError(f"se3 must be of shape (N,4,4), got {se3.shape}.")
    if R is None:
        R = se3[:, :3, :3]
    if T is None:
        T = se3[:, :3, 3:]
    if is_np:
        Rt = np.transpose(R, (0, 2, 1))
        top_right = -np.matmul(Rt, T)
        inv = np.tile(np.eye(4), (len(R), 1, 1))
    else:
        Rt = R.transpose(1, 2)
        top_right = -torch.bmm(Rt, T)
        inv = torch.eye(4, 4)[None].repeat(len(R), 1, 1).to(R.dtype).to(R.device)
    inv[:, :3, :3] = Rt
    inv[:, :3, 3:] = top_right
    return inv


def extri_intri_to_pose_encoding(extrinsics, intrinsics, image_size_hw=None, pose_encoding_type="absT_quaR_FoV"):
    assert pose_encoding_type == "absT_quaR_FoV"
    R = extrinsics[:, :, :3, :3]
    T = extrinsics[:, :, :3, 3]
    quat = mat_to_quat(R)
    H, W = image_size_hw
    fov_h = 2 * torch.atan((H / 2) / intrinsics[..., 1, 1])
    fov_w = 2 * torch.atan((W / 2) / intrinsics[..., 0, 0])
    return torch.cat([T, quat, fov_h[..., None], fov_w[..., None]], dim=-1).float()


def pose_encoding_to_extri_intri(pose_encoding, image_size_hw=None, pose_encoding_type="absT_quaR_FoV",
                                 build_intrinsics=True):
    assert pose_encoding_type == "absT_quaR_FoV"
    T = pose_encoding[..., :3]
    quat = pose_encoding[..., 3:7]
    fov_h = pose_encoding[..., 7]
    fov_w = pose_encoding[..., 8]
    R = quat_to_mat(quat)
    extrinsics = torch.cat([R, T[..., None]], dim=-1)
    intrinsics = None
    if build_intrinsics:
        H, W = image_size_hw
        fy = (H / 2.0) / torch.tan(fov_h / 2.0)
        fx = (W / 2.0) / torch.tan(fov_w / 2.0)
        intrinsics = torch.zeros(pose_encoding.shape[:2] + (3, 3), device=pose_encoding.device)
        intrinsics[..., 0, 0] = fx
        intrinsics[..., 1, 1] = fy
        intrinsics[..., 0, 2] = W / 2
        intrinsics[..., 1, 2] = H / 2
        intrinsics[..., 2, 2] = 1.0
    return extrinsics, intrinsics


# ----------------------------------------------------------------------------- CameraHead (A.5)
def camera_head_forward(p: Params, prefix: str, tokens_last: torch.Tensor, num_iterations: int = 4,
                        trunk_depth: int = 4, num_heads: int = 16):
    """tokens_last (B,S,P,2C) = last tapped aggregator layer.  Returns list of (B,S,9)."""
    tok = layer_norm(p, prefix + "token_norm", tokens_last[:, :, 0])
    B, S, C = tok.shape
    pred = None
    outs = []
    for _ in range(num_iterations):
        if pred is None:
            inp = linear(p, prefix + "embed_pose", p[prefix + "empty_pose_tokens"].expand(B, S, -1))
        else:
            inp = linear(p, prefix + "embed_pose", pred.detach())
        mod = linear(p, prefix + "poseLN_modulation.1", F.silu(inp))
        shift, scale, gate = mod.chunk(3, dim=-1)
        normed = F.layer_norm(tok, (C,), None, None, 1e-6)
        x = gate * (normed * (1 + scale) + shift) + tok
        for i in range(trunk_depth):
            x = block(p, f"{prefix}trunk.{i}.", x, num_heads)
        delta = mlp(p, prefix + "pose_branch.", layer_norm(p, prefix + "trunk_norm", x))
        pred = delta if pred is None else pred + delta
        act = torch.cat([pred[..., :3], pred[..., 3:7], F.relu(pred[..., 7:])], dim=-1)
        outs.append(act)
    return outs


# ----------------------------------------------------------------------------- DPTHead (A.6)
# Restated from the public upstream `vggt/heads/dpt_head.py`, `head_act.py`, `heads/utils.py` (upstream is not in the
# container: behaviour recalled).  Call sites: /root/reference/aligned_vggt/models/
# featureAligned_vggt.py:28-29 (construction: depth_head = DPTHead(dim_in=2*embed_dim, output_dim=2, activation="exp",
# conf_activation="expp1"); point_head = DPTHead(dim_in=2*embed_dim, output_dim=4, activation="inv_log",
# conf_activation="expp1")) and :166-168, :183-185 (forward, fp32: autocast disabled at :103).
# Recalled-uncertain item (?): the fusion blocks are built with nn.ReLU(inplace=True), so a ResidualConvUnit adds
# relu(x), not x, on its skip path; `relu_inplace=False` gives the other reading.
DPT_OUT_CHANNELS = (256, 512, 1024, 1024)
DPT_FEATURES = 256


def make_sincos_pos_embed(embed_dim: int, pos: torch.Tensor, omega_0: float = 100.0) -> torch.Tensor:
    omega = torch.arange(embed_dim // 2, dtype=torch.double, device=pos.device)
    omega = 1.0 / omega_0 ** (omega / (embed_dim / 2.0))
    out = torch.einsum("m,d->md", pos.reshape(-1).double(), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1).float()


def create_uv_grid(width: int, height: int, aspect_ratio: float, dtype=torch.float32, device=None) -> torch.Tensor:
    diag = (aspect_ratio ** 2 + 1.0) ** 0.5
    span_x, span_y = aspect_ratio / diag, 1.0 / diag
    xs = torch.linspace(-span_x * (width - 1) / width, span_x * (width - 1) / width, steps=width, dtype=dtype, device=device)
    ys = torch.linspace(-span_y * (height - 1) / height, span_y * (height - 1) / height, steps=height, dtype=dtype, device=device)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")
    return torch.stack((uu, vv), dim=-1)  # (height, width, 2)


def dpt_pos_embed(x: torch.Tensor, W: int, H: int, ratio: float = 0.1) -> torch.Tensor:
    """x (N,C,h,w) + ratio * sincos embedding of the uv grid (x-embedding in channels [0,C/2), y in [C/2,C))."""
    h, w, C = x.shape[-2], x.shape[-1], x.shape[1]
    grid = create_uv_grid(w, h, aspect_ratio=W / H, dtype=x.dtype, device=x.device)
    flat = grid.reshape(-1, 2)
    emb = torch.cat([make_sincos_pos_embed(C // 2, flat[:, 0]), make_sincos_pos_embed(C // 2, flat[:, 1])], dim=-1)
    emb = emb.view(h, w, C) * ratio
    return x + emb.permute(2, 0, 1)[None].to(x.dtype)


def _conv(p: Params, name: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, p[name + ".weight"], p.get(name + ".bias"), stride=stride, padding=padding)


def _residual_conv_unit(p: Params, pre: str, x: torch.Tensor, relu_inplace: bool) -> torch.Tensor:
    a = F.relu(x)
    out = _conv(p, pre + "conv2", F.relu(_conv(p, pre + "conv1", a, padding=1)), padding=1)
    return out + (a if relu_inplace else x)


def _fusion_block(p: Params, pre: str, x0: torch.Tensor, x1: Optional[torch.Tensor], size, relu_inplace: bool) -> torch.Tensor:
    out = x0
    if x1 is not None:  # has_residual
        out = out + _residual_conv_unit(p, pre + "resConfUnit1.", x1, relu_inplace)
    out = _residual_conv_unit(p, pre + "resConfUnit2.", out, relu_inplace)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=tuple(size), mode="bilinear", align_corners=True)
    return _conv(p, pre + "out_conv", out)


def dpt_activate(out: torch.Tensor, activation: str, conf_activation: str):
    """head_act.activate_head: (N,C,H,W) -> pred (N,H,W,C-1), conf (N,H,W)."""
    fmap = out.permute(0, 2, 3, 1)
    xyz, conf = fmap[..., :-1], fmap[..., -1]
    if activation == "exp":
        pts = torch.exp(xyz)
    elif activation == "inv_log":
        pts = torch.sign(xyz) * torch.expm1(torch.abs(xyz))
    else:
        raise ValueError(f"Unknown activation: {activation}")
    if conf_activation != "expp1":
        raise ValueError(f"Unknown conf_activation: {conf_activation}")
    return pts, 1 + conf.exp()


def dpt_head_forward(p: Params, prefix: str, taps, image_hw, patch_start_idx: int = 5, activation: str = "inv_log",
                     conf_activation: str = "expp1", patch_size: int = 14, relu_inplace: bool = True):
    """taps: the 4 tapped aggregator outputs (layers 4, 11, 17, 23), each (B,S,P,2C).  Returns pred (B,S,H,W,od-1),
    conf (B,S,H,W).  Per-frame computation, so upstream's frames_chunk_size split does not change the result."""
    H, W = image_hw
    B, S = taps[0].shape[:2]
    ph, pw = H // patch_size, W // patch_size
    feats = []
    for i, t in enumerate(taps):
        x = t[:, :, patch_start_idx:].reshape(B * S, -1, t.shape[-1])
        x = layer_norm(p, prefix + "norm", x)
        x = x.permute(0, 2, 1).reshape(B * S, x.shape[-1], ph, pw)
        x = _conv(p, f"{prefix}projects.{i}", x)
        x = dpt_pos_embed(x, W, H)
        if i == 0:
            x = F.conv_transpose2d(x, p[prefix + "resize_layers.0.weight"], p[prefix + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, p[prefix + "resize_layers.1.weight"], p[prefix + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = _conv(p, prefix + "resize_layers.3", x, stride=2, padding=1)
        feats.append(x)
    s = prefix + "scratch."
    rn = [_conv(p, f"{s}layer{i + 1}_rn", feats[i], padding=1) for i in range(4)]
    out = _fusion_block(p, s + "refinenet4.", rn[3], None, rn[2].shape[2:], relu_inplace)
    out = _fusion_block(p, s + "refinenet3.", out, rn[2], rn[1].shape[2:], relu_inplace)
    out = _fusion_block(p, s + "refinenet2.", out, rn[1], rn[0].shape[2:], relu_inplace)
    out = _fusion_block(p, s + "refinenet1.", out, rn[0], None, relu_inplace)
    out = _conv(p, s + "output_conv1", out, padding=1)
    out = F.interpolate(out, size=(ph * patch_size, pw * patch_size), mode="bilinear", align_corners=True)
    out = dpt_pos_embed(out, W, H)
    out = _conv(p, s + "output_conv2.2", F.relu(_conv(p, s + "output_conv2.0", out, padding=1)))
    pred, conf = dpt_activate(out, activation, conf_activation)
    return pred.view(B, S, *pred.shape[1:]), conf.view(B, S, *conf.shape[1:])

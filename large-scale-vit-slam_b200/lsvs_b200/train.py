"""Training path of the alignment head: forward WITH an autograd graph, so that `mode: train` of the
reference (only the head trains; its blocks run under torch.utils.checkpoint, alignment_head.py:361,385,498,527) works on the
drop-in.  The inference path stays the fused engine (csrc/engine.cu); this module is selected when gradients are enabled and a
head parameter requires them.

Division of labour.  The graph is torch's (autograd bookkeeping, LayerNorm / GELU / RoPE / residual glue and the 0.1 % of the
FLOPs spent in the fp32 decode); its heavy nodes are this repository's kernels:
  * every Linear of the 8 token blocks (97 % of the head's FLOPs), forward, input gradient and weight gradient, is the tcgen05
    GEMM (lsvs_gemm_bf16): y = x W^T, dx = dy W, dW = dy^T x, with bf16 operands (`precision=0`, the reference's bf16-mixed
    training arithmetic) or split-bf16 fp32-class operands (`precision=1`, used by the gradient checks);
  * attention forward / backward are lsvs_attention_f32_train / lsvs_attention_f32_backward (fp32, csrc/precise.cu).
No activation checkpointing (the reference needs it for 24 GB cards; a 32-frame chunk keeps < 6 GB of activations here).

Follows /root/reference/aligned_vggt/heads/alignment_head.py:224-345 (forward), :347-390 (frame / temporal attention on the raw
(B*P1, S, C) view), :427-540 (decode), aligned_vggt/layers/cross_attention.py:47-131, aligned_vggt/layers/gated_update.py:43-78.
"""
import ctypes
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import native as _n
from . import ops

_ll, _i, _f = ctypes.c_longlong, ctypes.c_int, ctypes.c_float


# ------------------------------------------------------------------------------------------------ native autograd nodes
def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _operand(t: torch.Tensor, precision: int, role: str) -> torch.Tensor:
    """fp32 (rows, K) -> GEMM operand: bf16 (rows, K), or the fp32-class split (rows, 3K): [hi | lo | hi] for the A role,
    [hi | hi | lo] for the W role (include/lsvs_b200.h, 'fp32-class operators')."""
    t = t.contiguous()
    if precision == 0:
        return t.to(torch.bfloat16)
    rows, K = t.shape
    if role == "a":
        out = torch.empty(rows, 3 * K, dtype=torch.bfloat16, device=t.device)
        _n.check(_n.lib().lsvs_cast_split(_n.ptr(t), _ll(K), _n.ptr(out), _ll(3 * K), _ll(rows), _i(K), _i(0), _n.stream_ptr()), "cast_split")
        return out
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], dim=1)


def _pad_cols(t: torch.Tensor, cols: int) -> torch.Tensor:
    return t if t.shape[1] == cols else F.pad(t, (0, cols - t.shape[1]))


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b on the tensor cores; backward = two more GEMMs on the same kernel."""

    @staticmethod
    def forward(ctx, x, W, b, precision):
        ctx.save_for_backward(x, W)
        ctx.precision, ctx.has_bias = precision, b is not None
        return ops.gemm(_operand(x, precision, "a"), _operand(W, precision, "w"), ops.EPI_BIAS_F32, bias=None if b is None else b.contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        p = ctx.precision
        dy = dy.contiguous()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:     # dx (M,K) = dy (M,N) W (N,K): reduction over N, the W-role operand is W^T
            dx = ops.gemm(_operand(dy, p, "a"), _operand(W.t(), p, "w"), ops.EPI_BIAS_F32)
        if ctx.needs_input_grad[1]:     # dW (N,K) = dy^T (N,M) x (M,K): reduction over the M rows (zero-padded to the k-block)
            Mp = _round_up(x.shape[0], 64)
            dW = ops.gemm(_operand(_pad_cols(dy.t(), Mp), p, "a"), _operand(_pad_cols(x.t(), Mp), p, "w"), ops.EPI_BIAS_F32)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dW, db, None


def linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor], precision: int) -> torch.Tensor:
    """nn.Linear over the last dim of x; output width must be a multiple of 128, input width of 64 (true for every Linear of the
    token blocks)."""
    lead = x.shape[:-1]
    y = _LinearTC.apply(x.reshape(-1, x.shape[-1]), W, b, precision)
    return y.view(*lead, W.shape[0])


class _AttentionF32(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v per (batch, head): q (batches*Lq, heads*hd), k / v (batches*Lk, heads*hd), fp32."""

    @staticmethod
    def forward(ctx, q, k, v, batches, heads, hd, Lq, Lk):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        D = heads * hd
        o = torch.empty_like(q)
        lse = torch.empty(batches, heads, Lq, dtype=torch.float32, device=q.device)
        _n.check(_n.lib().lsvs_attention_f32_train(_n.ptr(q), _ll(D), _n.ptr(k), _ll(D), _n.ptr(v), _ll(D), _n.ptr(o), _ll(D), _n.ptr(lse),
                                                   _i(batches), _i(heads), _i(hd), _i(Lq), _i(Lk), _f(hd ** -0.5), _n.stream_ptr()),
                 "attention_f32_train")
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.shape = (batches, heads, hd, Lq, Lk)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        batches, heads, hd, Lq, Lk = ctx.shape
        D = heads * hd
        do = do.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dbuf = torch.empty_like(lse)
        _n.check(_n.lib().lsvs_attention_f32_backward(_n.ptr(q), _ll(D), _n.ptr(k), _ll(D), _n.ptr(v), _ll(D), _n.ptr(o), _n.ptr(do), _ll(D),
                                                      _n.ptr(lse), _n.ptr(dbuf), _n.ptr(dq), _ll(D), _n.ptr(dk), _ll(D), _n.ptr(dv), _ll(D),
                                                      _i(batches), _i(heads), _i(hd), _i(Lq), _i(Lk), _f(hd ** -0.5), _n.stream_ptr()),
                 "attention_f32_backward")
        return dq, dk, dv, None, None, None, None, None


def attention(q, k, v, batches, heads, hd, Lq, Lk):
    return _AttentionF32.apply(q, k, v, batches, heads, hd, Lq, Lk)


# ------------------------------------------------------------------------------------------------ glue (torch, differentiable)
def _rope_tables(dim: int, n_pos: int, base: float, device):
    """cos / sin (n_pos, dim): angles pos * base^(-2j/dim), the half table duplicated (aligned_vggt/layers/rope.py:46-58)."""
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, device=device).float() / dim))
    ang = torch.outer(torch.arange(n_pos, device=device, dtype=torch.float32), inv)
    ang = torch.cat([ang, ang], dim=-1)
    return ang.cos(), ang.sin()


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def rope_1d(x, pos, base):
    """x (rows, heads, hd), pos (rows,) integer."""
    cos, sin = _rope_tables(x.shape[-1], int(pos.max()) + 1, base, x.device)
    return x * cos[pos][:, None] + _rot_half(x) * sin[pos][:, None]


def rope_2d(x, pos, base):
    """x (rows, heads, hd), pos (rows, 2) = (y, x): first half of hd rotates with y, second with x."""
    half = x.shape[-1] // 2
    cos, sin = _rope_tables(half, int(pos.max()) + 1, base, x.device)
    parts = []
    for part, idx in ((x[..., :half], pos[:, 0]), (x[..., half:], pos[:, 1])):
        parts.append(part * cos[idx][:, None] + _rot_half(part) * sin[idx][:, None])
    return torch.cat(parts, dim=-1)


def _ln(x, mod, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), mod.weight, mod.bias, eps)


def _expand_special(tok, B, S):
    """(1,2,X,C) -> (B,S,X,C): variant 0 for frame 0, variant 1 for the others (alignment_head.py:543-568)."""
    return torch.cat([tok[:, 0:1].expand(B, 1, *tok.shape[2:]), tok[:, 1:2].expand(B, S - 1, *tok.shape[2:])], dim=1)


def _mlp(x, mlp, precision):
    return linear(F.gelu(linear(x, mlp.fc1.weight, mlp.fc1.bias, precision)), mlp.fc2.weight, mlp.fc2.bias, precision)


def _self_block(x, blk, heads, batches, L, pos2d, base, precision):
    """UPSTREAM Block on (batches*L, C) rows: x += ls1 * attn(norm1 x); x += ls2 * mlp(norm2 x)."""
    rows, C = x.shape
    hd = C // heads
    qkv = linear(_ln(x, blk.norm1), blk.attn.qkv.weight, blk.attn.qkv.bias, precision).view(rows, 3, heads, hd)
    q = rope_2d(_ln(qkv[:, 0], blk.attn.q_norm), pos2d, base)
    k = rope_2d(_ln(qkv[:, 1], blk.attn.k_norm), pos2d, base)
    o = attention(q.reshape(rows, C), k.reshape(rows, C), qkv[:, 2].reshape(rows, C), batches, heads, hd, L, L)
    x = x + blk.ls1.gamma * linear(o, blk.attn.proj.weight, blk.attn.proj.bias, precision)
    return x + blk.ls2.gamma * _mlp(_ln(x, blk.norm2), blk.mlp, precision)


def _cross_block_tc(x, y, blk, heads, groups, Lq, Lk, pos_q, pos_k, base, precision):
    """CrossAttentionBlock (cross_attention.py:126-131) on x (groups*Lq, C), y (groups*Lk, C), tensor-core Linears."""
    C = x.shape[1]
    hd = C // heads
    xn, yn = _ln(x, blk.norm1), _ln(y, blk.norm3)
    q = linear(xn, blk.attn.q.weight, blk.attn.q.bias, precision).view(-1, heads, hd)
    k = linear(yn, blk.attn.k.weight, blk.attn.k.bias, precision).view(-1, heads, hd)
    v = linear(yn, blk.attn.v.weight, blk.attn.v.bias, precision)
    q = rope_1d(_ln(q, blk.attn.q_norm), pos_q, base)
    k = rope_1d(_ln(k, blk.attn.k_norm), pos_k, base)
    o = attention(q.reshape(-1, C), k.reshape(-1, C), v, groups, heads, hd, Lq, Lk)
    x = x + blk.ls1.gamma * linear(o, blk.attn.proj.weight, blk.attn.proj.bias, precision)
    return x + blk.ls2.gamma * _mlp(_ln(x, blk.norm2), blk.mlp, precision)


def _cross_block_small(x, y, blk, heads, pos_q, pos_k, base):
    """The decode's cross blocks (dim 512, a few dozen tokens): plain fp32 torch, as the reference decodes with autocast off (:340)."""
    B, N, C = x.shape
    M = y.shape[1]
    hd = C // heads
    xn, yn = _ln(x, blk.norm1), _ln(y, blk.norm3)
    q = F.linear(xn, blk.attn.q.weight, blk.attn.q.bias).view(B * N, heads, hd)
    k = F.linear(yn, blk.attn.k.weight, blk.attn.k.bias).view(B * M, heads, hd)
    v = F.linear(yn, blk.attn.v.weight, blk.attn.v.bias).view(B, M, heads, hd).transpose(1, 2)
    q = rope_1d(_ln(q, blk.attn.q_norm), pos_q.reshape(-1), base).view(B, N, heads, hd).transpose(1, 2)
    k = rope_1d(_ln(k, blk.attn.k_norm), pos_k.reshape(-1), base).view(B, M, heads, hd).transpose(1, 2)
    att = torch.softmax((q * hd ** -0.5) @ k.transpose(-2, -1), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    x = x + blk.ls1.gamma * F.linear(o, blk.attn.proj.weight, blk.attn.proj.bias)
    h = F.linear(F.gelu(F.linear(_ln(x, blk.norm2), blk.mlp.fc1.weight, blk.mlp.fc1.bias)), blk.mlp.fc2.weight, blk.mlp.fc2.bias)
    return x + blk.ls2.gamma * h


def _gated_update(gu, memory, update):
    """GatedUpdate.forward (gated_update.py:43-78): memory (B,N,D) unit rows, update (B,1,D)."""
    N = memory.shape[1]
    u_norm = update.norm(dim=-1, keepdim=True)
    mem_scaled = memory * u_norm
    inp = torch.cat([update.expand_as(memory), mem_scaled, memory.mean(dim=1, keepdim=True).expand_as(memory) * u_norm], dim=-1)
    deltas = []
    for i in range(N):
        m = getattr(gu.delta_mlps, str(i))
        l0, l2 = getattr(m, "0"), getattr(m, "2")
        deltas.append(F.linear(F.gelu(F.linear(inp[:, i], l0.weight, l0.bias)), l2.weight, l2.bias))
    diff = torch.stack(deltas, dim=1) - memory
    g0, g2 = getattr(gu.gate_mlp, "0"), getattr(gu.gate_mlp, "2")
    gate = torch.sigmoid(F.linear(F.gelu(F.linear(torch.cat([diff, mem_scaled], dim=-1), g0.weight, g0.bias)), g2.weight, g2.bias))
    orth = diff - (diff * memory).sum(-1, keepdim=True) * memory
    return F.normalize(memory + gate * F.normalize(orth, dim=-1), dim=-1)


def _decode(head, tok, memory_tokens, base):
    """AlignmentHead._decode_alignments (:427-540), fp32: tok (B,S,1024) processed per-frame alignment tokens."""
    B, S, _ = tok.shape
    dev, nm, heads = tok.device, head.num_memory_tokens, 8
    tokens = _ln(F.linear(tok, head.project_dec.weight, head.project_dec.bias), head.dec_norm)
    C = tokens.shape[-1]
    k_chunk = torch.arange(0, S + nm, device=dev)
    directional = None
    kv = tokens
    if nm > 0:
        k_chunk[-nm:] += S                                                                   # :451-452
        mean_norm = tokens.norm(dim=-1).mean(dim=-1, keepdim=True).unsqueeze(1)              # :469
        if memory_tokens is None:
            mem = head.memory_token.expand(B, -1, -1)
            init = F.linear(tokens[:, 0], head.frame_proj.weight, head.frame_proj.bias).view(B, -1, C)
            a = torch.sigmoid(head.alpha)
            directional = (1 - a) * mem + a * (init / init.norm(dim=-1, keepdim=True).clamp_min(1e-6))   # :475-478 (not re-normalised)
            effective = mem * mean_norm                                                      # :479 (the un-blended memory)
        else:
            assert memory_tokens.shape[0] == B, "Memory tokens must have same batch dimension as frame tokens"
            directional, effective = memory_tokens, memory_tokens * mean_norm
        kv = torch.cat([tokens, effective], dim=1)
    zero = torch.zeros(B, 1, dtype=torch.long, device=dev)
    k_chunk = k_chunk.view(1, -1).expand(B, -1)
    chunk_tok = tokens[:, :1]
    for i in range(2):
        chunk_tok = _cross_block_small(chunk_tok, kv, getattr(head.chunk_cross_blocks, str(i)), heads, zero, k_chunk, base)
    new_memory = _gated_update(head.gated_update, directional, chunk_tok) if nm > 0 else memory_tokens
    chunk_n = _ln(chunk_tok, head.chunk_norm)
    dec = lambda m, x: F.linear(F.gelu(F.linear(x, m.fc1.weight, m.fc1.bias)), m.fc2.weight, m.fc2.bias)
    frame_se3 = tokens.new_zeros(B, 0, 7)
    if S > 1:
        frame_tok = tokens[:, 1:]
        q_frame = torch.arange(1, S, device=dev).view(1, S - 1).expand(B, -1)
        for i in range(2):
            frame_tok = _cross_block_small(frame_tok, chunk_n, getattr(head.frame_cross_blocks, str(i)), heads, q_frame, zero, base)
        frame_se3 = dec(head.frame_se3_decoder, _ln(frame_tok, head.frame_norm))
    sim3 = dec(head.chunk_sim3_decoder, chunk_n)
    return torch.cat([sim3[..., :7], torch.exp(sim3[..., 7:])], dim=-1), frame_se3, new_memory       # :538


# ------------------------------------------------------------------------------------------------ the head
def alignment_head_forward_train(head, tokens: torch.Tensor, image_size: Tuple[int, int], next_num_overlap: int,
                                 overlap_tokens: Optional[torch.Tensor] = None, memory_tokens: Optional[torch.Tensor] = None,
                                 precision: int = 0):
    """AlignmentHead.forward with an autograd graph.  Same inputs / outputs as the inference path; gradients reach every head
    parameter and, for back-propagation through time over chunks, the `overlap_tokens` / `memory_tokens` of the previous chunk."""
    if not tokens.is_cuda:
        raise _n.NativeError("alignment head training path needs CUDA tensors (no CPU fallback on this path)")
    H, W = image_size
    B, S, P, _ = tokens.shape
    gh, gw = H // head.patch_size, W // head.patch_size
    if gh * gw + 5 != P:
        raise ValueError(f"Size of tokens and image do not match (P={P}, grid {gh}x{gw})")
    heads, base, D, P1 = 8, head.rope_freq, 1024, P + 1
    first = overlap_tokens is None
    if not first:
        assert overlap_tokens.shape[0] == B and overlap_tokens.shape[2] == P1 and overlap_tokens.shape[3] == D, \
            "Size of tokens and overlap tokens must match"
    T = S if first else overlap_tokens.shape[1]
    dev = tokens.device
    x = _ln(linear(tokens.float(), head.project_in.weight, head.project_in.bias, precision), head.token_norm)       # :242-247
    x = torch.cat([_expand_special(head.per_frame_alignment_token, B, S), x], dim=2)                              # :269-270
    # positions: 2-D (+1, six specials at 0) for the frame blocks (:301-310), temporal ids (:279-285)
    ys, xs = torch.meshgrid(torch.arange(gh, device=dev), torch.arange(gw, device=dev), indexing="ij")
    pos2d = torch.cat([torch.zeros(6, 2, dtype=torch.long, device=dev), torch.stack([ys.reshape(-1), xs.reshape(-1)], -1) + 1], 0).repeat(B * S, 1)
    ids = torch.arange(S, device=dev)
    q_ids = ids if first else ids + (S - (T - 1))
    k_ids = ids if first else torch.cat([ids[:1], ids[-(T - 1):]])
    x = x.reshape(B * S * P1, D)
    y_ctx = None if first else overlap_tokens.float().reshape(B * T * P1, D)
    for i in range(head.depth_aa):                                                                              # :317-335
        x = _self_block(x, getattr(head.frame_blocks, str(i)), heads, B * S, P1, pos2d, base, precision)
        # temporal cross attention on the RAW (B*P1, S, C) view (:372-377): groups of S consecutive flat rows
        y = x if first else y_ctx
        x = _cross_block_tc(x, y, getattr(head.temporal_blocks, str(i)), heads, B * P1, S, T, q_ids.repeat(B * P1), k_ids.repeat(B * P1),
                            base, precision)
    x = x.view(B, S, P1, D)
    sim3, se3, memory = _decode(head, x[:, :, 0], memory_tokens, base)
    new_overlap = torch.cat([x[:, :1], x[:, S - next_num_overlap:]], dim=1).contiguous()                          # :343
    return sim3, se3, memory, new_overlap


def wants_training_path(module) -> bool:
    """The autograd path is taken in train() mode with gradients enabled and at least one trainable alignment-head parameter; everything
    else (eval(), torch.no_grad(), a fully frozen head) runs the fused inference engine."""
    return module.training and torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


def pose_chain_train(chunk_sim3_enc, frame_se3_enc, cam_enc, prev_pose_enc, overlap: int, image_hw, gt_mean=None):
    """Differentiable twin of lsvs_pose_chain (featureAligned_vggt.py:97-143, :190-196) for the training losses on `pose_enc`
    (training/loss.py:133-146): torch ops on (B,S,4,4) matrices.  Returns (aligned pose_enc (B,S,9), point transform (B,4,4),
    chunk scale (B,))."""
    from aligned_vggt.utils.data import extri_to_pose_encoding, pose_encoding_to_extri
    from aligned_vggt.utils.geometry import averagePoseEncodings
    from . import posemath as pm
    B, S, _ = cam_enc.shape
    chunk_se3 = pose_encoding_to_extri(chunk_sim3_enc)                       # (B,1,4,4)
    scale = chunk_sim3_enc[..., -1].reshape(B)
    per_frame = torch.cat([chunk_se3, pose_encoding_to_extri(frame_se3_enc) @ chunk_se3], dim=1)            # (B,S,4,4)
    extr, intr = pm.pose_encoding_to_extri_intri(cam_enc.float(), image_hw)
    extr = pm.to_homogeneous(extr)
    point_identity = extr[:, 0].detach().clone()
    extr = extr @ pm.inverse_se3(extr[:, 0]).view(B, 1, 4, 4)               # first pose = identity
    extr = torch.cat([extr[..., :3, :3], extr[..., :3, 3:] * scale.view(B, 1, 1, 1)], dim=-1)                 # t *= chunk scale
    extr = pm.to_homogeneous(extr)
    if prev_pose_enc is not None:
        if gt_mean is not None:
            mean_T = gt_mean.to(extr)
        else:
            ctx = pose_encoding_to_extri(prev_pose_enc[:, -overlap:])
            cams = pm.inverse_se3(extr[:, :overlap]) @ ctx
            mean_T = pose_encoding_to_extri(averagePoseEncodings(extri_to_pose_encoding(cams))) if overlap > 1 else cams
        per_frame = per_frame @ mean_T
        point_T = pm.inverse_se3(per_frame[:, 0]) @ point_identity
    else:
        point_T = point_identity
    aligned = extr @ per_frame
    return pm.extri_intri_to_pose_encoding(aligned, intr, image_hw), point_T, scale


def apply_sim3_points_train(points, T, s):
    """p' = T[:3,:3] (s p) + T[:3,3] with gradients to T and s (alignment.py:491-526)."""
    B = points.shape[0]
    R, t = T[:, :3, :3], T[:, :3, 3]
    flat = points.reshape(B, -1, 3) * s.view(B, 1, 1)
    return (flat @ R.transpose(1, 2) + t[:, None]).view_as(points)

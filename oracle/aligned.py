"""fp32 restatement of the reference-owned code on the hot path (TEST INFRASTRUCTURE).

Every function cites the /root/reference file:line it follows.  ``oracle/make_golden.py`` checks
each of them against the real reference modules (imported unmodified on top of oracle/vggt_shim)
and writes tests/golden/*.npz from those reference outputs.

``p`` is a flat ``name -> tensor`` mapping (state_dict); ``pre`` a key prefix.
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import functional as OF

Params = Dict[str, torch.Tensor]


# =============================================================================================
# layers
# =============================================================================================
def cross_attention(p: Params, pre: str, x: torch.Tensor, y: torch.Tensor, pos_q, pos_k, heads: int,
                    rope_base: Optional[float] = 100.0) -> torch.Tensor:
    """CrossAttention.forward — aligned_vggt/layers/cross_attention.py:47-78.
    Separate q/k/v projections, per-head LayerNorm on q and k, 1-D RoPE with separate query / key
    position ids, softmax(q k^T / sqrt(d)) v (the reference's all-true mask is a no-op), proj."""
    B, N, C = x.shape
    M = y.shape[1]
    d = C // heads
    q = OF.linear(p, pre + "q", x).reshape(B, N, heads, d).transpose(1, 2)
    k = OF.linear(p, pre + "k", y).reshape(B, M, heads, d).transpose(1, 2)
    v = OF.linear(p, pre + "v", y).reshape(B, M, heads, d).transpose(1, 2)
    if pre + "q_norm.weight" in p:
        q = OF.layer_norm(p, pre + "q_norm", q)
        k = OF.layer_norm(p, pre + "k_norm", k)
    if rope_base is not None:
        q = OF.rope_apply_1d(q, pos_q, rope_base)
        k = OF.rope_apply_1d(k, pos_k, rope_base)
    att = torch.softmax((q * d ** -0.5) @ k.transpose(-2, -1), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return OF.linear(p, pre + "proj", o)


def cross_block(p: Params, pre: str, x: torch.Tensor, y: torch.Tensor, pos, heads: int,
                rope_base: Optional[float] = 100.0) -> torch.Tensor:
    """CrossAttentionBlock.forward — cross_attention.py:126-131."""
    a = cross_attention(p, pre + "attn.", OF.layer_norm(p, pre + "norm1", x), OF.layer_norm(p, pre + "norm3", y),
                        pos[0], pos[1], heads, rope_base)
    x = x + OF.layer_scale(p, pre + "ls1", a)
    m = OF.mlp(p, pre + "mlp.", OF.layer_norm(p, pre + "norm2", x))
    return x + OF.layer_scale(p, pre + "ls2", m)


def gated_update(p: Params, pre: str, memory: torch.Tensor, update: torch.Tensor, num_tokens: int = 8) -> torch.Tensor:
    """GatedUpdate.forward — aligned_vggt/layers/gated_update.py:43-78.  memory (B,N,D) unit rows,
    update (B,1,D)."""
    B, N, D = memory.shape
    assert N == num_tokens
    u_norm = update.norm(dim=-1, keepdim=True)
    upd = update.expand_as(memory)
    mem_scaled = memory * u_norm
    mean_scaled = memory.mean(dim=1, keepdim=True).expand_as(memory) * u_norm
    inp = torch.cat([upd, mem_scaled, mean_scaled], dim=-1)
    deltas = []
    for i in range(N):
        h = F.gelu(OF.linear(p, f"{pre}delta_mlps.{i}.0", inp[:, i]))
        deltas.append(OF.linear(p, f"{pre}delta_mlps.{i}.2", h))
    diff = torch.stack(deltas, dim=1) - memory
    g_in = torch.cat([diff, mem_scaled], dim=-1)
    gate = torch.sigmoid(OF.linear(p, pre + "gate_mlp.2", F.gelu(OF.linear(p, pre + "gate_mlp.0", g_in))))
    orth = diff - (diff * memory).sum(-1, keepdim=True) * memory
    return F.normalize(memory + gate * F.normalize(orth, dim=-1), dim=-1)


# =============================================================================================
# alignment head
# =============================================================================================
def decode_alignments(p: Params, pre: str, align_tok: torch.Tensor, first_chunk: bool,
                      memory_tokens: Optional[torch.Tensor], heads: int = 8, depth_decoder: int = 2,
                      num_memory_tokens: int = 8, rope_base: Optional[float] = 100.0):
    """AlignmentHead._decode_alignments (eval path) — alignment_head.py:427-540.
    align_tok (B,S,1024) -> chunk_sim3 (B,1,8), frame_se3 (B,S-1,7), memory (B,8,512)."""
    B, S, _ = align_tok.shape
    dev = align_tok.device
    nm = num_memory_tokens
    # position ids (:446-459)
    q_ids_frame = torch.arange(1, S, device=dev).view(1, S - 1).expand(B, -1)
    k_ids_frame = torch.zeros(1, 1, dtype=q_ids_frame.dtype, device=dev).expand(B, -1)
    if nm > 0:
        k_ids_chunk = torch.arange(0, S + nm, device=dev)
        k_ids_chunk[-nm:] += S
    else:
        k_ids_chunk = torch.arange(0, S, device=dev)
    q_ids_chunk = torch.zeros(1, 1, dtype=k_ids_chunk.dtype, device=dev).expand(B, -1)
    k_ids_chunk = k_ids_chunk.view(1, -1).expand(B, -1)

    tokens = OF.layer_norm(p, pre + "dec_norm", OF.linear(p, pre + "project_dec", align_tok))  # :463-465
    C = tokens.shape[-1]
    directional = None
    if nm > 0:
        mean_norm = tokens.norm(dim=-1).mean(dim=-1, keepdim=True).unsqueeze(1)  # (B,1,1) :469
        if memory_tokens is None:
            mem = p[pre + "memory_token"].expand(B, -1, -1)
            init = OF.linear(p, pre + "frame_proj", tokens[:, 0]).view(B, -1, C)
            init_dir = init / init.norm(dim=-1, keepdim=True).clamp_min(1e-6)
            a = torch.sigmoid(p[pre + "alpha"])
            directional = (1 - a) * mem + a * init_dir          # not re-normalised (:478)
            effective = mem * mean_norm                         # un-blended memory (:479)
        else:
            directional = memory_tokens
            effective = memory_tokens * mean_norm
        assert directional.shape[0] == B
        kv = torch.cat([tokens, effective], dim=1)
    else:
        kv = tokens

    chunk_tok = tokens[:, :1]
    for i in range(depth_decoder):  # :496-501
        chunk_tok = cross_block(p, f"{pre}chunk_cross_blocks.{i}.", chunk_tok, kv, (q_ids_chunk, k_ids_chunk), heads,
                                rope_base)
    new_memory = memory_tokens
    if nm > 0:
        new_memory = gated_update(p, pre + "gated_update.", directional, chunk_tok, nm)  # :506
    chunk_tok_n = OF.layer_norm(p, pre + "chunk_norm", chunk_tok)  # :507

    frame_tok = tokens[:, 1:]
    for i in range(depth_decoder):  # :525-530
        frame_tok = cross_block(p, f"{pre}frame_cross_blocks.{i}.", frame_tok, chunk_tok_n,
                                (q_ids_frame, k_ids_frame), heads, rope_base)
    frame_tok = OF.layer_norm(p, pre + "frame_norm", frame_tok)
    frame_se3 = OF.mlp(p, pre + "frame_se3_decoder.", frame_tok)        # (B,S-1,7)
    chunk_sim3 = OF.mlp(p, pre + "chunk_sim3_decoder.", chunk_tok_n)    # (B,1,8)
    chunk_sim3 = torch.cat([chunk_sim3[..., :7], torch.exp(chunk_sim3[..., 7:])], dim=-1)  # :538
    return chunk_sim3, frame_se3, new_memory


def alignment_head_forward(p: Params, pre: str, tokens: torch.Tensor, image_size: Tuple[int, int],
                           next_num_overlap: int, overlap_tokens: Optional[torch.Tensor] = None,
                           memory_tokens: Optional[torch.Tensor] = None, *, patch_size: int = 14, depth_aa: int = 4,
                           heads: int = 8, num_register_tokens: int = 4, num_memory_tokens: int = 8,
                           temporal_attention: bool = True, rope_base: float = 100.0, return_tokens: bool = False,
                           amp: bool = False):
    """AlignmentHead.forward (eval) — alignment_head.py:224-345.
    tokens (B,S,P,2048) -> chunk_sim3 (B,1,8), frame_se3 (B,S-1,7), memory (B,8,512), overlap (B,1+o,P+1,1024).
    amp=True emulates the reference's shipping precision (Lightning bf16-mixed, run_model.py:472): the token blocks
    run under bf16 autocast, the decode with autocast disabled (alignment_head.py:340)."""
    with torch.autocast(tokens.device.type, dtype=torch.bfloat16, enabled=amp):
        x, T, first_chunk, S, P1, C, B = _head_blocks(p, pre, tokens, image_size, overlap_tokens, patch_size, depth_aa, heads,
                                                      num_register_tokens, temporal_attention, rope_base)
    x = x.float()
    with torch.autocast(tokens.device.type, enabled=False):
        chunk_sim3, frame_se3, memory = decode_alignments(p, pre, x[:, :, 0], first_chunk, memory_tokens, heads,
                                                          num_memory_tokens=num_memory_tokens, rope_base=rope_base)
    new_overlap = torch.cat([x[:, :1], x[:, S - next_num_overlap:]], dim=1).contiguous()  # :343
    if return_tokens:
        return chunk_sim3, frame_se3, memory, new_overlap, x
    return chunk_sim3, frame_se3, memory, new_overlap


def _head_blocks(p, pre, tokens, image_size, overlap_tokens, patch_size, depth_aa, heads, num_register_tokens,
                 temporal_attention, rope_base):
    """project_in .. the 4 x (frame block, temporal cross block) loop of AlignmentHead.forward (:242-337)."""
    H, W = image_size
    x = OF.layer_norm(p, pre + "token_norm", OF.linear(p, pre + "project_in", tokens))  # :242-247
    B, S, P, C = x.shape
    T = None
    if overlap_tokens is not None:
        assert overlap_tokens.shape[0] == B and overlap_tokens.shape[2] == 1 + P and overlap_tokens.shape[3] == C, \
            "Size of tokens and overlap tokens must match"
        T = overlap_tokens.shape[1]
    first_chunk = overlap_tokens is None
    x = torch.cat([OF.expand_special(p[pre + "per_frame_alignment_token"], B, S), x], dim=2)  # :269-270
    P1 = P + 1
    n_special = 2 + num_register_tokens  # alignment + camera + register tokens (:93)
    gh, gw = H // patch_size, W // patch_size
    dev = x.device

    pos2d = OF.token_positions(B * S, gh, gw, n_special, dev)  # :301-310
    ids = torch.arange(S, device=dev)
    if temporal_attention:  # :278-285
        if T is not None:
            q_ids = (ids + (S - (T - 1))).view(1, S).expand(B * P1, -1)
            k_ids = torch.cat([ids[:1], ids[-(T - 1):]]).view(1, T).expand(B * P1, -1)
        else:
            q_ids = ids.view(1, S).expand(B * P1, -1)
            k_ids = q_ids
        pos_t = (q_ids, k_ids)
    else:
        # The reference's temporal_attention=False variant cannot run: __init__ stores self.aa_order before
        # rebinding the local aa_order (alignment_head.py:80,146), so forward asks for self.temporal_blocks
        # which was never created -> AttributeError (verified by oracle/make_golden.py).  Not restated.
        raise AttributeError("'AlignmentHead' object has no attribute 'temporal_blocks'")

    for i in range(depth_aa):  # :317-335
        x = OF.block(p, f"{pre}frame_blocks.{i}.", x.reshape(B * S, P1, C), heads, pos2d, rope_base)
        # RAW view (B,S,P1,C)->(B*P1,S,C): groups of S consecutive flat tokens, not a transpose (:372-377)
        xt = x.reshape(B, S * P1 * C).view(B * P1, S, C)
        yt = xt if first_chunk else overlap_tokens.reshape(B, T * P1 * C).view(B * P1, T, C)
        x = cross_block(p, f"{pre}temporal_blocks.{i}.", xt, yt, pos_t, heads, rope_base)

    return x.reshape(B, S, P1, C), T, first_chunk, S, P1, C, B


# =============================================================================================
# pose helpers
# =============================================================================================
def extri_to_pose_encoding(extr: torch.Tensor) -> torch.Tensor:
    """aligned_vggt/utils/data.py:12-30.  (B,S,>=3,4) -> (B,S,7) [t, quat xyzw]."""
    q = OF.mat_to_quat(extr[:, :, :3, :3])
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    return torch.cat([extr[:, :, :3, 3], q], dim=-1).float()


def pose_encoding_to_extri(enc: torch.Tensor) -> torch.Tensor:
    """aligned_vggt/utils/data.py:33-52.  (B,S,>=7) -> (B,S,4,4); reads [:3] and [3:7] only."""
    q = enc[..., 3:7]
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    R = OF.quat_to_mat(q)
    top = torch.cat([R, enc[..., :3, None]], dim=-1)
    bottom = torch.zeros_like(top[..., :1, :])
    bottom[..., 0, 3] = 1.0
    return torch.cat([top, bottom], dim=-2)


def average_pose_encodings(enc: torch.Tensor) -> torch.Tensor:
    """aligned_vggt/utils/geometry.py:4-37.  Markley quaternion mean (largest eigenvector of
    sum q q^T / N) + mean translation.  (B,N,7) -> (B,1,7).  Sign of the eigenvector is the one
    torch.linalg.eigh returns (a rotation is sign-invariant)."""
    B, N, _ = enc.shape
    t = enc[..., :3].mean(dim=1, keepdim=True)
    q = enc[..., 3:7]
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    M = (q.unsqueeze(-1) * q.unsqueeze(-2)).sum(dim=1) / N
    _, vec = torch.linalg.eigh(M)
    qm = vec[..., -1]
    qm = qm / qm.norm(dim=-1, keepdim=True)
    return torch.cat([t, qm.unsqueeze(1)], dim=-1).float()


def inv_se3(m: torch.Tensor) -> torch.Tensor:
    """closed-form SE(3) inverse over arbitrary leading dims (upstream closed_form_inverse_se3)."""
    sh = m.shape
    return OF.closed_form_inverse_se3(m.reshape(-1, sh[-2], sh[-1])).reshape(sh[:-2] + (4, 4))


# =============================================================================================
# Sim(3) apply
# =============================================================================================
def apply_sim3_points(points: torch.Tensor, T: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """apply_sim3_alignment_on_point_maps — aligned_vggt/utils/alignment.py:491-526.
    p' = T[:3,:3] (s p) + T[:3,3] per batch element.  points (B,S,H,W,3) or (S,H,W,3)."""
    if points.dim() == 4:
        points, T, s = points.unsqueeze(0), T.unsqueeze(0), s.unsqueeze(0)
    assert points.shape[0] == T.shape[0] == s.shape[0], "Inputs must have matching batch dimension"
    B = points.shape[0]
    sp = points * s.view(B, 1, 1, 1, 1)
    R = T[:, :3, :3].view(B, 1, 1, 1, 3, 3)
    t = T[:, :3, 3].view(B, 1, 1, 1, 3)
    return (R * sp.unsqueeze(-2)).sum(-1) + t


def apply_sim3_c2w(poses: torch.Tensor, T: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """apply_sim3_alignment_on_c2w — alignment.py:558-594.  poses (B,S,4,4) -> T @ [R, s t]."""
    if poses.dim() == 3:
        poses, T, s = poses.unsqueeze(0), T.unsqueeze(0), s.unsqueeze(0)
    assert poses.shape[0] == T.shape[0] == s.shape[0], "Inputs must have matching batch dimension"
    B = poses.shape[0]
    scaled = poses.clone()
    scaled[:, :, :3, 3] = scaled[:, :, :3, 3] * s.view(B, 1, 1)
    return T.unsqueeze(1) @ scaled


def apply_sim3_w2c(extr: torch.Tensor, T: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """apply_sim3_alignment_on_w2c — alignment.py:528-556.  extr (B,S,3|4,4) -> (B,S,4,4)."""
    if extr.dim() == 3:
        extr, T, s = extr.unsqueeze(0), T.unsqueeze(0), s.unsqueeze(0)
    assert extr.shape[0] == T.shape[0] == s.shape[0], "Inputs must have matching batch dimension"
    return inv_se3(apply_sim3_c2w(inv_se3(extr), T, s))


# =============================================================================================
# IRLS weighted Umeyama (point-aligned baseline)
# =============================================================================================
def weighted_umeyama(src: torch.Tensor, dst: torch.Tensor, w: torch.Tensor):
    """weighted_umeyama_sim3 — aligned_vggt/models/pointAligned_wrapped_vggt.py:159-217."""
    assert src.ndim == 2 and src.shape[1] == 3 and dst.shape == src.shape
    W = w.sum()
    if W < 1e-6:
        raise ValueError("Total weight too small for meaningful estimation")
    wc = w.view(-1, 1)
    mu_x = (wc * src).sum(0) / W
    mu_y = (wc * dst).sum(0) / W
    xc, yc = src - mu_x, dst - mu_y
    Sigma = (wc * yc).T @ xc / W
    U, Sv, Vh = torch.linalg.svd(Sigma, full_matrices=True)
    sgn = torch.sign(torch.det(U @ Vh))
    D = torch.stack([torch.ones_like(sgn), torch.ones_like(sgn), sgn])
    R = U @ torch.diag(D) @ Vh
    var_x = (w * (xc ** 2).sum(1)).sum() / W
    s = (Sv * D).sum() / var_x
    t = mu_y - s * (R @ mu_x)
    return R, t, s


def irls_umeyama(src, dst, conf_src, conf_dst, conf_threshold_factor=0.5, delta=0.1, max_iters=20, tol=1e-9):
    """irls_sim3_umeyama — pointAligned_wrapped_vggt.py:219-305."""
    assert src.shape[0] == dst.shape[0]
    src, dst = src.reshape(-1, 3), dst.reshape(-1, 3)
    comb = torch.sqrt(conf_src.reshape(-1) * conf_dst.reshape(-1))
    keep = comb >= conf_threshold_factor * torch.median(comb)
    src, dst, comb = src[keep], dst[keep], comb[keep]
    R, t, s = weighted_umeyama(src, dst, comb.clone())
    for _ in range(max_iters):
        res = torch.linalg.norm(s * (src @ R.T) + t - dst, dim=1)
        rw = torch.where(res <= delta, torch.ones_like(res), delta / res.clamp_min(1e-12))
        R2, t2, s2 = weighted_umeyama(src, dst, comb * rw)
        dR, dt, ds = torch.norm(R2 - R), torch.norm(t2 - t), torch.abs(s2 - s)
        R, t, s = R2, t2, s2
        if dR < tol and dt < tol and ds < tol:
            break
    return R, t, s


# =============================================================================================
# chunk scheduler
# =============================================================================================
def generate_chunks(num_frames: int, mode: str, seq_width: int, overlap: int) -> List[List[int]]:
    """generate_chunks — aligned_vggt/utils/data.py:155-207 (deterministic modes)."""
    out: List[List[int]] = []
    if mode == "chunk_gt":
        for i in range(0, num_frames - seq_width + 1, seq_width):
            out.append(list(range(i, i + seq_width)))
        if len(out) * seq_width < num_frames:
            out.append(list(range(len(out) * seq_width, num_frames)))
    elif mode == "chunk_overlap":
        if num_frames < seq_width:
            out.append(list(range(num_frames)))
        else:
            step = seq_width - overlap
            for i in range(0, num_frames - seq_width + 1, step):
                out.append(list(range(i, i + seq_width)))
            if len(out) * step < num_frames - overlap:
                out.append(list(range(len(out) * step, num_frames)))
    elif mode == "all":
        out = [list(range(num_frames))]
    else:
        raise ValueError(f"Unknown sequence generation mode: {mode}")
    return out


# =============================================================================================
# model-level forward (FeatureAlignedVGGT) with the decoder heads replaced by given raw maps
# =============================================================================================
def compose_alignment(chunk_sim3: torch.Tensor, frame_se3: torch.Tensor):
    """featureAligned_vggt.py:97-101.  -> per_frame_se3 (B,S,4,4), scale (B,1)."""
    chunk_se3 = pose_encoding_to_extri(chunk_sim3)
    per_frame = pose_encoding_to_extri(frame_se3) @ chunk_se3
    return torch.cat([chunk_se3, per_frame], dim=1), chunk_sim3[..., -1]


def scale_lse(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """scale_lse_solver — /root/reference/aligned_vggt/utils/alignment.py:113-129: |sum(x*y) / sum(x**2)|, (n,3) each."""
    assert x.shape == y.shape, "x shape not equal to y shape"
    return ((x * y).sum() / (x ** 2).sum()).abs()


def pose_chain(pose_enc_cam: torch.Tensor, image_hw, per_frame_se3: torch.Tensor, scale: torch.Tensor,
               prev_pose_enc: Optional[torch.Tensor], overlap: int, gt_mean: Optional[torch.Tensor] = None,
               gt_scale_poses: Optional[torch.Tensor] = None):
    """featureAligned_vggt.py:106-143.  Returns (aligned_pose_enc (B,S,9), per_frame_se3 incl. mean transform (B,S,4,4),
    point_identity_alignment (B,4,4)).  gt_mean (B,1,4,4): `mean_camera_transform = gt_poses[:,:1]` (:123-124, only looked at
    with a previous chunk).  gt_scale_poses (B,S,4,4) padded gt: the pose-aligned baseline's position-scale alignment
    (poseAligned_wrapped_vggt.py:84-104); the per-batch scales are then returned as a fourth value."""
    B = pose_enc_cam.shape[0]
    extr3, intr = OF.pose_encoding_to_extri_intri(pose_enc_cam, image_size_hw=image_hw)
    extr = F.pad(extr3, (0, 0, 0, 1))
    extr[:, :, 3, 3] = 1.0
    ident = OF.closed_form_inverse_se3(extr[:, 0])
    point_identity = extr[:, 0].clone()
    extr = extr @ ident.view(B, 1, 4, 4)
    extr[:, :, :3, 3] *= scale.view(B, 1, 1)
    batch_scales = None
    if gt_scale_poses is not None and extr.shape[1] > 1:
        centred = gt_scale_poses @ OF.closed_form_inverse_se3(gt_scale_poses[:, 0]).view(B, 1, 4, 4)
        batch_scales = torch.stack([scale_lse(extr[b, :, :3, 3], centred[b, :, :3, 3]) for b in range(B)])
        extr[:, :, :3, 3] *= batch_scales.view(B, 1, 1)
    if prev_pose_enc is not None and gt_mean is not None:
        mean_T = gt_mean.to(extr)
    elif prev_pose_enc is not None:
        ctx = pose_encoding_to_extri(prev_pose_enc[:, -overlap:])
        cam_T = inv_se3(extr[:, :overlap]) @ ctx
        if overlap > 1:
            mean_T = pose_encoding_to_extri(average_pose_encodings(extri_to_pose_encoding(cam_T)))
        else:
            mean_T = cam_T
    else:
        mean_T = torch.eye(4, dtype=extr.dtype, device=extr.device).view(1, 1, 4, 4).expand(B, -1, -1, -1)
    per_frame_se3 = per_frame_se3 @ mean_T
    aligned = extr @ per_frame_se3
    enc = OF.extri_intri_to_pose_encoding(aligned, intr, image_size_hw=image_hw)
    if gt_scale_poses is not None:
        return enc, per_frame_se3, point_identity, batch_scales
    return enc, per_frame_se3, point_identity


def point_transform(per_frame_se3: torch.Tensor, point_identity: torch.Tensor, has_context: bool) -> torch.Tensor:
    """featureAligned_vggt.py:190-196 -> (B,4,4) applied to scale*points."""
    if has_context:
        return inv_se3(per_frame_se3[:, 0]) @ point_identity
    return point_identity


def feature_aligned_forward(p: Params, images: torch.Tensor, num_overlap: int, context: Optional[dict] = None, *,
                            raw_points: Optional[torch.Tensor] = None, raw_depth: Optional[torch.Tensor] = None,
                            depth: int = 24, dino_depth: int = 24, taps=(4, 11, 17, 23), depth_aa: int = 4,
                            num_memory_tokens: int = 8, amp: bool = False, gt_poses: Optional[torch.Tensor] = None) -> dict:
    """FeatureAlignedVGGT.forward (eval, enable_camera=True) — featureAligned_vggt.py:48-225.
    The DPT heads are out of scope: ``raw_points`` (B,S,H,W,3) / ``raw_depth`` (B,S,H,W,1)
    stand in for their outputs so the Sim(3) application can be checked.  Returns this chunk's tensors
    (not the accumulated lists) plus the context entries needed by the next chunk."""
    B, S, _, H, W = images.shape
    with torch.autocast(images.device.type, dtype=torch.bfloat16, enabled=amp):  # amp: the reference's bf16-mixed inference precision
        toks, _ = OF.aggregator_forward(p, "aggregator.", images, depth=depth, dino_depth=dino_depth, keep=taps)
    taps_t = [toks[i].float() for i in taps]
    return _feature_aligned_tail(p, images, taps_t, num_overlap, context, raw_points, raw_depth, depth_aa, num_memory_tokens, amp, gt_poses)


def _feature_aligned_tail(p, images, taps_t, num_overlap, context, raw_points, raw_depth, depth_aa, num_memory_tokens, amp, gt_poses):
    """Everything of FeatureAlignedVGGT.forward after the Aggregator (featureAligned_vggt.py:84-225)."""
    B, S, _, H, W = images.shape
    ctx_overlap = ctx_mem = prev_pose = None
    if context is not None:
        ctx_overlap = context["overlap_tokens"]
        ctx_mem = context["memory_tokens"] if num_memory_tokens > 0 else None
        prev_pose = context["pose_enc"]
    overlap = num_overlap if S > num_overlap else S - 1  # :93
    chunk_sim3, frame_se3, memory, overlap_tokens = alignment_head_forward(
        p, "alignment_head.", taps_t[-1], (H, W), overlap, ctx_overlap, ctx_mem, depth_aa=depth_aa,
        num_memory_tokens=num_memory_tokens, amp=amp)
    per_frame, scale = compose_alignment(chunk_sim3, frame_se3)
    cam_enc = OF.camera_head_forward(p, "camera_head.", taps_t[-1])[-1]
    pose_enc, per_frame, pt_ident = pose_chain(cam_enc, (H, W), per_frame, scale, prev_pose, overlap,
                                               gt_mean=None if gt_poses is None else gt_poses[:, :1])
    out = {"taps": taps_t, "chunk_sim3_alignment_enc": chunk_sim3, "frame_se3_alignment_enc": frame_se3,
           "memory_tokens": memory, "overlap_tokens": overlap_tokens, "pose_enc": pose_enc, "camera_pose_enc": cam_enc}
    if raw_depth is not None:
        out["depth"] = raw_depth * scale.view(B, 1, 1, 1, 1)  # :171
    if raw_points is not None:
        Tp = point_transform(per_frame, pt_ident, context is not None)
        out["world_points"] = apply_sim3_points(raw_points, Tp, scale.view(B))  # :198-207
        out["point_transform"] = Tp
    return out


def feature_aligned_forward_sliced(p: Params, images: torch.Tensor, num_overlap: int, context: Optional[dict] = None, *,
                                   raw_points: Optional[torch.Tensor] = None, raw_depth: Optional[torch.Tensor] = None,
                                   depth: int = 24, dino_depth: int = 24, taps=(4, 11, 17, 23), depth_aa: int = 4,
                                   num_memory_tokens: int = 8):
    """The same computation as feature_aligned_forward (fp32), as a generator that yields (unit name, relative cost) after each
    unit of work — patch embedding, every DINO block, every frame / global block, the tail (heads, pose chain, Sim(3) apply) —
    and returns the output dict (StopIteration.value).  bench.py's CPU reference arm times a bounded slice of a full-size chunk
    per step with it; tests/test_oracle_golden.py checks it against the monolithic function."""
    B, S, _, H, W = images.shape
    pre, n_reg, heads, patch = "aggregator.", 4, 16, 14
    mean = torch.tensor(OF.RESNET_MEAN, dtype=images.dtype, device=images.device).view(1, 1, 3, 1, 1)
    std = torch.tensor(OF.RESNET_STD, dtype=images.dtype, device=images.device).view(1, 1, 3, 1, 1)
    x = ((images - mean) / std).view(B * S, 3, H, W)
    dp = pre + "patch_embed."
    gh, gw = H // patch, W // patch
    x = torch.nn.functional.conv2d(x, p[dp + "patch_embed.proj.weight"], p[dp + "patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    x = torch.cat([p[dp + "cls_token"].expand(B * S, -1, -1), x], dim=1) + OF.interpolate_pos_embed(p[dp + "pos_embed"], gh, gw)
    x = torch.cat([x[:, :1], p[dp + "register_tokens"].expand(B * S, -1, -1), x[:, 1:]], dim=1)
    yield "patch_embed", 0.3
    for i in range(dino_depth):
        x = OF.block(p, f"{dp}blocks.{i}.", x, heads, ln_eps=1e-6)
        yield f"dino.{i}", 1.0
    patch_tok = OF.layer_norm(p, dp + "norm", x, 1e-6)[:, 1 + n_reg:]
    C = patch_tok.shape[-1]
    cam = OF.expand_special(p[pre + "camera_token"], B, S).reshape(B * S, 1, C)
    reg = OF.expand_special(p[pre + "register_token"], B, S).reshape(B * S, n_reg, C)
    tokens = torch.cat([cam, reg, patch_tok], dim=1)
    pos = OF.token_positions(B * S, gh, gw, 1 + n_reg, images.device)
    P = tokens.shape[1]
    global_cost = 1.0 + 2.0 * (S * P) / 13184.0   # attention over all S*P tokens: twice the block's GEMM work at S*P = 13184
    kept = {}
    for i in range(depth):
        tokens = OF.block(p, f"{pre}frame_blocks.{i}.", tokens.view(B * S, P, C), heads, pos, 100.0)
        frame_out = tokens.view(B, S, P, C)
        yield f"frame.{i}", 1.0
        tokens = OF.block(p, f"{pre}global_blocks.{i}.", tokens.view(B, S * P, C), heads, pos.view(B, S * P, 2), 100.0)
        if i in taps:
            kept[i] = torch.cat([frame_out, tokens.view(B, S, P, C)], dim=-1)
        yield f"global.{i}", global_cost
    out = _feature_aligned_tail(p, images, [kept[i].float() for i in taps], num_overlap, context, raw_points, raw_depth, depth_aa,
                                num_memory_tokens, False, None)
    yield "tail", 8.0
    return out


def pose_aligned_forward(p: Params, images: torch.Tensor, num_overlap: int, context: Optional[dict] = None, *,
                         raw_points: Optional[torch.Tensor] = None, raw_depth: Optional[torch.Tensor] = None, depth: int = 24,
                         dino_depth: int = 24, taps=(4, 11, 17, 23), gt_poses: Optional[torch.Tensor] = None) -> dict:
    """pose-aligned baseline VGGT.forward (eval) — aligned_vggt/models/poseAligned_wrapped_vggt.py:36-204.
    Same chain as the feature-aligned model with identity learned alignment and unit scale (:107-130, :171-187);
    gt_poses (B,S,3,4): position-scale alignment + gt chunk transform (:84-109), scale applied to depth / points (:144-147, :166-169).
    raw_points / raw_depth stand in for (or are) the DPT head outputs."""
    B, S, _, H, W = images.shape
    toks, _ = OF.aggregator_forward(p, "aggregator.", images, depth=depth, dino_depth=dino_depth, keep=taps)
    cam_enc = OF.camera_head_forward(p, "camera_head.", toks[taps[-1]])[-1]
    eye = torch.eye(4).view(1, 1, 4, 4).expand(B, S, -1, -1)
    prev = context["pose_enc"] if context is not None else None
    batch_scales = None
    if gt_poses is not None:
        gt4 = F.pad(gt_poses, (0, 0, 0, 1))
        gt4[:, :, 3, 3] = 1.0
        pose_enc, per_frame, pt_ident, batch_scales = pose_chain(cam_enc, (H, W), eye, torch.ones(B, 1), prev, num_overlap,
                                                                 gt_mean=gt4[:, :1], gt_scale_poses=gt4)
    else:
        pose_enc, per_frame, pt_ident = pose_chain(cam_enc, (H, W), eye, torch.ones(B, 1), prev, num_overlap)
    s = torch.ones(B) if batch_scales is None else batch_scales
    out = {"pose_enc": pose_enc, "camera_pose_enc": cam_enc, "taps": [toks[i].float() for i in taps], "batch_scales": s}
    if raw_depth is not None:
        out["depth"] = raw_depth * s.view(B, 1, 1, 1, 1)
    if raw_points is not None:
        Tp = point_transform(per_frame, pt_ident, context is not None)
        out["world_points"] = apply_sim3_points(raw_points, Tp, s)
        out["point_transform"] = Tp
    return out


def point_aligned_forward(p: Params, images: torch.Tensor, num_overlap: int, context: Optional[dict] = None, *,
                          raw_points: torch.Tensor, raw_points_conf: torch.Tensor, raw_depth: Optional[torch.Tensor] = None,
                          depth: int = 24, dino_depth: int = 24, taps=(4, 11, 17, 23)) -> dict:
    """point-aligned baseline VGGT.forward (eval) — aligned_vggt/models/pointAligned_wrapped_vggt.py:34-157: IRLS Umeyama of
    the overlap point maps onto the previous chunk's aligned ones (:74-99), applied to points (:100), poses (:113-122) and
    depth (:135-138).  context: {"world_points", "world_points_conf"} of the previous chunk (aligned)."""
    B, S, _, H, W = images.shape
    toks, _ = OF.aggregator_forward(p, "aggregator.", images, depth=depth, dino_depth=dino_depth, keep=taps)
    cam_enc = OF.camera_head_forward(p, "camera_head.", toks[taps[-1]])[-1]
    T = torch.eye(4).repeat(B, 1, 1)
    s = torch.ones(B)
    if context is not None:
        for b in range(B):
            r, t, sc = irls_umeyama(raw_points[b, :num_overlap], context["world_points"][b, -num_overlap:],
                                    raw_points_conf[b, :num_overlap], context["world_points_conf"][b, -num_overlap:])
            T[b, :3, :3], T[b, :3, 3], s[b] = r, t, sc
    out = {"world_points": apply_sim3_points(raw_points, T, s), "world_points_conf": raw_points_conf,
           "pose_enc": pose_enc_apply_sim3(cam_enc, (H, W), T, s), "transform": T, "scales": s}
    if raw_depth is not None:
        out["depth"] = raw_depth * s.view(B, 1, 1, 1, 1)
    return out


def pose_enc_apply_sim3(pose_enc: torch.Tensor, image_hw, T: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """pointAligned_wrapped_vggt.py:113-122: pose_enc -> extrinsics -> apply_sim3_alignment_on_w2c -> pose_enc."""
    extr, intr = OF.pose_encoding_to_extri_intri(pose_enc, image_size_hw=image_hw)
    aligned = apply_sim3_w2c(extr, T, s)
    return OF.extri_intri_to_pose_encoding(aligned, intr, image_size_hw=image_hw)


def apply_sim3_alignment(T: torch.Tensor, s: torch.Tensor, pose_enc: torch.Tensor, image_hw, points: Optional[torch.Tensor] = None,
                         depths: Optional[torch.Tensor] = None):
    """apply_sim3_alignment — /root/reference/aligned_vggt/utils/alignment.py:449-489 (out of place)."""
    B = T.shape[0]
    return (pose_enc_apply_sim3(pose_enc, image_hw, T, s),
            None if points is None else apply_sim3_points(points, T, s),
            None if depths is None else depths * s.view(B, 1, 1, 1, 1))


# ----------------------------------------------------------------------------- evaluation-side geometry
def unproject_depth(depth_map: torch.Tensor, extrinsics: torch.Tensor, intrinsics: torch.Tensor) -> torch.Tensor:
    """unproject_depth_map_to_point_map — /root/reference/aligned_vggt/utils/geometry.py:39-75.
    depth (B,S,H,W,1), extrinsics (B,S,3,4) w2c, intrinsics (B,S,3,3) -> (B,S,H,W,3)."""
    B, S, H, W, _ = depth_map.shape
    u, v = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")  # geometry.py:152-158
    pix = torch.stack((u, v, torch.ones_like(u)), dim=-1).float().view(-1, 3)
    rays = (torch.inverse(intrinsics) @ pix.t()[None, None]).permute(0, 1, 3, 2)  # (B,S,HW,3)
    cam = rays * depth_map.reshape(B, S, -1, 1)
    poses = OF.closed_form_inverse_se3(extrinsics.reshape(B * S, 3, 4)).reshape(B, S, 4, 4)
    world = cam @ poses[..., :3, :3].transpose(-1, -2) + poses[..., None, :3, 3]
    return world.view(B, S, H, W, 3)


def depth_scale_align(d_pred: torch.Tensor, d_gt: torch.Tensor, mask: torch.Tensor, conf: torch.Tensor) -> torch.Tensor:
    """the solver of scale_align_from_depths — /root/reference/aligned_vggt/utils/alignment.py:259-313: (B,N) each -> scales (B)."""
    x, y, m, w_conf = d_pred, d_gt, mask.float(), conf
    N = x.shape[1]
    sum_valid = m.sum(dim=-1, keepdim=True).clamp_min(1.0)
    min_depth = 0.1 * ((y * m).sum(dim=-1, keepdim=True) / sum_valid)
    w = m * w_conf * (1.0 / torch.max(y, min_depth).clamp_min(1e-6))
    sign = torch.sign(x)
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    x_pos, y_pos = x * sign, y * sign
    r = y_pos / x_pos.clamp_min(1e-6)
    w_eff = w * x_pos
    r_sorted, idx = torch.sort(r, dim=-1)
    cumsum = torch.gather(w_eff, -1, idx).cumsum(-1)
    idx_med = torch.searchsorted(cumsum, 0.5 * cumsum[:, -1:], side="left").clamp(max=N - 1)
    scales = torch.gather(r_sorted, -1, idx_med).squeeze(-1)
    scales[scales <= 0] *= -1
    return scales


def convert_dict_lists(chunked: dict, overlap: int) -> dict:
    """convertDictListsToTensors — /root/reference/aligned_vggt/utils/data.py:54-87 (out of place, tensors only)."""
    out = {}
    for k, items in chunked.items():
        if isinstance(items, list) and torch.is_tensor(items[0]):
            out[k] = torch.cat([t if i == 0 else t[:, overlap:] for i, t in enumerate(items)], dim=1)
    return out

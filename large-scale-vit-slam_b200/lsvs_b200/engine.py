"""ctypes wrapper of the native engine (csrc/engine.cu): parameter push + module forwards.
Everything here is plumbing: device memory comes from torch, the work runs in liblsvs_b200.so."""
import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import native as _n

_i, _f, _ll, _vp = ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_void_p


class EngineConfig(ctypes.Structure):
    """mirror of `lsvs_engine_config` in include/lsvs_b200.h"""
    _fields_ = [("embed_dim", _i), ("num_heads", _i), ("patch_size", _i), ("num_register_tokens", _i), ("depth", _i),
                ("dino_depth", _i), ("head_depth_aa", _i), ("num_memory_tokens", _i), ("with_alignment_head", _i),
                ("with_camera_head", _i), ("rope_base", _f), ("precision", _i)]


def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """DINOv2 interpolate_pos_encoding (UPSTREAM A.2): bicubic + antialias resize of the learned M x M table to
    (gh, gw); identity for the native square grid.  Weight preprocessing, runs once per image size."""
    n = pos_embed.shape[1] - 1
    m = int(math.sqrt(n))
    if gh == m and gw == m:
        return pos_embed[0].float().contiguous()
    pe = pos_embed.float()
    c = pe.shape[-1]
    patch = F.interpolate(pe[:, 1:].reshape(1, m, m, c).permute(0, 3, 1, 2), size=(gh, gw), mode="bicubic", antialias=True)
    patch = patch.permute(0, 2, 3, 1).reshape(gh * gw, c)
    return torch.cat([pe[0, :1], patch], dim=0).contiguous()


def _dpt_weight_layout(name: str, t: torch.Tensor) -> torch.Tensor:
    """torch conv layouts -> the (rows, cols) layouts of include/lsvs_b200.h (lsvs_dpt_head_forward)."""
    if not name.endswith((".weight", ".bias")) or ".norm." in name:
        return t
    if "resize_layers.0." in name or "resize_layers.1." in name:  # ConvTranspose2d (ic, oc, k, k) -> ((ky, kx, oc), ic)
        return t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]) if t.dim() == 4 else t
    if t.dim() == 4:  # Conv2d (oc, ic, k, k) -> (oc, (ky, kx, ic))
        t = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)
    if "output_conv2.0." in name:  # 32 output channels zero-padded to the 64-wide GEMM tile
        pad = torch.zeros((64 - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    return t


class Engine:
    """One native engine per (model, device).  `sync(named_params)` pushes changed parameters."""

    def __init__(self, depth=24, dino_depth=24, head_depth_aa=4, num_memory_tokens=8, with_alignment_head=True,
                 with_camera_head=True, rope_base=100.0, precision=None):
        """precision (include/lsvs_b200.h lsvs_engine_config::precision): 0 bf16 tensor-core operands (default; env
        LSVS_PRECISION overrides the default), 1 fp32-class alignment head + camera-head trunk, 2 fp32-class everywhere."""
        if precision is None:
            precision = int(os.environ.get("LSVS_PRECISION", "0"))
        if precision not in (0, 1, 2):
            raise ValueError(f"precision must be 0, 1 or 2, got {precision}")
        self.precision = precision
        self.cfg = EngineConfig(1024, 16, 14, 4, depth, dino_depth, head_depth_aa, num_memory_tokens,
                                int(with_alignment_head), int(with_camera_head), rope_base, precision)
        self._h = _vp()
        _n.check(_n.lib().lsvs_engine_create(ctypes.byref(self.cfg), ctypes.byref(self._h)), "engine_create")
        self._seen: Dict[str, Tuple[int, int]] = {}
        self._pos_key = None
        self.device = None

    def __del__(self):
        try:
            if self._h:
                _n.lib().lsvs_engine_destroy(self._h)
                self._h = _vp()
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def sync(self, named_params) -> int:
        """Push parameters whose storage or version changed since the last call; returns how many were pushed."""
        pushed = 0
        pos = None
        for name, p in named_params:
            if not p.is_cuda:
                raise _n.NativeError(f"parameter {name} lives on {p.device}: move the model to a CUDA device "
                                     "(there is no CPU fallback on this path)")
            if self.device is None:
                self.device = p.device
            if name.endswith("patch_embed.pos_embed"):
                pos = p
            key = (p.data_ptr(), p._version)
            if self._seen.get(name) == key:
                continue
            t = p.detach()
            if ".depth_head." in "." + name or ".point_head." in "." + name:
                t = _dpt_weight_layout(name, t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            rows, cols = (t.shape[0], t[0].numel()) if t.dim() >= 2 and name.endswith(".weight") else (0, 0)
            _n.check(_n.lib().lsvs_engine_set_param(self._h, name.encode(), _n.ptr(t), _ll(t.numel()), _i(rows), _i(cols),
                                                    _n.stream_ptr()), f"engine_set_param({name})")
            self._seen[name] = key
            pushed += 1
            if name.endswith("patch_embed.pos_embed"):
                self._pos_key = None
        if pushed:
            _n.check(_n.lib().lsvs_engine_finalize(self._h, _n.stream_ptr()), "engine_finalize")
        self._pos_param = pos if pos is not None else getattr(self, "_pos_param", None)
        return pushed

    def _ensure_pos(self, gh: int, gw: int):
        if self._pos_key == (gh, gw):
            return
        pe = interpolate_pos_embed(self._pos_param.detach(), gh, gw)
        _n.check(_n.lib().lsvs_engine_set_pos_embed(self._h, _n.ptr(pe), _i(gh), _i(gw), _n.stream_ptr()), "engine_set_pos_embed")
        torch.cuda.current_stream().synchronize()  # `pe` is a temporary; once per image size
        self._pos_key = (gh, gw)

    # ------------------------------------------------------------------ forwards
    def aggregator_forward(self, images: torch.Tensor, tap_layers: Sequence[int]) -> List[torch.Tensor]:
        B, S, C, H, W = images.shape
        assert C == 3
        img = images.detach()
        if img.dtype != torch.float32 or not img.is_contiguous():
            img = img.float().contiguous()
        gh, gw = H // 14, W // 14
        self._ensure_pos(gh, gw)
        P = 5 + gh * gw
        uniq = sorted(set(int(t) for t in tap_layers))
        bufs = {t: torch.empty(B, S, P, 2048, dtype=torch.float32, device=img.device) for t in uniq}
        ptrs = (_vp * len(uniq))(*[bufs[t].data_ptr() for t in uniq])
        ids = (_i * len(uniq))(*uniq)
        _n.check(_n.lib().lsvs_aggregator_forward(self._h, _n.ptr(img), _i(B), _i(S), _i(H), _i(W), ptrs, ids, _i(len(uniq)),
                                                  _n.stream_ptr()), "aggregator_forward")
        return [bufs[int(t)] for t in tap_layers]

    def alignment_head_forward(self, tokens: torch.Tensor, image_size, next_overlap: int,
                               overlap_tokens: Optional[torch.Tensor], memory_tokens: Optional[torch.Tensor]):
        B, S, P, C = tokens.shape
        H, W = image_size
        dev = tokens.device
        as_bf16 = tokens.dtype == torch.bfloat16 and self.precision == 0   # shipped by the chunk scheduler: used as they are
        tok = tokens.detach().contiguous() if as_bf16 else tokens.detach().float().contiguous()
        T = 0
        ov = mem = None
        if overlap_tokens is not None:
            assert overlap_tokens.shape[0] == B and overlap_tokens.shape[2] == 1 + P and overlap_tokens.shape[3] == 1024, \
                "Size of tokens and overlap tokens must match"
            T = overlap_tokens.shape[1]
            ov = overlap_tokens.detach().to(dev, torch.float32).contiguous()
        if memory_tokens is not None:
            assert memory_tokens.shape[0] == B, "Memory tokens must have same batch dimension as frame tokens"
            mem = memory_tokens.detach().to(dev, torch.float32).contiguous()
        sim3 = torch.empty(B, 1, 8, device=dev)
        se3 = torch.empty(B, max(S - 1, 0), 7, device=dev)
        nm = int(self.cfg.num_memory_tokens)
        mem_out = torch.empty(B, nm, 512, device=dev) if nm > 0 else None
        ov_out = torch.empty(B, 1 + next_overlap, P + 1, 1024, device=dev)
        fn = _n.lib().lsvs_alignment_head_forward_bf16 if as_bf16 else _n.lib().lsvs_alignment_head_forward
        _n.check(fn(self._h, _n.ptr(tok), _i(B), _i(S), _i(P), _i(H), _i(W), _i(next_overlap),
                                                      _n.ptr(ov), _i(T), _n.ptr(mem), _n.ptr(sim3), _n.ptr(se3), _n.ptr(mem_out),
                                                      _n.ptr(ov_out), _n.stream_ptr()), "alignment_head_forward")
        if nm == 0:  # the reference hands the caller's memory_tokens argument back untouched (alignment_head.py:504-506)
            mem_out = memory_tokens
        return sim3, se3, mem_out, ov_out

    def alignment_head_prefix(self, tokens: torch.Tensor, image_size) -> torch.Tensor:
        """The context-free part of the head (project_in, token_norm, alignment token, first frame block): tokens (B,S,P,2048) ->
        fp32 token stream (B,S,P+1,1024).  Runs on the rank that encoded the chunk (lsvs_alignment_head_prefix)."""
        B, S, P, C = tokens.shape
        H, W = image_size
        tok = tokens.detach().float().contiguous()
        out = torch.empty(B, S, P + 1, 1024, dtype=torch.float32, device=tok.device)
        _n.check(_n.lib().lsvs_alignment_head_prefix(self._h, _n.ptr(tok), _i(B), _i(S), _i(P), _i(H), _i(W), _n.ptr(out), _n.stream_ptr()),
                 "alignment_head_prefix")
        return out

    def alignment_head_resume(self, prefix: torch.Tensor, image_size, next_overlap: int, overlap_tokens: Optional[torch.Tensor],
                              memory_tokens: Optional[torch.Tensor]):
        """The rest of the head from a prefix stream (B,S,P+1,1024): same outputs as alignment_head_forward, bit for bit."""
        B, S, P1, C = prefix.shape
        assert C == 1024
        P = P1 - 1
        H, W = image_size
        dev = prefix.device
        x0 = prefix.detach().float().contiguous()
        T = 0
        ov = mem = None
        if overlap_tokens is not None:
            assert overlap_tokens.shape[0] == B and overlap_tokens.shape[2] == P1 and overlap_tokens.shape[3] == 1024, \
                "Size of tokens and overlap tokens must match"
            T = overlap_tokens.shape[1]
            ov = overlap_tokens.detach().to(dev, torch.float32).contiguous()
        if memory_tokens is not None:
            mem = memory_tokens.detach().to(dev, torch.float32).contiguous()
        nm = int(self.cfg.num_memory_tokens)
        sim3 = torch.empty(B, 1, 8, device=dev)
        se3 = torch.empty(B, max(S - 1, 0), 7, device=dev)
        mem_out = torch.empty(B, nm, 512, device=dev) if nm > 0 else None
        ov_out = torch.empty(B, 1 + next_overlap, P1, 1024, device=dev)
        _n.check(_n.lib().lsvs_alignment_head_resume(self._h, _n.ptr(x0), _i(B), _i(S), _i(P), _i(H), _i(W), _i(next_overlap), _n.ptr(ov),
                                                     _i(T), _n.ptr(mem), _n.ptr(sim3), _n.ptr(se3), _n.ptr(mem_out), _n.ptr(ov_out),
                                                     _n.stream_ptr()), "alignment_head_resume")
        return sim3, se3, (mem_out if nm > 0 else memory_tokens), ov_out

    def alignment_decode_forward(self, align_tokens: torch.Tensor, memory_tokens: Optional[torch.Tensor]):
        """fp32 decode stage alone (alignment_head.py:427-540): (B,S,1024) -> sim3 (B,1,8), se3 (B,S-1,7), memory (B,8,512)."""
        B, S, C = align_tokens.shape
        assert C == 1024
        dev = align_tokens.device
        tok = align_tokens.detach().float().contiguous()
        mem = None if memory_tokens is None else memory_tokens.detach().to(dev, torch.float32).contiguous()
        nm = int(self.cfg.num_memory_tokens)
        sim3, se3 = torch.empty(B, 1, 8, device=dev), torch.empty(B, max(S - 1, 0), 7, device=dev)
        mem_out = torch.empty(B, nm, 512, device=dev) if nm > 0 else None
        _n.check(_n.lib().lsvs_alignment_decode_forward(self._h, _n.ptr(tok), _i(B), _i(S), _n.ptr(mem), _n.ptr(sim3), _n.ptr(se3),
                                                        _n.ptr(mem_out), _n.stream_ptr()), "alignment_decode_forward")
        return sim3, se3, (mem_out if nm > 0 else memory_tokens)

    def dpt_head_forward(self, prefix: str, taps: Sequence[torch.Tensor], image_hw, output_dim: int, activation: str,
                         frames_chunk: int = 8):
        """UPSTREAM DPTHead.forward: 4 tapped layers (B,S,P,2048) -> pred (B,S,H,W,output_dim-1), conf (B,S,H,W)."""
        B, S, P, C = taps[0].shape
        H, W = image_hw
        assert len(taps) == 4 and C == 2048
        Ho, Wo = 14 * (H // 14), 14 * (W // 14)
        ts = [t.detach().float().contiguous() for t in taps]
        dev = ts[0].device
        pred = torch.empty(B, S, Ho, Wo, output_dim - 1, device=dev)
        conf = torch.empty(B, S, Ho, Wo, device=dev)
        ptrs = (_vp * 4)(*[t.data_ptr() for t in ts])
        act = {"exp": 0, "inv_log": 1}[activation]
        _n.check(_n.lib().lsvs_dpt_head_forward(self._h, prefix.encode(), ptrs, _i(B * S), _i(P), _i(H), _i(W), _i(output_dim), _i(act),
                                                _n.ptr(pred), _n.ptr(conf), _i(frames_chunk if frames_chunk else 0), _n.stream_ptr()),
                 "dpt_head_forward")
        return pred, conf

    def camera_head_forward(self, tokens_last: torch.Tensor, num_iterations: int = 4, all_iterations: bool = False):
        """(B,S,P,2048) -> pose_enc (B,S,9) of the last iteration, or with `all_iterations` the list of every iteration's
        activated encoding (UPSTREAM CameraHead.forward's return value)."""
        B, S, P, C = tokens_last.shape
        tok = tokens_last.detach().float().contiguous()
        out = torch.empty(B, S, 9, device=tok.device)
        iters = torch.empty(num_iterations, B, S, 9, device=tok.device) if all_iterations else None
        _n.check(_n.lib().lsvs_camera_head_forward(self._h, _n.ptr(tok), _i(B), _i(S), _i(P), _i(num_iterations), _n.ptr(out),
                                                   _n.ptr(iters), _n.stream_ptr()), "camera_head_forward")
        return list(iters.unbind(0)) if all_iterations else out


GT_MEAN, GT_SCALE = 1, 2   # LSVS_GT_MEAN / LSVS_GT_SCALE of include/lsvs_b200.h


def pose_chain(chunk_sim3: torch.Tensor, frame_se3: torch.Tensor, cam_enc: torch.Tensor, prev_pose_enc: Optional[torch.Tensor],
               overlap: int, image_hw, gt_poses: Optional[torch.Tensor] = None, gt_mode: int = 0):
    """Pose / Sim(3) composition of one chunk (featureAligned_vggt.py:97-143, :190-196) in one kernel.
    Returns (aligned pose_enc (B,S,9), point transform (B,4,4), chunk scale (B,)).
    gt_poses (B,S,3|4,4) + gt_mode (GT_MEAN | GT_SCALE): the ground-truth variants of the chain (lsvs_pose_chain_gt)."""
    B, S, _ = cam_enc.shape
    dev = cam_enc.device
    H, W = image_hw
    cs = chunk_sim3.detach().float().contiguous()
    fs = frame_se3.detach().float().contiguous()
    ce = cam_enc.detach().float().contiguous()
    prev = None
    S_prev = 0
    if prev_pose_enc is not None:
        prev = prev_pose_enc.detach().to(dev, torch.float32).contiguous()
        S_prev = prev.shape[1]
    pose = torch.empty(B, S, 9, device=dev)
    pt = torch.empty(B, 4, 4, device=dev)
    sc = torch.empty(B, device=dev)
    if gt_poses is not None and gt_mode:
        gt = gt_poses.detach().to(dev, torch.float32).contiguous()
        if gt.dim() != 4 or gt.shape[0] != B or gt.shape[1] != S or gt.shape[-1] != 4 or gt.shape[-2] not in (3, 4):
            raise ValueError(f"gt_poses must be (B,S,3,4) or (B,S,4,4) with B={B}, S={S}; got {tuple(gt.shape)}")
        _n.check(_n.lib().lsvs_pose_chain_gt(_n.ptr(cs), _n.ptr(fs), _n.ptr(ce), _n.ptr(prev), _i(S_prev), _i(overlap), _i(B), _i(S),
                                             _i(H), _i(W), _n.ptr(gt), _i(gt.shape[-2]), _i(gt_mode), _n.ptr(pose), _n.ptr(pt),
                                             _n.ptr(sc), _n.stream_ptr()), "pose_chain_gt")
        return pose, pt, sc
    _n.check(_n.lib().lsvs_pose_chain(_n.ptr(cs), _n.ptr(fs), _n.ptr(ce), _n.ptr(prev), _i(S_prev), _i(overlap), _i(B), _i(S),
                                      _i(H), _i(W), _n.ptr(pose), _n.ptr(pt), _n.ptr(sc), _n.stream_ptr()), "pose_chain")
    return pose, pt, sc


def pose_enc_apply_sim3(pose_enc: torch.Tensor, T: torch.Tensor, s: torch.Tensor, image_hw) -> torch.Tensor:
    """pose_enc (B,S,9) w2c encodings -> Sim(3)-aligned encodings (pointAligned_wrapped_vggt.py:113-122)."""
    B, S, _ = pose_enc.shape
    H, W = image_hw
    pe, Tm, sc = pose_enc.detach().float().contiguous(), T.detach().float().contiguous(), s.detach().float().reshape(B).contiguous()
    out = torch.empty_like(pe)
    _n.check(_n.lib().lsvs_pose_enc_apply_sim3(_n.ptr(pe), _n.ptr(Tm), _n.ptr(sc), _n.ptr(out), _i(B), _i(S), _i(H), _i(W), _n.stream_ptr()),
             "pose_enc_apply_sim3")
    return out

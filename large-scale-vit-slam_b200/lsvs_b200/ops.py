"""Thin Python wrappers over the op-level C-ABI entry points (used by tests, bench and the module mirrors).
Tensors must be CUDA and contiguous; nothing here computes on the CPU."""
import ctypes

import torch

from . import native as _n

EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_F32, EPI_RESID_F32, EPI_HEADNORM64_BF16, EPI_HEADNORM128_BF16 = range(6)
ROPE_NONE, ROPE_2D, ROPE_1D = range(3)

_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


class GemmEpilogue(ctypes.Structure):
    """mirror of `lsvs_gemm_epilogue` in include/lsvs_b200.h"""
    _fields_ = [("bias", _vp), ("out", _vp), ("ldo", _i), ("gamma", _vp), ("resid", _vp), ("ldr", _i), ("out2", _vp),
                ("ld2", _i), ("qn_w", _vp), ("qn_b", _vp), ("kn_w", _vp), ("kn_b", _vp), ("n_q_cols", _i),
                ("n_k_cols", _i), ("ln_eps", _f), ("rope_mode", _i), ("rope_tab", _vp), ("tokens_per_frame", _i),
                ("n_special", _i), ("grid_w", _i), ("pos_ids", _vp), ("pos_period", _i)]


def _p(t):
    return None if t is None else t.data_ptr()


def _cuda(t, dtype, name):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _n.NativeError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} on {t.device}")
    return t


def rope_table(n_pos: int, n_freq: int, base: float = 100.0, device="cuda") -> torch.Tensor:
    tab = torch.empty(n_pos, n_freq, 2, dtype=torch.float32, device=device)
    _n.check(_n.lib().lsvs_rope_table(_n.ptr(tab), _i(n_pos), _i(n_freq), _f(base), _n.stream_ptr()), "rope_table")
    return tab


def gemm(a: torch.Tensor, w: torch.Tensor, kind: int, *, bias=None, out=None, gamma=None, resid=None, out2=None,
         qn=None, kn=None, n_q_cols=0, n_k_cols=0, ln_eps=1e-5, rope_mode=ROPE_NONE, rope_tab=None,
         tokens_per_frame=0, n_special=0, grid_w=0, pos_ids=None) -> torch.Tensor:
    """C = A @ W^T with a fused epilogue (see include/lsvs_b200.h).  a (M,K) bf16 (row stride allowed),
    w (N,K) bf16.  Returns `out` (or `resid` for the residual epilogue)."""
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.is_cuda and a.dtype == torch.bfloat16 and a.stride(1) == 1 and w.is_cuda and w.dtype == torch.bfloat16 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    e = GemmEpilogue()
    e.bias = _p(_cuda(bias, torch.float32, "bias"))
    if kind == EPI_RESID_F32:
        assert resid is not None and resid.dtype == torch.float32 and resid.stride(-1) == 1
        e.resid, e.ldr = _p(resid), resid.stride(0)
        e.gamma = _p(_cuda(gamma, torch.float32, "gamma"))
        if out2 is not None:
            e.out2, e.ld2 = _p(out2), out2.stride(0)
    else:
        if out is None:
            out = torch.empty(M, N, dtype=torch.float32 if kind == EPI_BIAS_F32 else torch.bfloat16, device=a.device)
        assert out.stride(-1) == 1
        e.out, e.ldo = _p(out), out.stride(0)
    if qn is not None:
        e.qn_w, e.qn_b = _p(_cuda(qn[0], torch.float32, "q_norm.weight")), _p(_cuda(qn[1], torch.float32, "q_norm.bias"))
    if kn is not None:
        e.kn_w, e.kn_b = _p(_cuda(kn[0], torch.float32, "k_norm.weight")), _p(_cuda(kn[1], torch.float32, "k_norm.bias"))
    e.n_q_cols, e.n_k_cols, e.ln_eps, e.rope_mode = n_q_cols, n_k_cols, ln_eps, rope_mode
    e.rope_tab = _p(_cuda(rope_tab, torch.float32, "rope_tab"))
    e.tokens_per_frame, e.n_special, e.grid_w = tokens_per_frame, n_special, grid_w
    if pos_ids is not None:
        _cuda(pos_ids, torch.int32, "pos_ids")
        e.pos_ids, e.pos_period = _p(pos_ids), pos_ids.numel()
    _n.check(_n.lib().lsvs_gemm_bf16(_n.ptr(a), _i(a.stride(0)), _n.ptr(w), _i(w.stride(0)), _i(M), _i(N), _i(K), _i(kind),
                                     ctypes.byref(e), _n.stream_ptr()), "gemm_bf16")
    return resid if kind == EPI_RESID_F32 else out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, batches: int, heads: int, head_dim: int, Lq: int,
              Lk: int, scale: float = None, out: torch.Tensor = None) -> torch.Tensor:
    """softmax(q k^T scale) v.  q (batches*Lq, >=heads*head_dim) bf16 with row stride; k, v (batches*Lk, ...)."""
    for t in (q, k, v):
        assert t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    D = heads * head_dim
    if out is None:
        out = torch.empty(batches * Lq, D, dtype=torch.bfloat16, device=q.device)
    if scale is None:
        scale = head_dim ** -0.5
    _n.check(_n.lib().lsvs_attention_bf16(_n.ptr(q), _i(q.stride(0)), _n.ptr(k), _i(k.stride(0)), _n.ptr(v), _i(v.stride(0)),
                                          _n.ptr(out), _i(out.stride(0)), _i(batches), _i(heads), _i(head_dim), _i(Lq), _i(Lk),
                                          _f(scale), _n.stream_ptr()), "attention_bf16")
    return out


# ---- DPT head building blocks (include/lsvs_b200.h: lsvs_conv2d_nhwc_bf16, lsvs_dpt_resample)
DPT_POS_EMBED, DPT_PAD, DPT_CONVT_SHUFFLE, DPT_IM2COL_S2, DPT_BILINEAR = range(5)


def conv2d_nhwc(x: torch.Tensor, w: torch.Tensor, bias=None, res1=None, res2=None, taps: int = 9, relu: bool = False,
                mask_border: bool = True) -> torch.Tensor:
    """x padded NHWC bf16 (frames, hp, wp, C) with a zero border; w bf16 (OC, taps*C) ordered (ky, kx, c)."""
    frames, hp, wp, C = x.shape
    OC = w.shape[0]
    x = _cuda(x, torch.bfloat16, "x"); w = _cuda(w, torch.bfloat16, "w")
    out = torch.empty(frames, hp, wp, OC, dtype=torch.bfloat16, device=x.device)
    _n.check(_n.lib().lsvs_conv2d_nhwc_bf16(_vp(_p(x)), _vp(_p(w)), _vp(_p(_cuda(bias, torch.float32, "bias"))),
                                           _vp(_p(_cuda(res1, torch.bfloat16, "res1"))), _vp(_p(_cuda(res2, torch.bfloat16, "res2"))),
                                           _vp(_p(out)), _i(frames), _i(hp), _i(wp), _i(C), _i(OC), _i(taps), _i(int(relu)),
                                           _i(int(mask_border)), _n.stream_ptr()), "conv2d_nhwc_bf16")
    return out


def dpt_resample(op: int, x: torch.Tensor, out: torch.Tensor, frames: int, h: int, w: int, C: int, a: int = 0, b: int = 0,
                 aspect: float = 1.0, ratio: float = 0.0) -> torch.Tensor:
    _n.check(_n.lib().lsvs_dpt_resample(_i(op), _vp(_p(_cuda(x, torch.bfloat16, "in"))), _vp(_p(_cuda(out, torch.bfloat16, "out"))),
                                       _i(frames), _i(h), _i(w), _i(C), _i(a), _i(b), _f(aspect), _f(ratio), _n.stream_ptr()),
             "dpt_resample")
    return out

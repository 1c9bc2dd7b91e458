"""Deterministic, platform-independent synthetic weights (TEST INFRASTRUCTURE).

``fill_state_dict`` fills a ``name -> shape`` spec from a numpy PCG64 stream keyed by (seed, name), so the
build container (reference + oracle) and the GPU box (product + oracle) get bit-identical weights without
shipping them.  Values are "random-init like" (weights N(0, 0.02)) but with every bias / token / norm
parameter away from its trivial default so that each term of the path is exercised.
"""
import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def _normal(rng, shape, std):
    return (rng.standard_normal(tuple(shape), dtype=np.float32) * np.float32(std)).astype(np.float32)


def synth_param(name: str, shape: Tuple[int, ...], seed: int = 0, ls_gamma: float = None) -> torch.Tensor:
    rng = _rng(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    shape = tuple(shape)
    if leaf == "gamma":  # LayerScale: keep the module's init magnitude unless overridden, +-10 %
        base = ls_gamma if ls_gamma is not None else (1.0 if ".patch_embed.blocks." in name else 0.01)
        v = base * (1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32))
    elif leaf == "memory_token":  # orthonormal rows (alignment_head.py:211-214)
        a = rng.standard_normal((shape[-1], shape[-2]))
        q, _ = np.linalg.qr(a)
        v = q.T.reshape(shape)
    elif leaf == "alpha":
        v = np.full(shape, 0.1)
    elif leaf == "mask_token":
        v = np.zeros(shape)
    elif leaf in ("camera_token", "register_token", "cls_token", "register_tokens", "per_frame_alignment_token",
                  "pos_embed", "empty_pose_tokens"):
        v = _normal(rng, shape, 0.02)
    elif leaf == "weight" and len(shape) == 1:  # LayerNorm scale
        v = 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)
    elif leaf == "weight":
        std = 0.1 if (parent == "2" and "gate_mlp" in name) else 0.02
        v = _normal(rng, shape, std)
    elif leaf == "bias":
        v = _normal(rng, shape, 0.02)
    else:
        raise KeyError(f"no synthetic rule for parameter {name!r} {shape}")
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def fill_state_dict(spec: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0, ls_gamma: float = None) -> Dict[str, torch.Tensor]:
    return {name: synth_param(name, shape, seed, ls_gamma) for name, shape in spec}


def spec_of(module: torch.nn.Module):
    """(name, shape) pairs of a module's persistent state (works for meta-device modules)."""
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]


def checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used to assert both sides generated the same weights."""
    tot = 0.0
    for k in sorted(sd):
        t = sd[k].double()
        tot += float(t.sum()) + 0.5 * float(t.abs().sum())
    return tot

"""Shared helpers for the GPU parity tests: metrics in the north_star's terms."""
import math

import numpy as np
import torch

from oracle import functional as OF
from oracle import weights as OW

TOK_REL_L2 = 1e-2      # north_star: bf16-vs-fp32 token rel-L2 <= 1e-2
TRANS_REL = 1e-3       # per-frame Sim(3)/SE(3) translation within 1e-3 relative
ROT_DEG = 0.05         # rotation within 0.05 degrees
APPLY_TOL = 1e-5       # Sim(3) apply within 1e-5 (fp32)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def rot_err_deg(q_a, q_b):
    """angle between rotations given as (unnormalised) xyzw quaternions, in degrees (max over leading dims)."""
    qa = torch.nn.functional.normalize(q_a.double().cpu(), dim=-1)
    qb = torch.nn.functional.normalize(q_b.double().cpu(), dim=-1)
    d = (qa * qb).sum(-1).abs().clamp(max=1.0)
    return float((2 * torch.acos(d)).max() * 180.0 / math.pi)


def trans_rel_err(t_a, t_b):
    """max over frames of |t_a - t_b| / max(|t_b|, mean |t_b|)."""
    ta, tb = t_a.double().cpu(), t_b.double().cpu()
    n = tb.norm(dim=-1)
    den = torch.maximum(n, n.mean().expand_as(n)).clamp_min(1e-12)
    return float(((ta - tb).norm(dim=-1) / den).max())


def scalar_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())


def pose_metrics(enc_a, enc_b):
    """enc (..., >=7) = [t(3), quat xyzw(4), ...] -> dict(trans_rel, rot_deg)."""
    return {"trans_rel": trans_rel_err(enc_a[..., :3], enc_b[..., :3]), "rot_deg": rot_err_deg(enc_a[..., 3:7], enc_b[..., 3:7])}


def synth_images(seed, B, S, H, W):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(g.random((B, S, 3, H, W), dtype=np.float32))


def load_synth_weights(model, seed=0, ls_gamma=None):
    """Fill a product module with the same deterministic weights oracle/make_golden.py used."""
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    sd = OW.fill_state_dict(spec, seed=seed, ls_gamma=ls_gamma)
    model.load_state_dict(sd, strict=True)
    return sd


REPORT = {}


def report(key, value):
    """Collect measured deviations; written to gpurun_out/precision_report.json when the directory exists (GPU box runs)."""
    import json
    import os
    REPORT[key] = value
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "precision_report.json")
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.update(REPORT)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)

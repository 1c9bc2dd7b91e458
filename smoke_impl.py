"""Body of __graft_entry__.smoke(): small hot-path invocation on cuda:0 checked against the oracle."""
import numpy as np
import torch


def rnd(seed, *shape, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(g.standard_normal(shape, dtype=np.float32) * np.float32(scale))


def run():
    from oracle import aligned as OA
    from oracle import functional as OF
    from aligned_vggt.utils import alignment as A
    dev = torch.device("cuda:0")
    pts = rnd(1, 2, 3, 16, 20, 3, scale=10.0)
    q = torch.nn.functional.normalize(rnd(2, 2, 4), dim=-1)
    T = torch.eye(4).repeat(2, 1, 1)
    T[:, :3, :3] = OF.quat_to_mat(q)
    T[:, :3, 3] = rnd(3, 2, 3)
    s = torch.tensor([0.5, 2.0])
    got = A.apply_sim3_alignment_on_point_maps(pts.to(dev), T.to(dev), s.to(dev)).cpu()
    ref = OA.apply_sim3_points(pts, T, s)
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"sim3 apply mismatch {err}"

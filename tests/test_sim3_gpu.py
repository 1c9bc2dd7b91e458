"""GPU parity: Sim(3) application through the C ABI vs the oracle / golden vectors (tolerance 1e-5, fp32)."""
import pytest
import torch

from conftest import rnd
from oracle import aligned as OA
from oracle import functional as OF

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star: "Sim(3) apply within 1e-5 in fp32"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def make_T(seed, B):
    q = torch.nn.functional.normalize(rnd(seed, B, 4), dim=-1)
    T = torch.eye(4).repeat(B, 1, 1)
    T[:, :3, :3] = OF.quat_to_mat(q)
    T[:, :3, 3] = rnd(seed + 1, B, 3, scale=5.0)
    s = 0.5 + 1.5 * torch.rand(B, generator=torch.Generator().manual_seed(seed))
    return T, s


def test_points_golden(golden):
    from aligned_vggt.utils import alignment as A
    g = golden("geometry.npz")
    out = A.apply_sim3_alignment_on_point_maps(g["pts"].cuda(), g["T"].cuda(), g["s"].cuda()).cpu()
    assert rel(out, g["sim3_points"]) < TOL
    out = A.apply_sim3_alignment_on_w2c(g["extr"].cuda(), g["T"].cuda(), g["s"].cuda()).cpu()
    assert rel(out, g["sim3_w2c"]) < TOL
    out = A.apply_sim3_alignment_on_c2w(g["c2w"].cuda(), g["T"].cuda(), g["s"].cuda()).cpu()
    assert rel(out, g["sim3_c2w"]) < TOL


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 2, 3, 5), (3, 2, 7, 9), (2, 4, 154, 518), (1, 3, 17, 13), (5, 1, 2, 2)])
def test_points_vs_oracle(shape):
    from aligned_vggt.utils import alignment as A
    B, S, H, W = shape
    pts = rnd(7, B, S, H, W, 3, scale=10.0)
    T, s = make_T(11, B)
    got = A.apply_sim3_alignment_on_point_maps(pts.cuda(), T.cuda(), s.cuda()).cpu()
    assert got.shape == pts.shape
    assert rel(got, OA.apply_sim3_points(pts, T, s)) < TOL


def test_points_unbatched_and_identity_and_errors():
    from aligned_vggt.utils import alignment as A
    pts = rnd(8, 2, 6, 10, 3)
    T, s = make_T(12, 1)
    got = A.apply_sim3_alignment_on_point_maps(pts.cuda(), T[0].cuda(), s[0].cuda()).cpu()
    assert got.shape == (1, 2, 6, 10, 3)
    assert rel(got[0], OA.apply_sim3_points(pts, T[0], s[0])[0]) < TOL
    same = A.apply_sim3_alignment_on_point_maps(pts.cuda(), torch.eye(4).cuda(), torch.ones(()).cuda()).cpu()
    assert torch.equal(same[0], pts)  # identity Sim(3) is exact
    with pytest.raises(AssertionError):
        A.apply_sim3_alignment_on_point_maps(pts[None].cuda(), torch.eye(4).repeat(2, 1, 1).cuda(), torch.ones(1).cuda())


def test_points_full_size_roundtrip():
    """BASELINE config-2 size (32 x 154 x 518 points): apply then apply the inverse Sim(3) returns the input."""
    from aligned_vggt.utils import alignment as A
    B, S, H, W = 1, 32, 154, 518
    pts = (torch.randn(B, S, H, W, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) * 10)
    T, s = make_T(13, B)
    fwd = A.apply_sim3_alignment_on_point_maps(pts, T.cuda(), s.cuda())
    Tinv = torch.eye(4).repeat(B, 1, 1)
    Tinv[:, :3, :3] = T[:, :3, :3].transpose(1, 2) / s.view(B, 1, 1)
    Tinv[:, :3, 3] = -(Tinv[:, :3, :3] @ T[:, :3, 3:]).squeeze(-1)
    back = A.apply_sim3_alignment_on_point_maps(fwd, Tinv.cuda(), torch.ones(B).cuda())
    assert rel(back, pts) < 1e-5
    # linearity: f(a p) - t = a (f(p) - t)
    t = T[:, :3, 3].cuda().view(B, 1, 1, 1, 3)
    f2 = A.apply_sim3_alignment_on_point_maps(2 * pts, T.cuda(), s.cuda())
    assert rel(f2 - t, 2 * (fwd - t)) < 1e-5


def test_depth_scale_and_poses():
    from aligned_vggt.utils import alignment as A
    d = rnd(9, 3, 2, 5, 7, 1).abs()
    s = torch.tensor([0.5, 1.5, 3.0])
    assert torch.equal(A.scale_depth(d.cuda(), s.cuda()).cpu(), d * s.view(3, 1, 1, 1, 1))
    B, S = 3, 5
    T, s = make_T(14, B)
    extr = torch.cat([OF.quat_to_mat(torch.nn.functional.normalize(rnd(15, B, S, 4), dim=-1)), rnd(16, B, S, 3, 1)], -1)
    got = A.apply_sim3_alignment_on_w2c(extr.cuda(), T.cuda(), s.cuda()).cpu()
    assert rel(got, OA.apply_sim3_w2c(extr, T, s)) < TOL
    c2w = OA.inv_se3(extr)
    assert rel(A.apply_sim3_alignment_on_c2w(c2w.cuda(), T.cuda(), s.cuda()).cpu(), OA.apply_sim3_c2w(c2w, T, s)) < TOL


def test_apply_sim3_alignment_on_dict_golden(golden):
    """reference alignment.py:428-489: numpy transforms / scales, pose encodings + points + depth (depth in place)."""
    from aligned_vggt.utils import alignment as A
    from parity_util import ROT_DEG, TRANS_REL, pose_metrics, scalar_rel
    g = golden("sim3_dict.npz")
    dep = g["dep"].cuda()
    pred = {"pose_enc": g["enc"].cuda(), "world_points": g["pts"].cuda(), "depth": dep}
    A.apply_sim3_alignment_on_dict(pred, (g["H"], g["W"]), g["T"].numpy(), g["s"].numpy())
    m = pose_metrics(pred["pose_enc"], g["out_pose_enc"])
    assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG and scalar_rel(pred["pose_enc"][..., 7:], g["out_pose_enc"][..., 7:]) < 1e-5
    assert float((pred["world_points"].cpu() - g["out_world_points"]).abs().max()) <= 1e-5 * float(g["out_world_points"].abs().max())
    assert pred["depth"].data_ptr() == dep.data_ptr()                       # scaled in place like `depths *= ...` (:487)
    assert float((dep.cpu() - g["out_depth"]).abs().max()) <= 1e-6 * float(g["out_depth"].abs().max())
    only_pose = {"pose_enc": g["enc"].cuda()}
    A.apply_sim3_alignment_on_dict(only_pose, (g["H"], g["W"]), g["T"].numpy(), g["s"].numpy())
    assert set(only_pose) == {"pose_enc"} and torch.equal(only_pose["pose_enc"], pred["pose_enc"])

"""GPU parity of the baseline rows: IRLS weighted Umeyama, the pose-aligned chain, the point-aligned
pose update, and the two `VGGT` wrapper classes, vs the golden vectors generated from the reference's own code."""
import pytest
import torch

from conftest import rnd
from oracle import aligned as OA
from oracle import functional as OF
from oracle import weights as OW
from parity_util import ROT_DEG, TRANS_REL, load_synth_weights, pose_metrics, rel_l2, scalar_rel, synth_images

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def rot_deg(Ra, Rb):
    c = ((Ra.double().cpu().T @ Rb.double().cpu()).trace() - 1) / 2
    return float(torch.rad2deg(torch.acos(c.clamp(-1, 1))))


def test_irls_umeyama_golden(golden):
    from aligned_vggt.models.pointAligned_wrapped_vggt import irls_sim3_umeyama, weighted_umeyama_sim3
    g = golden("geometry.npz")
    R, t, s = irls_sim3_umeyama(g["u_src"].cuda(), g["u_dst"].cuda(), g["u_cs"].cuda(), g["u_cd"].cuda())
    assert rot_deg(R, g["irls_R"]) < ROT_DEG and rel_l2(t, g["irls_t"]) < TRANS_REL and abs(float(s) - g["irls_s"]) < 1e-4 * g["irls_s"]
    w = torch.sqrt(g["u_cs"] * g["u_cd"]).reshape(-1)
    R, t, s = weighted_umeyama_sim3(g["u_src"].reshape(-1, 3).cuda(), g["u_dst"].reshape(-1, 3).cuda(), w.cuda())
    assert rot_deg(R, g["um_R"]) < ROT_DEG and rel_l2(t, g["um_t"]) < TRANS_REL and abs(float(s) - g["um_s"]) < 1e-4 * g["um_s"]


@pytest.mark.parametrize("n,h,w,noise,outliers", [(1, 3, 5, 0.0, 0), (2, 30, 40, 0.02, 25), (8, 154, 518, 0.01, 5000), (3, 17, 13, 0.2, 40)])
def test_irls_umeyama_vs_oracle(n, h, w, noise, outliers):
    """incl. BASELINE config-2 overlap size (8 x 154 x 518 = 638 k points): exact median select + 21 solves on device."""
    from aligned_vggt.models.pointAligned_wrapped_vggt import irls_sim3_umeyama
    src = rnd(1, n, h, w, 3, scale=4.0)
    Rg = OF.quat_to_mat(torch.nn.functional.normalize(rnd(2, 4), dim=-1))
    dst = 1.7 * (src @ Rg.T) + torch.tensor([0.5, -1.0, 2.0]) + rnd(3, n, h, w, 3, scale=noise)
    if outliers:
        idx = torch.randperm(n * h * w, generator=torch.Generator().manual_seed(4))[:outliers]
        dst.view(-1, 3)[idx] += 5.0
    cs, cd = 1 + torch.exp(rnd(5, n, h, w)), 1 + torch.exp(rnd(6, n, h, w))
    Rr, tr, sr = OA.irls_umeyama(src, dst, cs, cd)
    R, t, s = irls_sim3_umeyama(src.cuda(), dst.cuda(), cs.cuda(), cd.cuda())
    assert rot_deg(R, Rr) < ROT_DEG, rot_deg(R, Rr)
    assert float((t.cpu() - tr).norm() / tr.norm()) < TRANS_REL
    assert abs(float(s) - float(sr)) < 1e-3 * float(sr)
    if noise == 0.0:  # noise-free points: exact recovery of the generating Sim(3) (KAT)
        assert rot_deg(R, Rg) < 1e-2 and abs(float(s) - 1.7) < 1e-4


def test_irls_umeyama_errors():
    from aligned_vggt.models.pointAligned_wrapped_vggt import irls_sim3_umeyama
    from lsvs_b200 import native
    src = rnd(1, 1, 4, 4, 3)
    with pytest.raises(ValueError):  # zero confidences -> total weight too small (pointAligned_wrapped_vggt.py:184-185)
        irls_sim3_umeyama(src.cuda(), src.cuda(), torch.zeros(1, 4, 4).cuda(), torch.zeros(1, 4, 4).cuda())
    with pytest.raises(native.NativeError):
        irls_sim3_umeyama(src, src, torch.ones(1, 4, 4), torch.ones(1, 4, 4))


def test_pose_enc_apply_sim3_golden(golden):
    from lsvs_b200.engine import pose_enc_apply_sim3
    g = golden("pose_enc_sim3.npz")
    out = pose_enc_apply_sim3(g["enc"].cuda(), g["T"].cuda(), g["s"].cuda(), (g["H"], g["W"]))
    m = pose_metrics(out, g["out"])
    assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG and scalar_rel(out[..., 7:], g["out"][..., 7:]) < 1e-5


def test_pose_aligned_model_golden(golden):
    from aligned_vggt.models.poseAligned_wrapped_vggt import VGGT
    g = golden("model_pose_aligned_small.npz")
    taps = (0, 0, 1, 1)
    model = VGGT(enable_point=False, enable_depth=False, enable_track=False, depth=2, patch_embed_depth=2, intermediate_layer_indices=taps)
    sd = load_synth_weights(model, seed=0)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"])
    model = model.cuda().eval()
    S, H, W, ov = g["S"], g["H"], g["W"], g["ov"]
    imgs = [synth_images(100 + i, 1, S, H, W) for i in range(2)]
    pts = [rnd(200 + i, 1, S, H, W, 3, scale=5.0) for i in range(2)]
    conf = torch.ones(1, S, H, W)
    p1 = model(imgs[0].cuda(), ov, None, raw_points=pts[0].cuda(), raw_points_conf=conf.cuda())
    e1 = p1["pose_enc"][-1].clone()
    p2 = model(imgs[1].cuda(), ov, p1, raw_points=pts[1].cuda(), raw_points_conf=conf.cuda())
    e2 = p2["pose_enc"][-1]
    # reference's own bf16-autocast deviation is not available for this wrapper (camera head runs fp32 there): use the
    # feature-aligned calibration band (tests/test_model_gpu.py): pose_enc within 1e-2 rel / 1 deg of the fp32 run
    for e, k in ((e1, "c1_pose_enc"), (e2, "c2_pose_enc")):
        m = pose_metrics(e, g[k])
        assert m["trans_rel"] < 1e-2 and m["rot_deg"] < 1.0, (k, m)
    # the point maps get exactly the transform the path decoded
    o1 = OA.pose_aligned_forward(sd, imgs[0], ov, None, raw_points=pts[0], depth=2, dino_depth=2, taps=taps)
    o2 = OA.pose_aligned_forward(sd, imgs[1], ov, {"pose_enc": o1["pose_enc"]}, raw_points=pts[1], depth=2, dino_depth=2, taps=taps)
    assert rel_l2(p2["world_points"][0], o1["world_points"]) < 2e-2 and rel_l2(p2["world_points"][1], o2["world_points"]) < 2e-2
    assert len(p2["pose_enc"]) == 2 and "images" in p2


def test_point_aligned_model_chain():
    """point-aligned wrapper: two chunks whose raw point maps differ by a known Sim(3) on the overlap -> the second chunk is
    mapped onto the first (the reference's wrapper cannot be run without the DPT head, so this is a property test on
    top of the function-level goldens)."""
    from aligned_vggt.models.pointAligned_wrapped_vggt import VGGT
    taps = (0, 0, 0, 0)
    model = VGGT(enable_point=False, enable_depth=False, enable_track=False, depth=1, patch_embed_depth=1, intermediate_layer_indices=taps)
    load_synth_weights(model, seed=0)
    model = model.cuda().eval()
    S, H, W, ov = 4, 28, 42, 2
    imgs = [synth_images(100 + i, 1, S, H, W).cuda() for i in range(2)]
    world = rnd(7, 1, S + 2, H, W, 3, scale=5.0)            # 6 frames of "true" points; chunk 2 starts at frame 2
    Rg = OF.quat_to_mat(torch.nn.functional.normalize(rnd(8, 4), dim=-1))
    sg, tg = 0.8, torch.tensor([1.0, 2.0, -0.5])
    pts1 = world[:, :S]
    pts2 = ((world[:, 2:] - tg) @ Rg) / sg                  # chunk-2 frame: world = sg * R * p + tg
    conf = 1 + torch.exp(rnd(9, 1, S, H, W))
    dep = rnd(10, 1, S, H, W, 1).abs() + 0.1
    p1 = model(imgs[0], ov, None, raw_points=pts1.cuda(), raw_points_conf=conf.cuda(), raw_depth=dep.cuda(), raw_depth_conf=conf.cuda())
    assert torch.equal(p1["world_points"][0].cpu(), pts1)   # first chunk: identity Sim(3) is exact
    cam1 = p1["pose_enc"][0].clone()
    p2 = model(imgs[1], ov, p1, raw_points=pts2.cuda(), raw_points_conf=conf.cuda(), raw_depth=dep.cuda(), raw_depth_conf=conf.cuda())
    assert rel_l2(p2["world_points"][1], world[:, 2:]) < 1e-4                      # chunk 2 lands on the world points
    assert rel_l2(p2["depth"][1], dep * sg) < 1e-4                                 # depth scaled by the estimated scale
    # pose update == oracle's pose_enc -> w2c -> Sim(3) -> pose_enc with the generating transform
    cam2_raw = model.camera_head([model.aggregator(imgs[1])[0][0]])[-1].cpu()
    T = torch.eye(4)[None].clone(); T[0, :3, :3] = Rg; T[0, :3, 3] = tg
    ref = OA.pose_enc_apply_sim3(cam2_raw, (H, W), T, torch.tensor([sg]))
    m = pose_metrics(p2["pose_enc"][1], ref)
    assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG, m
    assert torch.equal(p2["pose_enc"][0], cam1)


# ---- gt_poses variants and the baseline wrappers with their DPT heads (tests/golden/model_baselines_gt_dpt.npz: outputs of the
# ---- reference classes' own forwards; two chained chunks whose overlap frames are the same images) -------------------------
def _gt_dpt_images(g):
    import numpy as np
    S, H, W, ov = g["S"], g["H"], g["W"], g["ov"]
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(500 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i in range(2)]
    imgs[1][:, :ov] = imgs[0][:, -ov:]
    return imgs


def _check_pose(e, ref, what):
    m = pose_metrics(e, ref)   # bf16 encoder + camera trunk vs the fp32 reference run: calibration band of tests/test_model_gpu.py
    assert m["trans_rel"] < 1e-2 and m["rot_deg"] < 1.0, (what, m)


def test_feature_aligned_gt_poses_golden(golden):
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    g = golden("model_baselines_gt_dpt.npz")
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False, depth=1, patch_embed_depth=1,
                               intermediate_layer_indices=(0, 0, 0, 0))
    sd = load_synth_weights(model, seed=0)
    assert abs(OW.checksum(sd) - g["fa_wsum"]) < 1e-6 * abs(g["fa_wsum"])
    model = model.cuda().eval()
    imgs, ov = _gt_dpt_images(g), g["ov"]
    p1 = model(imgs[0].cuda(), ov, None, gt_poses=g["fa_gt1"].cuda())
    _check_pose(p1["pose_enc"][-1], g["fa_c1_pose_enc"], "c1")
    ctx_plain = {k: (list(v) if isinstance(v, list) else v) for k, v in p1.items()}
    p2 = model(imgs[1].cuda(), ov, p1, gt_poses=g["fa_gt2"])          # CPU gt tensor: moved like the reference's .to(extr)
    _check_pose(p2["pose_enc"][-1], g["fa_c2_pose_enc"], "c2")
    p2_plain = model(imgs[1].cuda(), ov, ctx_plain)
    assert float((p2_plain["pose_enc"][-1] - p2["pose_enc"][-1]).abs().max()) > 1e-2


@pytest.mark.parametrize("tag", ["pa", "pagt"])
def test_pose_aligned_with_dpt_heads_golden(golden, tag):
    from aligned_vggt.models.poseAligned_wrapped_vggt import VGGT
    g = golden("model_baselines_gt_dpt.npz")
    sub, ov = g["sub"], g["ov"]
    model = VGGT(enable_track=False, depth=1, patch_embed_depth=1, intermediate_layer_indices=(0, 0, 0, 0))
    sd = load_synth_weights(model, seed=2)
    assert abs(OW.checksum(sd) - g["pa_wsum"]) < 1e-6 * abs(g["pa_wsum"])
    model = model.cuda().eval()
    imgs = _gt_dpt_images(g)
    p = None
    for ci in (1, 2):
        gt = g[f"pagt_gt{ci}"].cuda() if tag == "pagt" else None
        p = model(imgs[ci - 1].cuda(), ov, p, gt_poses=gt)
        assert len(p["depth"]) == ci and len(p["world_points"]) == ci
        _check_pose(p["pose_enc"][-1], g[f"{tag}_c{ci}_pose_enc"], (tag, ci))
        for k, tol in (("depth", 3e-2), ("depth_conf", 3e-2), ("world_points", 5e-2), ("world_points_conf", 3e-2)):
            got, ref = p[k][-1][:, :, ::sub, ::sub], g[f"{tag}_c{ci}_{k}"]
            assert got.shape == ref.shape and rel_l2(got, ref) < tol, (tag, ci, k, rel_l2(got, ref))
    if tag == "pagt":
        with pytest.raises(RuntimeError):   # the reference pads one row: only (B,S,3,4) works there
            model(imgs[0].cuda(), ov, None, gt_poses=torch.eye(4).expand(1, g["S"], 4, 4).cuda())


def test_point_aligned_with_dpt_heads_golden(golden):
    from aligned_vggt.models.pointAligned_wrapped_vggt import VGGT
    g = golden("model_baselines_gt_dpt.npz")
    sub, ov = g["sub"], g["ov"]
    model = VGGT(enable_track=False, depth=1, patch_embed_depth=1, intermediate_layer_indices=(0, 0, 0, 0))
    load_synth_weights(model, seed=2)
    model = model.cuda().eval()
    imgs = _gt_dpt_images(g)
    p = None
    for ci in (1, 2):
        p = model(imgs[ci - 1].cuda(), ov, p)
        _check_pose(p["pose_enc"][-1], g[f"pt_c{ci}_pose_enc"], ci)
        for k, tol in (("depth", 3e-2), ("depth_conf", 3e-2), ("world_points", 5e-2), ("world_points_conf", 3e-2)):
            got, ref = p[k][-1][:, :, ::sub, ::sub], g[f"pt_c{ci}_{k}"]
            assert got.shape == ref.shape and rel_l2(got, ref) < tol, (ci, k, rel_l2(got, ref))

"""GPU parity of the module-level drop-ins (Aggregator, AlignmentHead, CameraHead, FeatureAlignedVGGT, pose chain)
against the golden vectors generated from the reference's own code and against the CPU oracle.

Tolerances (north_star): token rel-L2 <= 1e-2 (bf16 tensor-core path vs fp32 reference); per-frame Sim(3)/SE(3)
translation 1e-3 relative, rotation 0.05 deg for the fp32 decode / pose stages given the same inputs.
End to end (bf16 encoder feeding the fp32 decode) the decoded transforms inherit the encoder's bf16 error, exactly
as they do in the reference's own shipping precision (Lightning bf16-mixed): the tests measure that deviation on the
CPU oracle (amp=True) and require the CUDA path to be no further from fp32 than 3x it (both are single noise
realisations of the same bf16 rounding level, so a factor is needed; measured ratios are 0.4x .. 2x)."""
import numpy as np
import pytest
import torch

from conftest import rnd
from oracle import aligned as OA
from oracle import functional as OF
from oracle import weights as OW
from parity_util import (ROT_DEG, TOK_REL_L2, TRANS_REL, load_synth_weights, pose_metrics, rel_l2, scalar_rel, synth_images)

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


# BASELINE config 1, bf16 pipeline vs the reference's fp32 outputs — (relative translation, rotation in degrees).  The deviation is
# one realisation of bf16 rounding noise: any change of summation order in a kernel gives another one (two builds of this round
# measured Sim(3) 1.8e-3 / 0.16 deg and 0.6e-3 / 0.27 deg; SE(3) 6.2e-3 / 0.49 deg, pose 8.8e-3 / 0.59 deg —
# profiles/r2_precision_report.json), so the bounds are 3x the typical measured values, not a tight envelope of one build.
CONFIG1_BOUNDS = {"chunk_sim3_alignment_enc": (5e-3, 0.6), "frame_se3_alignment_enc": (1.8e-2, 1.5), "pose_enc": (2.6e-2, 1.8)}


def _report(key, value):
    import json
    import os
    from conftest import ROOT
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "precision_report.json")
        old = json.load(open(path)) if os.path.exists(path) else {}
        old[key] = value
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)


def within(mine, ref_amp, north_star, slack=3.0):
    return mine <= max(slack * ref_amp, north_star)


# ------------------------------------------------------------------------------------------------ fp32 stages
@pytest.mark.parametrize("B,S", [(1, 4), (2, 5), (1, 32), (1, 2)])
def test_decode_fp32_parity(B, S):
    """AlignmentHead._decode_alignments + GatedUpdate in fp32 (alignment_head.py:427-540): same inputs, north_star tolerance."""
    from aligned_vggt.heads.alignment_head import AlignmentHead
    head = AlignmentHead()
    sd = load_synth_weights(head, seed=11, ls_gamma=0.3)
    head = head.cuda()
    tok = rnd(50 + S, B, S, 1024, scale=2.0)
    ref1 = OA.decode_alignments(sd, "", tok, True, None)
    got1 = head._decode_alignments(tok.cuda(), 0, True, None)
    tok2 = rnd(60 + S, B, S, 1024, scale=2.0)
    ref2 = OA.decode_alignments(sd, "", tok2, False, ref1[2])
    got2 = head._decode_alignments(tok2.cuda(), 0, False, ref1[2].cuda())
    for ref, got in ((ref1, got1), (ref2, got2)):
        m = pose_metrics(got[0], ref[0])
        assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG, m
        assert scalar_rel(got[0][..., 7], ref[0][..., 7]) < 1e-4
        m = pose_metrics(got[1], ref[1])
        assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG, m
        assert rel_l2(got[2], ref[2]) < 1e-4
        assert float((got[2].norm(dim=-1) - 1).abs().max()) < 1e-5  # memory rows stay unit norm


@pytest.mark.parametrize("B,S,ov", [(1, 4, 2), (2, 6, 1), (1, 32, 8), (1, 3, 3)])
def test_pose_chain_parity(B, S, ov):
    """featureAligned_vggt.py:97-143,190-196 in one kernel vs the restated reference math."""
    from lsvs_b200.engine import pose_chain
    H, W = 154, 518
    sim3 = torch.cat([rnd(1, B, 1, 3), rnd(2, B, 1, 4), 0.5 + torch.rand(B, 1, 1, generator=torch.Generator().manual_seed(3))], -1)
    se3 = torch.cat([rnd(4, B, S - 1, 3), rnd(5, B, S - 1, 4)], -1)
    cam = torch.cat([rnd(6, B, S, 3), torch.nn.functional.normalize(rnd(7, B, S, 4), dim=-1) * 1.1, 0.5 + 0.3 * torch.rand(B, S, 2, generator=torch.Generator().manual_seed(8))], -1)
    prev_q = torch.nn.functional.normalize(rnd(9, B, S + 1, 4), dim=-1)
    prev = torch.cat([rnd(10, B, S + 1, 3), torch.where(prev_q[..., 3:] < 0, -prev_q, prev_q), 0.6 * torch.ones(B, S + 1, 2)], -1)
    for prev_enc in (None, prev):
        per_frame, scale = OA.compose_alignment(sim3, se3)
        ref_enc, pf, ident = OA.pose_chain(cam, (H, W), per_frame, scale, prev_enc, ov)
        ref_T = OA.point_transform(pf, ident, prev_enc is not None)
        enc, T, sc = pose_chain(sim3.cuda(), se3.cuda(), cam.cuda(), None if prev_enc is None else prev_enc.cuda(), ov, (H, W))
        m = pose_metrics(enc, ref_enc)
        assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG, m
        assert scalar_rel(enc[..., 7:], ref_enc[..., 7:]) < 1e-5
        assert rel_l2(T, ref_T) < 1e-5 and rel_l2(sc, scale.reshape(B)) < 1e-7


# ------------------------------------------------------------------------------------------------ alignment head
def test_alignment_head_golden(golden):
    from aligned_vggt.heads.alignment_head import AlignmentHead
    g = golden("head_temporal.npz")
    head = AlignmentHead()
    sd = load_synth_weights(head, seed=7, ls_gamma=0.2)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"]), "synthetic weights differ from the golden run"
    head = head.cuda()
    S, gh, gw, ov = g["S"], g["gh"], g["gw"], g["ov"]
    P = 5 + gh * gw
    tok = [rnd(30, 1, S, P, 2048), rnd(31, 1, S, P, 2048)]
    r1 = head(tok[0].cuda(), (gh * 14, gw * 14), ov)
    r2 = head(tok[1].cuda(), (gh * 14, gw * 14), ov, overlap_tokens=r1[3], memory_tokens=r1[2])
    # the reference's own bf16-autocast deviation from fp32 on the same case
    a1 = OA.alignment_head_forward(sd, "", tok[0], (gh * 14, gw * 14), ov, amp=True)
    a2 = OA.alignment_head_forward(sd, "", tok[1], (gh * 14, gw * 14), ov, a1[3], a1[2], amp=True)
    for c, r, a in (("c1", r1, a1), ("c2", r2, a2)):
        assert r[3].shape == (1, 1 + ov, P + 1, 1024) and r[3].is_contiguous()
        assert rel_l2(r[3], g[c + "_overlap"]) < TOK_REL_L2
        assert rel_l2(r[2], g[c + "_mem"]) < TOK_REL_L2
        for i, key in ((0, "_sim3"), (1, "_se3")):
            mine, amp = pose_metrics(r[i], g[c + key]), pose_metrics(a[i], g[c + key])
            assert within(mine["trans_rel"], amp["trans_rel"], TRANS_REL), (c, key, mine, amp)
            assert within(mine["rot_deg"], amp["rot_deg"], ROT_DEG), (c, key, mine, amp)


def test_alignment_head_errors():
    from aligned_vggt.heads.alignment_head import AlignmentHead
    from lsvs_b200 import native
    with pytest.raises(AttributeError):
        AlignmentHead(temporal_attention=False)
    with pytest.raises(ValueError):
        AlignmentHead(depth_aa=4, aa_block_size=3)
    head = AlignmentHead()
    with pytest.raises(native.NativeError):  # parameters still on the CPU: no fallback
        head(torch.zeros(1, 2, 29, 2048, device="cuda"), (56, 84), 1)
    head = head.cuda()
    with pytest.raises(AssertionError):  # overlap tokens of another patch grid (alignment_head.py:251)
        head(torch.zeros(1, 2, 29, 2048, device="cuda"), (56, 84), 1, overlap_tokens=torch.zeros(1, 2, 31, 1024, device="cuda"))
    with pytest.raises(ValueError):  # token count does not match the image size
        head(torch.zeros(1, 2, 30, 2048, device="cuda"), (56, 84), 1)


# ------------------------------------------------------------------------------------------------ aggregator
@pytest.mark.parametrize("S,H,W", [(2, 28, 42), (1, 518, 518), (3, 154, 518)])
def test_aggregator_vs_oracle(S, H, W):
    """1+1 DINO / 1 alternating pair on odd shapes (square image: native 37x37 position grid, no interpolation)."""
    from lsvs_b200.modules import Aggregator
    agg = Aggregator(depth=1, patch_embed_depth=1)
    sd = load_synth_weights(agg, seed=21, ls_gamma=0.2)
    agg = agg.cuda()
    img = synth_images(5, 1, S, H, W)
    ref, start = OF.aggregator_forward(sd, "", img, depth=1, dino_depth=1)
    out, start2 = agg(img.cuda())
    assert start == start2 == 5 and out[0].shape == ref[0].shape == (1, S, 5 + (H // 14) * (W // 14), 2048)
    assert rel_l2(out[0], ref[0]) < TOK_REL_L2
    assert rel_l2(out[0][..., :1024], ref[0][..., :1024]) < TOK_REL_L2  # frame half and global half separately
    assert rel_l2(out[0][..., 1024:], ref[0][..., 1024:]) < TOK_REL_L2


# ------------------------------------------------------------------------------------------------ full model
def _run_model_case(golden, tag, depth, dino, taps, precision=None):
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    g = golden(f"model_{tag}.npz")
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False, depth=depth, patch_embed_depth=dino,
                               intermediate_layer_indices=taps, precision=precision)
    sd = load_synth_weights(model, seed=0)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"]), "synthetic weights differ from the golden run"
    model = model.cuda().eval()
    S, H, W, ov, st = g["S"], g["H"], g["W"], g["ov"], g["sample_stride"]
    imgs = [synth_images(100 + i, 1, S, H, W) for i in range(2)]
    pts = [rnd(200 + i, 1, S, H, W, 3, scale=5.0) for i in range(2)]
    dep = [rnd(300 + i, 1, S, H, W, 1).abs() + 0.1 for i in range(2)]
    out = []
    p = None
    for i in range(2):
        p = model(imgs[i].cuda(), ov, p, raw_depth=dep[i].cuda(), raw_points=pts[i].cuda())
        snap = {k: (v[-1] if isinstance(v, list) else v).clone() for k, v in p.items()}
        snap["chunk_sim3_alignment_enc"] = p["chunk_sim3_alignment_enc"][:, -1:]
        snap["frame_se3_alignment_enc"] = p["frame_se3_alignment_enc"][:, -(S - 1):]
        snap["tap_last"] = model.aggregator(imgs[i].cuda())[0][taps[-1]]
        out.append(snap)
    assert set(p.keys()) >= {"pose_enc", "chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "overlap_tokens", "memory_tokens", "depth", "world_points", "images"}
    assert len(p["pose_enc"]) == 2 and p["chunk_sim3_alignment_enc"].shape == (1, 2, 8) and p["frame_se3_alignment_enc"].shape == (1, 2 * (S - 1), 7)
    return g, sd, out, imgs, pts, dep, (S, H, W, ov, st)


def _check_tokens(g, out, st):
    for c, snap in (("c1", out[0]), ("c2", out[1])):
        assert rel_l2(snap["tap_last"][..., ::st], g[c + "_tap_last"]) < TOK_REL_L2
        assert rel_l2(snap["overlap_tokens"][..., ::st], g[c + "_overlap_tokens"]) < TOK_REL_L2
        assert rel_l2(snap["memory_tokens"], g[c + "_memory_tokens"]) < TOK_REL_L2


def test_model_small_golden(golden):
    taps = (0, 0, 1, 1)
    g, sd, out, imgs, pts, dep, (S, H, W, ov, st) = _run_model_case(golden, "small", 2, 2, taps)
    _check_tokens(g, out, st)
    # reference's own bf16-autocast deviation on this case (CPU oracle, amp=True)
    a1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, depth=2, dino_depth=2, taps=taps, amp=True)
    ctx = {"overlap_tokens": a1["overlap_tokens"], "memory_tokens": a1["memory_tokens"], "pose_enc": a1["pose_enc"]}
    a2 = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, depth=2, dino_depth=2, taps=taps, amp=True)
    for c, snap, amp in (("c1", out[0], a1), ("c2", out[1], a2)):
        for key in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "pose_enc"):
            mine, ref_amp = pose_metrics(snap[key], g[f"{c}_{key}"]), pose_metrics(amp[key], g[f"{c}_{key}"])
            assert within(mine["trans_rel"], ref_amp["trans_rel"], TRANS_REL), (c, key, mine, ref_amp)
            assert within(mine["rot_deg"], ref_amp["rot_deg"], ROT_DEG), (c, key, mine, ref_amp)
    # Sim(3) application on the stand-in maps: exactly the transform the path decoded (1e-5, fp32)
    for i, snap in enumerate(out):
        scale = snap["chunk_sim3_alignment_enc"][0, 0, 7].cpu()
        assert rel_l2(snap["depth"], dep[i] * scale) < 1e-6
    o1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, depth=2, dino_depth=2, taps=taps, raw_points=pts[0])
    assert rel_l2(out[0]["world_points"], o1["world_points"]) < 2e-2  # transform from bf16 tokens vs fp32 oracle


def test_model_full_golden(golden):
    """BASELINE config 1 (4 frames of 518x154, full 24+24+24 depth, two chained chunks) vs the reference run."""
    g, sd, out, imgs, pts, dep, (S, H, W, ov, st) = _run_model_case(golden, "full", 24, 24, (4, 11, 17, 23))
    _check_tokens(g, out, st)
    measured = {}
    for c, snap in (("c1", out[0]), ("c2", out[1])):
        for key in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "pose_enc"):
            measured[f"{c}_{key}"] = pose_metrics(snap[key], g[f"{c}_{key}"])
    _report("model_full_config1_precision0", measured)
    # bf16 pipeline vs the reference's fp32 outputs: bounds calibrated on the measured deviation (B200: see DESIGN.md section 3;
    # the fp32-class pipeline meets 1e-3 / 0.05 deg on the benchmarked configuration, tests/test_precision_gpu.py)
    for name, m in measured.items():
        tr, rd = CONFIG1_BOUNDS[name.split("_", 1)[1]]
        assert m["trans_rel"] < tr and m["rot_deg"] < rd, (name, m)


def test_model_full_golden_fp32_class(golden):
    """BASELINE config 1 with every block fp32-class (engine precision 2): the north_star Sim(3) / SE(3) / pose numbers hold against
    the reference's fp32 outputs, both chained chunks."""
    g, sd, out, imgs, pts, dep, (S, H, W, ov, st) = _run_model_case(golden, "full", 24, 24, (4, 11, 17, 23), precision=2)
    measured = {}
    for c, snap in (("c1", out[0]), ("c2", out[1])):
        assert rel_l2(snap["tap_last"][..., ::st], g[c + "_tap_last"]) < 1e-3
        assert rel_l2(snap["overlap_tokens"][..., ::st], g[c + "_overlap_tokens"]) < 1e-3
        for key in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "pose_enc"):
            m = measured[f"{c}_{key}"] = pose_metrics(snap[key], g[f"{c}_{key}"])
            assert m["trans_rel"] < TRANS_REL and m["rot_deg"] < ROT_DEG, (c, key, m)
    _report("model_full_config1_precision2", measured)


def test_model_state_and_errors():
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    from lsvs_b200 import native
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False, depth=1, patch_embed_depth=1,
                               intermediate_layer_indices=(0, 0, 0, 0))
    img = synth_images(1, 1, 2, 28, 42)
    with pytest.raises(native.NativeError):
        model(img.cuda(), 1)  # weights on CPU -> loud failure, no fallback
    model = model.cuda().eval()
    p1 = model(img.cuda(), 1)
    with torch.no_grad():  # in-place weight update must reach the engine (version counter)
        model.alignment_head.chunk_sim3_decoder.fc2.bias.add_(1.0)
    p2 = model(img.cuda(), 1)
    d = (p2["chunk_sim3_alignment_enc"] - p1["chunk_sim3_alignment_enc"]).cpu()[0, 0]
    assert torch.allclose(d[:7], torch.ones(7), atol=1e-4)
    model(img.cuda(), 1, gt_poses=torch.zeros(1, 2, 3, 4))          # first chunk: gt_poses is not looked at (:122-124)
    with pytest.raises(RuntimeError):                                # with context the reference's matmul needs (B,S,4,4)
        model(img.cuda(), 1, p2, gt_poses=torch.zeros(1, 2, 3, 4))
    # S <= num_overlap: overlap falls back to S-1 (featureAligned_vggt.py:93)
    p3 = model(img.cuda(), 5)
    assert p3["overlap_tokens"].shape[1] == 2

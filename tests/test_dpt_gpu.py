"""GPU parity of the DPT dense-prediction heads through the C ABI: the shifted-row tensor-core
convolution and the resampling kernels against torch fp32 on the same (bf16-rounded) inputs, the whole head and the
model-level depth / world_points against the oracle restatement of UPSTREAM DPTHead (oracle/functional.py).

Tolerances: single kernels 1e-2 relative (bf16 output rounding 4e-3); whole head rel-L2 <= 3e-2 on the pre-activation
logits (fp32 reference vs bf16 tensor-core convolutions through ~25 layers), stated per test."""
import pytest
import torch
import torch.nn.functional as F

from parity_util import rel_l2

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def _padded(x_nchw):
    """(N,C,h,w) fp32 -> padded NHWC bf16 (N,h+2,w+2,C) on cuda"""
    return F.pad(x_nchw, (1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous().bfloat16().cuda()


@pytest.mark.parametrize("C,OC,h,w,frames", [(64, 128, 13, 21, 2), (256, 256, 11, 37, 3), (128, 64, 20, 30, 1), (512, 256, 6, 19, 2)])
def test_conv3x3_matches_torch(C, OC, h, w, frames):
    from lsvs_b200 import ops
    g = torch.Generator().manual_seed(C + OC)
    x = _bf(torch.randn(frames, C, h, w, generator=g))
    wt = _bf(torch.randn(OC, C, 3, 3, generator=g) / (3 * C ** 0.5))
    bias = torch.randn(OC, generator=g)
    r1 = _bf(torch.randn(frames, OC, h, w, generator=g))
    r2 = _bf(torch.randn(frames, OC, h, w, generator=g))
    w2 = wt.permute(0, 2, 3, 1).reshape(OC, 9 * C).contiguous().bfloat16().cuda()
    for relu, use_res in [(False, False), (True, True)]:
        ref = F.conv2d(x, wt, bias, padding=1)
        if use_res:
            ref = ref + r1 + r2
        if relu:
            ref = F.relu(ref)
        out = ops.conv2d_nhwc(_padded(x), w2, bias.cuda(), _padded(r1) if use_res else None, _padded(r2) if use_res else None,
                              taps=9, relu=relu, mask_border=True)
        torch.cuda.synchronize()
        o = out.float().cpu()
        assert float(o[:, 0].abs().max()) == 0 and float(o[:, -1].abs().max()) == 0  # border rows / cols are zero padding
        assert float(o[:, :, 0].abs().max()) == 0 and float(o[:, :, -1].abs().max()) == 0
        got = o[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        assert rel_l2(got, ref) < 1e-2, (relu, use_res, rel_l2(got, ref))


def test_conv1x1_and_unmasked():
    from lsvs_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn(2, 256, 9, 14, generator=g))
    wt = _bf(torch.randn(256, 256, 1, 1, generator=g) / 16)
    bias = torch.randn(256, generator=g)
    out = ops.conv2d_nhwc(_padded(x), wt.reshape(256, 256).bfloat16().cuda(), bias.cuda(), taps=1, mask_border=False)
    ref = F.conv2d(F.pad(x, (1, 1, 1, 1)), wt, bias)  # unmasked: the border carries the bias
    assert rel_l2(out.float().cpu().permute(0, 3, 1, 2), ref) < 1e-2


def test_resample_kernels_match_torch():
    from lsvs_b200 import ops
    from oracle import functional as OF
    g = torch.Generator().manual_seed(7)
    frames, h, w, C = 2, 5, 7, 64
    x = _bf(torch.randn(frames, C, h, w, generator=g))
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().bfloat16().cuda()
    # pad
    out = torch.empty(frames, h + 2, w + 2, C, dtype=torch.bfloat16, device="cuda")
    ops.dpt_resample(ops.DPT_PAD, x_nhwc, out, frames, h, w, C)
    assert torch.equal(out.cpu(), _padded(x).cpu())
    # position embedding (ratio 0.1, aspect W/H of the image)
    pe = x_nhwc.clone()
    ops.dpt_resample(ops.DPT_POS_EMBED, None, pe, frames, h, w, C, aspect=518 / 154, ratio=0.1)
    ref = OF.dpt_pos_embed(x, 518, 154)
    assert float((pe.float().cpu().permute(0, 3, 1, 2) - ref).abs().max()) < 2e-2  # bf16 rounding of values ~N(0,1)
    assert rel_l2(pe.float().cpu().permute(0, 3, 1, 2) - x, ref - x) < 2e-1  # the embedding itself (0.1 amplitude, bf16 ulp 4e-3..8e-3)
    # bilinear, align_corners=True, padded in / out
    for (ho, wo) in [(10, 14), (11, 37), (7, 9)]:
        out = torch.empty(frames, ho + 2, wo + 2, C, dtype=torch.bfloat16, device="cuda")
        ops.dpt_resample(ops.DPT_BILINEAR, _padded(x), out, frames, h, w, C, a=ho, b=wo)
        ref = F.interpolate(x, size=(ho, wo), mode="bilinear", align_corners=True)
        o = out.float().cpu()
        assert float(o[:, 0].abs().max()) == 0 and float(o[:, :, -1].abs().max()) == 0
        assert float((o[:, 1:-1, 1:-1].permute(0, 3, 1, 2) - ref).abs().max()) < 2e-2
    # im2col 3x3 stride 2 pad 1 against F.unfold
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty(frames * ho * wo, 9 * C, dtype=torch.bfloat16, device="cuda")
    ops.dpt_resample(ops.DPT_IM2COL_S2, x_nhwc, out, frames, h, w, C)
    unf = F.unfold(x, 3, padding=1, stride=2).view(frames, C, 9, ho * wo).permute(0, 3, 2, 1).reshape(frames * ho * wo, 9 * C)
    assert torch.equal(out.float().cpu(), unf)
    # transposed-conv shuffle: GEMM output [(f,y,x)][(i,j,c)] -> padded NHWC
    k = 2
    wt = _bf(torch.randn(C, C, k, k, generator=g) / 8)
    gemm_out = torch.einsum("nchw,cdij->nhwijd", x, wt).reshape(frames * h * w, k * k * C)
    out = torch.empty(frames, k * h + 2, k * w + 2, C, dtype=torch.bfloat16, device="cuda")
    ops.dpt_resample(ops.DPT_CONVT_SHUFFLE, gemm_out.bfloat16().cuda().contiguous(), out, frames, h, w, C, a=k)
    ref = F.conv_transpose2d(x, wt, stride=k)
    assert float((out.float().cpu()[:, 1:-1, 1:-1].permute(0, 3, 1, 2) - ref).abs().max()) < 3e-2


def _logits(pred, conf, activation):
    """invert activate_head so that errors are measured on the network's own output scale"""
    p = pred.double().cpu()
    x = torch.log(p) if activation == "exp" else torch.sign(p) * torch.log1p(p.abs())
    return torch.cat([x, torch.log(conf.double().cpu() - 1).unsqueeze(-1)], dim=-1)


@pytest.mark.parametrize("H,W,frames,od,activation,prefix", [(56, 84, 3, 2, "exp", "depth_head."), (154, 518, 2, 4, "inv_log", "point_head."),
                                                            (154, 518, 9, 2, "exp", "depth_head.")])
def test_dpt_head_matches_oracle(H, W, frames, od, activation, prefix):
    from lsvs_b200.modules import DPTHead
    from oracle import functional as OF
    from oracle import weights as OW
    head = DPTHead(dim_in=2048, output_dim=od, activation=activation, conf_activation="expp1", prefix=prefix)
    sd = OW.fill_state_dict([(k, tuple(v.shape)) for k, v in head.state_dict().items()], seed=3)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().eval()
    P = 5 + (H // 14) * (W // 14)
    g = torch.Generator().manual_seed(11)
    taps = [torch.randn(1, frames, P, 2048, generator=g) for _ in range(4)]
    images = torch.zeros(1, frames, 3, H, W)
    with torch.no_grad():
        # 9 frames with upstream's frames_chunk_size=8 -> two passes of the frame-chunk loop; otherwise the engine's own choice
        pred, conf = head([t.cuda() for t in taps], images=images.cuda(), patch_start_idx=5, frames_chunk_size=8 if frames > 8 else None)
    torch.cuda.synchronize()
    n_ref = min(frames, 2)  # the fp32 CPU oracle on the first frames (per-frame computation)
    ref_pred, ref_conf = OF.dpt_head_forward(sd, "", [t[:, :n_ref] for t in taps], (H, W), activation=activation)
    assert pred.shape == (1, frames, H, W, od - 1) and conf.shape == (1, frames, H, W)
    assert torch.isfinite(pred).all() and torch.isfinite(conf).all()
    la, lb = _logits(pred[:, :n_ref], conf[:, :n_ref], activation), _logits(ref_pred, ref_conf, activation)
    err = rel_l2(la, lb)
    from parity_util import report
    report(f"dpt_{prefix}{H}x{W}x{frames}_precision0_logits_rel_l2", err)
    assert err < 3e-2, err
    assert rel_l2(pred[:, :n_ref], ref_pred) < 3e-2 and rel_l2(conf[:, :n_ref], ref_conf) < 3e-2
    if frames > 8:  # frames beyond the first chunk of 8 go through the second pass of the frame-chunk loop
        with torch.no_grad():
            p2, c2 = head([t[:, 8:].cuda() for t in taps], images=images[:, 8:].cuda(), patch_start_idx=5)
        assert torch.equal(p2, pred[:, 8:]) and torch.equal(c2, conf[:, 8:])


@pytest.mark.parametrize("H,W,frames,od,activation,prefix", [(56, 84, 3, 2, "exp", "depth_head."), (154, 518, 2, 4, "inv_log", "point_head.")])
def test_dpt_head_fp32_class_matches_oracle(H, W, frames, od, activation, prefix, monkeypatch):
    """precision >= 1: the DPT heads run like the reference runs them — fp32 activations (autocast disabled,
    featureAligned_vggt.py:103) — with split-bf16 GEMM operands on the tensor cores: logits within 2e-4 of the fp32 oracle
    (bf16 activations: 3e-2 above)."""
    from lsvs_b200.modules import DPTHead
    from oracle import functional as OF
    from oracle import weights as OW
    monkeypatch.setenv("LSVS_PRECISION", "1")
    head = DPTHead(dim_in=2048, output_dim=od, activation=activation, conf_activation="expp1", prefix=prefix)
    sd = OW.fill_state_dict([(k, tuple(v.shape)) for k, v in head.state_dict().items()], seed=3)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().eval()
    P = 5 + (H // 14) * (W // 14)
    g = torch.Generator().manual_seed(11)
    taps = [torch.randn(1, frames, P, 2048, generator=g) for _ in range(4)]
    images = torch.zeros(1, frames, 3, H, W)
    with torch.no_grad():
        pred, conf = head([t.cuda() for t in taps], images=images.cuda(), patch_start_idx=5)
        pred1, conf1 = head([t.cuda() for t in taps], images=images.cuda(), patch_start_idx=5, frames_chunk_size=1)
    torch.cuda.synchronize()
    assert torch.equal(pred, pred1) and torch.equal(conf, conf1)   # per-frame computation: independent of the frames per pass
    ref_pred, ref_conf = OF.dpt_head_forward(sd, "", taps, (H, W), activation=activation)
    la, lb = _logits(pred, conf, activation), _logits(ref_pred, ref_conf, activation)
    err = rel_l2(la, lb)
    from parity_util import report
    report(f"dpt_{prefix}{H}x{W}_precision1_logits_rel_l2", err)
    assert err < 2e-4, err
    assert rel_l2(pred, ref_pred) < 2e-4 and rel_l2(conf, ref_conf) < 2e-4


def test_model_depth_and_points_match_oracle():
    """FeatureAlignedVGGT with the DPT heads enabled (reduced depth): depth * scale and Sim(3)-applied world points."""
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    from oracle import aligned as OA
    from oracle import functional as OF
    from parity_util import load_synth_weights, synth_images
    model = FeatureAlignedVGGT(enable_point=True, enable_depth=True, enable_track=False, depth=1, patch_embed_depth=1,
                               intermediate_layer_indices=(0, 0, 0, 0))
    sd = load_synth_weights(model, seed=2)
    model = model.cuda().eval()
    images = synth_images(4, 1, 3, 56, 84)
    with torch.no_grad():
        out = model(images.cuda(), num_overlap=1)
    torch.cuda.synchronize()
    ref = OA.feature_aligned_forward(sd, images, 1, depth=1, dino_depth=1, taps=(0, 0, 0, 0))
    taps = ref["taps"]
    d_ref, dc_ref = OF.dpt_head_forward(sd, "depth_head.", taps, (56, 84), activation="exp")
    p_ref, pc_ref = OF.dpt_head_forward(sd, "point_head.", taps, (56, 84), activation="inv_log")
    scale = ref["chunk_sim3_alignment_enc"][..., -1].reshape(1)
    ref2 = OA.feature_aligned_forward(sd, images, 1, depth=1, dino_depth=1, taps=(0, 0, 0, 0), raw_points=p_ref, raw_depth=d_ref)
    assert set(["depth", "depth_conf", "world_points", "world_points_conf"]) <= set(out.keys())
    assert rel_l2(out["depth"][0], ref2["depth"]) < 3e-2
    assert rel_l2(out["depth_conf"][0], dc_ref) < 3e-2
    assert rel_l2(out["world_points"][0], ref2["world_points"]) < 5e-2
    assert rel_l2(out["world_points_conf"][0], pc_ref) < 3e-2
    assert float(scale) > 0


def test_model_dpt_golden(golden):
    """two chained chunks against the REFERENCE forward's depth / world_points (tests/golden/model_dpt_small.npz)."""
    import numpy as np
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    from oracle import weights as OW
    from parity_util import load_synth_weights
    g = golden("model_dpt_small.npz")
    S, H, W, ov, sub = g["S"], g["H"], g["W"], g["ov"], g["sub"]
    model = FeatureAlignedVGGT(enable_point=True, enable_depth=True, enable_track=False, depth=1, patch_embed_depth=1,
                               intermediate_layer_indices=(0, 0, 0, 0))
    sd = load_synth_weights(model, seed=2)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"]), "synthetic weights differ from the golden run"
    model = model.cuda().eval()
    p = None
    for ci in (1, 2):
        img = torch.from_numpy(np.random.Generator(np.random.PCG64(400 + ci - 1)).random((1, S, 3, H, W), dtype=np.float32))
        with torch.no_grad():
            p = model(img.cuda(), ov, p)
        assert len(p["depth"]) == ci and len(p["world_points_conf"]) == ci  # accumulated per chunk like the reference (:173-181)
        for k, tol in (("depth", 3e-2), ("depth_conf", 3e-2), ("world_points", 5e-2), ("world_points_conf", 3e-2)):
            got = p[k][-1][:, :, ::sub, ::sub]
            ref = torch.as_tensor(g[f"c{ci}_{k}"])
            assert got.shape == ref.shape, (k, got.shape, ref.shape)
            assert rel_l2(got, ref) < tol, (ci, k, rel_l2(got, ref))

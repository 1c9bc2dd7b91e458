"""Build-container script: validate the oracle restatement against the REAL reference code and write
tests/golden/*.npz (TEST INFRASTRUCTURE; needs /root/reference, so it never runs on the GPU box).

    python -m oracle.make_golden [--full]

The reference modules are imported unmodified from /root/reference on top of oracle/vggt_shim (the
un-vendored upstream).  Every golden array is an output of the reference's own code.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "vggt_shim"))
sys.path.insert(1, "/root/reference")

from oracle import aligned as OA          # noqa: E402
from oracle import functional as OF       # noqa: E402
from oracle import weights as OW          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rnd(seed, *shape, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.standard_normal(shape, dtype=np.float32) * np.float32(scale)))


def maxdiff(a, b):
    return float((a - b).abs().max())


def check(name, a, b, tol):
    d = maxdiff(a, b)
    ref = float(b.abs().max())
    print(f"  {name:38s} max|diff|={d:.3e}  (max|ref|={ref:.3e})")
    assert d <= tol * max(1.0, ref), f"{name}: restatement deviates from the reference ({d} > {tol})"


def save(fname, **arrs):
    np.savez_compressed(os.path.join(GOLD, fname), **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                                      for k, v in arrs.items()})
    print(f"  wrote tests/golden/{fname}")


# --------------------------------------------------------------------------------------------
def case_layers():
    """rope.py, gated_update.py, cross_attention.py."""
    print("[layers]")
    from aligned_vggt.layers.rope import RotaryPositionEmbedding
    from aligned_vggt.layers.gated_update import GatedUpdate
    from aligned_vggt.layers.cross_attention import CrossAttentionBlock
    x = rnd(1, 3, 8, 5, 64)
    pos = torch.tensor([[0, 3, 4, 9, 70]]).expand(3, -1)
    ref = RotaryPositionEmbedding(100.0)(x, pos)
    check("rope1d", OF.rope_apply_1d(x, pos, 100.0), ref, 1e-6)

    gu = GatedUpdate(512, 8)
    sd = OW.fill_state_dict(OW.spec_of(gu), seed=3)
    gu.load_state_dict(sd, strict=True)
    mem = torch.nn.functional.normalize(rnd(2, 2, 8, 512), dim=-1)
    upd = rnd(4, 2, 1, 512)
    ref_g = gu(mem, upd)
    check("gated_update", OA.gated_update(sd, "", mem, upd), ref_g, 1e-5)

    cb = CrossAttentionBlock(dim=512, num_heads=8, init_values=0.01, qk_norm=True, rope=RotaryPositionEmbedding(100.0))
    sdc = OW.fill_state_dict(OW.spec_of(cb), seed=5, ls_gamma=0.3)
    cb.load_state_dict(sdc, strict=True)
    xq, yk = rnd(6, 2, 5, 512), rnd(7, 2, 7, 512)
    pq = torch.tensor([[0, 1, 2, 3, 4]]).expand(2, -1)
    pk = torch.tensor([[0, 1, 2, 3, 4, 10, 11]]).expand(2, -1)
    ref_c = cb(xq, yk, pos=(pq, pk))
    check("cross_block", OA.cross_block(sdc, "", xq, yk, (pq, pk), 8), ref_c, 1e-5)
    save("layers.npz", rope_x=x, rope_pos=pos, rope_out=ref, gu_mem=mem, gu_upd=upd, gu_out=ref_g,
         cb_x=xq, cb_y=yk, cb_pq=pq, cb_pk=pk, cb_out=ref_c)


def case_geometry():
    """alignment.py Sim(3) apply, geometry.py pose average, data.py pose enc / chunks, IRLS Umeyama."""
    print("[geometry]")
    from aligned_vggt.utils import alignment as RA
    from aligned_vggt.utils import data as RD
    from aligned_vggt.utils import geometry as RG
    from aligned_vggt.models.pointAligned_wrapped_vggt import irls_sim3_umeyama, weighted_umeyama_sim3
    B, S, H, W = 2, 3, 5, 7
    pts = rnd(10, B, S, H, W, 3, scale=10.0)
    q = rnd(11, B, 4)
    R = OF.quat_to_mat(q / q.norm(dim=-1, keepdim=True))
    T = torch.eye(4).repeat(B, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = rnd(12, B, 3, scale=3.0)
    s = torch.tensor([0.7, 1.9])
    ref_p = RA.apply_sim3_alignment_on_point_maps(pts, T, s)
    check("sim3_points", OA.apply_sim3_points(pts, T, s), ref_p, 1e-6)
    qe = rnd(13, B, S, 4)
    extr = torch.cat([OF.quat_to_mat(qe), rnd(14, B, S, 3, 1, scale=2.0)], dim=-1)  # (B,S,3,4)
    ref_w = RA.apply_sim3_alignment_on_w2c(extr.clone(), T, s)
    check("sim3_w2c", OA.apply_sim3_w2c(extr, T, s), ref_w, 1e-5)
    c2w = OA.inv_se3(extr)
    ref_c = RA.apply_sim3_alignment_on_c2w(c2w.clone(), T, s)
    check("sim3_c2w", OA.apply_sim3_c2w(c2w, T, s), ref_c, 1e-6)

    enc = torch.cat([rnd(15, B, 4, 3), torch.nn.functional.normalize(rnd(16, 1, 1, 4) + 0.05 * rnd(17, B, 4, 4), dim=-1)], -1)
    ref_avg = RG.averagePoseEncodings(enc)
    mine = OA.average_pose_encodings(enc)
    sgn = torch.sign((mine[..., 3:] * ref_avg[..., 3:]).sum(-1, keepdim=True))
    check("average_pose(t)", mine[..., :3], ref_avg[..., :3], 1e-6)
    check("average_pose(q up to sign)", mine[..., 3:] * sgn, ref_avg[..., 3:], 1e-5)
    e44 = RD.pose_encoding_to_extri(enc)
    check("pose_encoding_to_extri", OA.pose_encoding_to_extri(enc), e44, 1e-6)
    back = RD.extri_to_pose_encoding(e44)
    check("extri_to_pose_encoding", OA.extri_to_pose_encoding(e44), back, 1e-6)

    # IRLS Umeyama: dst = s R src + t + noise, a few gross outliers
    n, h, w = 2, 12, 16
    src = rnd(20, n, h, w, 3, scale=4.0)
    Rg, tg, sg = R[0], T[0, :3, 3], 1.3
    dst = sg * (src @ Rg.T) + tg + rnd(21, n, h, w, 3, scale=0.01)
    dst[0, 0, :5] += 3.0
    cs, cd = 1 + torch.exp(rnd(22, n, h, w)), 1 + torch.exp(rnd(23, n, h, w))
    Rr, tr, sr = irls_sim3_umeyama(src, dst, cs, cd)
    Ro, to, so = OA.irls_umeyama(src, dst, cs, cd)
    check("irls R", Ro, Rr, 1e-5); check("irls t", to, tr, 1e-4); check("irls s", so, sr, 1e-5)
    wts = torch.sqrt(cs * cd).reshape(-1)
    Ru, tu, su = weighted_umeyama_sim3(src.reshape(-1, 3), dst.reshape(-1, 3), wts)
    Rm, tm, sm = OA.weighted_umeyama(src.reshape(-1, 3), dst.reshape(-1, 3), wts)
    check("umeyama R", Rm, Ru, 1e-5); check("umeyama t", tm, tu, 1e-4); check("umeyama s", sm, su, 1e-5)

    chunks = {}
    for (nf, mode, wd, ov) in [(1000, "chunk_overlap", 32, 8), (14, "chunk_overlap", 5, 1), (3, "chunk_overlap", 5, 1),
                               (20, "chunk_overlap", 8, 2), (17, "chunk_gt", 5, 0), (9, "all", 4, 1), (32, "chunk_overlap", 32, 8),
                               (33, "chunk_overlap", 32, 8)]:
        ref = RD.generate_chunks(nf, mode, wd, ov)
        assert OA.generate_chunks(nf, mode, wd, ov) == ref, (nf, mode, wd, ov)
        chunks[f"chunks_{nf}_{mode}_{wd}_{ov}"] = np.array([[c[0], c[-1], len(c)] for c in ref])
    print(f"  generate_chunks: {len(chunks)} cases identical (1000/32/8 -> {len(RD.generate_chunks(1000,'chunk_overlap',32,8))} chunks)")
    save("geometry.npz", pts=pts, T=T, s=s, sim3_points=ref_p, extr=extr, sim3_w2c=ref_w, c2w=c2w, sim3_c2w=ref_c,
         enc=enc, avg=ref_avg, enc_extr=e44, enc_back=back, u_src=src, u_dst=dst, u_cs=cs, u_cd=cd,
         irls_R=Rr, irls_t=tr, irls_s=sr, um_R=Ru, um_t=tu, um_s=su, **chunks)


def case_head():
    """AlignmentHead (real reference class) over two chained chunks; both temporal and global variants."""
    print("[alignment head]")
    from aligned_vggt.heads.alignment_head import AlignmentHead
    S, gh, gw, ov = 4, 4, 6, 2
    P = 5 + gh * gw
    # temporal_attention=False is unusable in the reference itself (aa_order bug, alignment_head.py:80,146)
    try:
        AlignmentHead(in_dim=2048, temporal_attention=False).eval()(rnd(1, 1, 2, 5 + 4, 2048), (28, 28), 1)
        raise SystemExit("reference temporal_attention=False unexpectedly works; restate it")
    except AttributeError as e:
        print(f"  reference temporal_attention=False fails as expected: {e}")
    for variant, temporal in (("temporal", True),):
        head = AlignmentHead(in_dim=2048, patch_size=14, num_memory_tokens=8, temporal_attention=temporal).eval()
        sd = OW.fill_state_dict(OW.spec_of(head), seed=7, ls_gamma=0.2)
        head.load_state_dict(sd, strict=True)
        tok1, tok2 = rnd(30, 1, S, P, 2048), rnd(31, 1, S, P, 2048)
        with torch.no_grad():
            r1 = head(tok1, (gh * 14, gw * 14), ov)
            r2 = head(tok2, (gh * 14, gw * 14), ov, overlap_tokens=r1[3], memory_tokens=r1[2])
            o1 = OA.alignment_head_forward(sd, "", tok1, (gh * 14, gw * 14), ov, temporal_attention=temporal)
            o2 = OA.alignment_head_forward(sd, "", tok2, (gh * 14, gw * 14), ov, o1[3], o1[2], temporal_attention=temporal)
        for nm, a, b in zip(("sim3", "se3", "memory", "overlap"), o1, r1):
            check(f"{variant} chunk1 {nm}", a, b, 2e-5)
        for nm, a, b in zip(("sim3", "se3", "memory", "overlap"), o2, r2):
            check(f"{variant} chunk2 {nm}", a, b, 2e-5)
        save(f"head_{variant}.npz", S=S, gh=gh, gw=gw, ov=ov, wsum=OW.checksum(sd),
             c1_sim3=r1[0], c1_se3=r1[1], c1_mem=r1[2], c1_overlap=r1[3],
             c2_sim3=r2[0], c2_se3=r2[1], c2_mem=r2[2], c2_overlap=r2[3])


def build_reference_model(depths=None):
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    if depths:
        os.environ["VGGT_SHIM_DEPTH"] = f"{depths[0]},{depths[1]}"
    else:
        os.environ.pop("VGGT_SHIM_DEPTH", None)
    t0 = time.time()
    with torch.device("meta"):
        model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False)
    spec = OW.spec_of(model)
    sd = OW.fill_state_dict(spec, seed=0)
    model = model.to_empty(device="cpu")
    model.load_state_dict(sd, strict=True)
    model.eval()
    print(f"  reference model built: {sum(v.numel() for v in sd.values())/1e6:.0f} M params in {time.time()-t0:.1f}s")
    return model, sd


def run_two_chunks(model, sd, S, H, W, ov, taps, depth, dino_depth, tag, sample=None):
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(100 + i)).random((1, S, 3, H, W), dtype=np.float32))
            for i in range(2)]
    model.intermediate_layer_indices = list(taps)
    captured = []
    hook = model.alignment_head.register_forward_pre_hook(lambda m, a: captured.append(a[0].detach().clone()))
    with torch.no_grad():
        t0 = time.time()
        ref1 = model(imgs[0], ov)
        t1 = time.time()
        snap1 = {k: (v[-1] if isinstance(v, list) else v).clone() for k, v in ref1.items() if k != "images"}
        ref2 = model(imgs[1], ov, ref1)
        t2 = time.time()
    hook.remove()
    print(f"  reference forward: chunk1 {t1-t0:.1f}s, chunk2 {t2-t1:.1f}s  ({torch.get_num_threads()} threads)")
    snap2 = {k: (v[-1] if isinstance(v, list) else v).clone() for k, v in ref2.items() if k != "images"}
    snap2["chunk_sim3_alignment_enc"] = ref2["chunk_sim3_alignment_enc"][:, -1:]
    snap2["frame_se3_alignment_enc"] = ref2["frame_se3_alignment_enc"][:, -(S - 1):]
    with torch.no_grad():
        o1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, depth=depth, dino_depth=dino_depth, taps=taps)
        ctx = {"overlap_tokens": o1["overlap_tokens"], "memory_tokens": o1["memory_tokens"], "pose_enc": o1["pose_enc"]}
        o2 = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, depth=depth, dino_depth=dino_depth, taps=taps)
    tol = 5e-4
    for ci, (o, r, cap) in enumerate(((o1, snap1, captured[0]), (o2, snap2, captured[1])), 1):
        check(f"{tag} c{ci} last tap", o["taps"][-1], cap, tol)
        for k in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "memory_tokens", "overlap_tokens", "pose_enc"):
            check(f"{tag} c{ci} {k}", o[k], r[k], tol)
    arrs = {"S": S, "H": H, "W": W, "ov": ov, "taps": np.array(taps), "wsum": OW.checksum(sd),
            "secs_chunk1": t1 - t0, "secs_chunk2": t2 - t1, "threads": torch.get_num_threads()}
    st = sample or 1
    for ci, (r, cap) in enumerate(((snap1, captured[0]), (snap2, captured[1])), 1):
        arrs[f"c{ci}_tap_last"] = cap[..., ::st].contiguous() if st > 1 else cap
        for k in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "memory_tokens", "pose_enc"):
            arrs[f"c{ci}_{k}"] = r[k]
        arrs[f"c{ci}_overlap_tokens"] = r["overlap_tokens"][..., ::st].contiguous() if st > 1 else r["overlap_tokens"]
    arrs["sample_stride"] = st
    save(f"model_{tag}.npz", **arrs)


def case_spec():
    """state_dict key/shape contract of the reference model classes (full depth), for the drop-in's spec test."""
    print("[state_dict spec]")
    import json
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    os.environ.pop("VGGT_SHIM_DEPTH", None)
    with torch.device("meta"):
        model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False)
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f)
    print(f"  {len(spec)} keys, {sum(int(np.prod(v)) if v else 1 for v in spec.values())/1e6:.0f} M parameters -> tests/golden/state_dict_spec.json")


def case_model_small():
    print("[FeatureAlignedVGGT, depth 2/2, S=4, 56x84, overlap 2]")
    model, sd = build_reference_model((2, 2))
    run_two_chunks(model, sd, 4, 56, 84, 2, (0, 0, 1, 1), 2, 2, "small")


def case_pose_aligned_small():
    """pose-aligned baseline (poseAligned_wrapped_vggt.VGGT, real reference class) over two chained chunks."""
    print("[pose-aligned VGGT, depth 2/2, S=4, 56x84, overlap 2]")
    from aligned_vggt.models.poseAligned_wrapped_vggt import VGGT
    os.environ["VGGT_SHIM_DEPTH"] = "2,2"
    with torch.device("meta"):
        model = VGGT(enable_point=False, enable_depth=False, enable_track=False)
    sd = OW.fill_state_dict(OW.spec_of(model), seed=0)
    model = model.to_empty(device="cpu")
    model.load_state_dict(sd, strict=True)
    model.eval()
    model.intermediate_layer_indices = [0, 0, 1, 1]
    S, H, W, ov = 4, 56, 84, 2
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(100 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i in range(2)]
    r1 = model(imgs[0], ov)
    e1 = r1["pose_enc"][-1].clone()
    r2 = model(imgs[1], ov, r1)
    e2 = r2["pose_enc"][-1].clone()
    o1 = OA.pose_aligned_forward(sd, imgs[0], ov, None, depth=2, dino_depth=2, taps=(0, 0, 1, 1))
    o2 = OA.pose_aligned_forward(sd, imgs[1], ov, {"pose_enc": o1["pose_enc"]}, depth=2, dino_depth=2, taps=(0, 0, 1, 1))
    check("pose-aligned c1 pose_enc", o1["pose_enc"], e1, 5e-4)
    check("pose-aligned c2 pose_enc", o2["pose_enc"], e2, 5e-4)
    save("model_pose_aligned_small.npz", S=S, H=H, W=W, ov=ov, wsum=OW.checksum(sd), c1_pose_enc=e1, c2_pose_enc=e2)
    # function-level fixture for the point-aligned pose update (pointAligned_wrapped_vggt.py:113-122), real reference helpers
    from aligned_vggt.utils.alignment import apply_sim3_alignment_on_w2c
    from vggt.vggt.utils.pose_enc import extri_intri_to_pose_encoding, pose_encoding_to_extri_intri
    q = rnd(11, 2, 4)
    T = torch.eye(4).repeat(2, 1, 1)
    T[:, :3, :3] = OF.quat_to_mat(q / q.norm(dim=-1, keepdim=True))
    T[:, :3, 3] = rnd(12, 2, 3, scale=3.0)
    s = torch.tensor([0.7, 1.9])
    enc = torch.cat([rnd(40, 2, 5, 3), torch.nn.functional.normalize(rnd(41, 2, 5, 4), dim=-1), 0.5 + 0.3 * torch.rand(2, 5, 2, generator=torch.Generator().manual_seed(42))], -1)
    extr, intr = pose_encoding_to_extri_intri(enc, (H, W))
    ref = extri_intri_to_pose_encoding(apply_sim3_alignment_on_w2c(extr, T, s), intr, (H, W))
    check("pose_enc_apply_sim3", OA.pose_enc_apply_sim3(enc, (H, W), T, s), ref, 1e-5)
    save("pose_enc_sim3.npz", enc=enc, T=T, s=s, H=H, W=W, out=ref)


def case_model_dpt_small():
    """FeatureAlignedVGGT with the DPT depth / point heads enabled — the REFERENCE class's own forward (depth scaling :171,
    point transform :187-207) over the shim's DPTHead — two chained chunks; stored 4x subsampled in H and W."""
    print("[FeatureAlignedVGGT + DPT heads, depth 1/1, S=3, 56x84, overlap 1]")
    import json
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    os.environ["VGGT_SHIM_DEPTH"] = "1,1"
    with torch.device("meta"):
        model = FeatureAlignedVGGT(enable_point=True, enable_depth=True, enable_track=False)
    spec = OW.spec_of(model)
    with open(os.path.join(GOLD, "state_dict_spec_dpt.json"), "w") as f:
        json.dump({k: list(v) for k, v in spec if k.startswith(("depth_head.", "point_head."))}, f)
    sd = OW.fill_state_dict(spec, seed=2)
    model = model.to_empty(device="cpu")
    model.load_state_dict(sd, strict=True)
    model.eval()
    taps = (0, 0, 0, 0)
    model.intermediate_layer_indices = list(taps)
    S, H, W, ov = 3, 56, 84, 1
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(400 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i in range(2)]
    with torch.no_grad():
        ref1 = model(imgs[0], ov)
        snap1 = {k: v[-1].clone() for k, v in ref1.items() if k in ("depth", "depth_conf", "world_points", "world_points_conf")}
        ref2 = model(imgs[1], ov, ref1)
        snap2 = {k: v[-1].clone() for k, v in ref2.items() if k in ("depth", "depth_conf", "world_points", "world_points_conf")}
        # restatement: aggregator / head / pose chain as before, DPT through oracle.functional, Sim(3) apply through oracle.aligned
        outs, ctx = [], None
        for i in range(2):
            o = OA.feature_aligned_forward(sd, imgs[i], ov, ctx, depth=1, dino_depth=1, taps=taps)
            d, dc = OF.dpt_head_forward(sd, "depth_head.", o["taps"], (H, W), activation="exp")
            p, pc = OF.dpt_head_forward(sd, "point_head.", o["taps"], (H, W), activation="inv_log")
            o = OA.feature_aligned_forward(sd, imgs[i], ov, ctx, depth=1, dino_depth=1, taps=taps, raw_points=p, raw_depth=d)
            o["depth_conf"], o["world_points_conf"] = dc, pc
            outs.append(o)
            ctx = {"overlap_tokens": o["overlap_tokens"], "memory_tokens": o["memory_tokens"], "pose_enc": o["pose_enc"]}
    arrs = {"S": S, "H": H, "W": W, "ov": ov, "wsum": OW.checksum(sd), "sub": 4}
    for ci, (o, r) in enumerate(zip(outs, (snap1, snap2)), 1):
        for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
            check(f"dpt c{ci} {k}", o[k], r[k], 5e-4)
            arrs[f"c{ci}_{k}"] = r[k][:, :, ::4, ::4].contiguous()
    save("model_dpt_small.npz", **arrs)


def _synth_gt_poses(seed, S, rows):
    """(1,S,rows,4) world-to-camera matrices of a slowly moving camera (rotation a few degrees per frame)."""
    q = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 0.0, 1.0]) + 0.05 * rnd(seed, S, 4), dim=-1)
    g = torch.eye(4).repeat(S, 1, 1)
    g[:, :3, :3] = OF.quat_to_mat(q)
    g[:, :3, 3] = rnd(seed + 1, S, 3, scale=0.5) + torch.arange(S).view(S, 1) * torch.tensor([0.3, 0.0, 1.0])
    return g[None, :, :rows].contiguous()


def case_baselines_gt_dpt():
    """The gt_poses variants (sample modes chunk_gt / two_chunks) and the two baseline wrappers WITH their DPT heads: outputs of
    the reference classes' own forwards over the shim, two chained chunks whose overlap frames are the same images."""
    print("[gt_poses paths + baseline wrappers with DPT heads, depth 1/1, S=4, 56x84, overlap 2]")
    import json
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    from aligned_vggt.models.pointAligned_wrapped_vggt import VGGT as PointVGGT
    from aligned_vggt.models.poseAligned_wrapped_vggt import VGGT as PoseVGGT
    os.environ["VGGT_SHIM_DEPTH"] = "1,1"
    taps = (0, 0, 0, 0)
    S, H, W, ov, sub = 4, 56, 84, 2, 4
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(500 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i in range(2)]
    imgs[1][:, :ov] = imgs[0][:, -ov:]
    arrs = {"S": S, "H": H, "W": W, "ov": ov, "sub": sub}

    def build(cls, seed, **kw):
        with torch.device("meta"):
            m = cls(enable_track=False, **kw)
        spec = OW.spec_of(m)
        sd = OW.fill_state_dict(spec, seed=seed)
        m = m.to_empty(device="cpu")
        m.load_state_dict(sd, strict=True)
        m.eval()
        m.intermediate_layer_indices = list(taps)
        return m, sd, spec

    # (a) feature-aligned, gt_poses (1,S,4,4) given with the second chunk
    model, sd, _ = build(FeatureAlignedVGGT, 0, enable_point=False, enable_depth=False)
    gt4 = [_synth_gt_poses(600 + 10 * i, S, 4) for i in range(2)]
    r1 = model(imgs[0], ov, None, gt_poses=gt4[0])
    e1 = r1["pose_enc"][-1].clone()
    r2 = model(imgs[1], ov, r1, gt_poses=gt4[1])
    e2 = r2["pose_enc"][-1].clone()
    o1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, depth=1, dino_depth=1, taps=taps, gt_poses=gt4[0])
    ctx = {"overlap_tokens": o1["overlap_tokens"], "memory_tokens": o1["memory_tokens"], "pose_enc": o1["pose_enc"]}
    pts = rnd(620, 1, S, H, W, 3, scale=5.0)
    o2 = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, depth=1, dino_depth=1, taps=taps, gt_poses=gt4[1], raw_points=pts)
    check("feature-aligned gt c1 pose_enc", o1["pose_enc"], e1, 5e-4)
    check("feature-aligned gt c2 pose_enc", o2["pose_enc"], e2, 5e-4)
    arrs.update(fa_wsum=OW.checksum(sd), fa_gt1=gt4[0], fa_gt2=gt4[1], fa_c1_pose_enc=e1, fa_c2_pose_enc=e2)

    # (b) pose-aligned with DPT heads: plain, and with gt_poses (1,S,3,4) on both chunks
    model, sd, spec = build(PoseVGGT, 2, enable_point=True, enable_depth=True)
    with open(os.path.join(GOLD, "state_dict_spec_baselines.json"), "w") as f:
        json.dump({k: list(v) for k, v in spec}, f)
    arrs["pa_wsum"] = OW.checksum(sd)
    gt3 = [_synth_gt_poses(640 + 10 * i, S, 3) for i in range(2)]
    keys = ("pose_enc", "depth", "depth_conf", "world_points", "world_points_conf")
    for tag, gts in (("pa", (None, None)), ("pagt", gt3)):
        r1 = model(imgs[0], ov, None, gt_poses=None if gts[0] is None else gts[0].clone())
        s1 = {k: r1[k][-1].clone() for k in keys}
        r2 = model(imgs[1], ov, r1, gt_poses=None if gts[1] is None else gts[1].clone())
        s2 = {k: r2[k][-1].clone() for k in keys}
        ctx = None
        for ci, (img, snap, gt) in enumerate(((imgs[0], s1, gts[0]), (imgs[1], s2, gts[1])), 1):
            o = OA.pose_aligned_forward(sd, img, ov, ctx, depth=1, dino_depth=1, taps=taps, gt_poses=gt)
            d, dc = OF.dpt_head_forward(sd, "depth_head.", o["taps"], (H, W), activation="exp")
            pp, pc = OF.dpt_head_forward(sd, "point_head.", o["taps"], (H, W), activation="inv_log")
            o = OA.pose_aligned_forward(sd, img, ov, ctx, depth=1, dino_depth=1, taps=taps, gt_poses=gt, raw_points=pp, raw_depth=d)
            ctx = {"pose_enc": o["pose_enc"]}
            for k, v in (("pose_enc", o["pose_enc"]), ("depth", o["depth"]), ("depth_conf", dc), ("world_points", o["world_points"]),
                         ("world_points_conf", pc)):
                check(f"{tag} c{ci} {k}", v, snap[k], 5e-4)
                arrs[f"{tag}_c{ci}_{k}"] = snap[k] if k == "pose_enc" else snap[k][:, :, ::sub, ::sub].contiguous()
            if gt is not None:
                arrs[f"{tag}_gt{ci}"] = gt
                arrs[f"{tag}_c{ci}_scale"] = o["batch_scales"]

    # (c) point-aligned with DPT heads (same weights as (b): identical parameter tree)
    model, sd_c, _ = build(PointVGGT, 2, enable_point=True, enable_depth=True)
    assert OW.checksum(sd_c) == OW.checksum(sd)
    r1 = model(imgs[0], ov, None)
    s1 = {k: r1[k][-1].clone() for k in keys}
    r2 = model(imgs[1], ov, r1)
    s2 = {k: r2[k][-1].clone() for k in keys}
    ctx = None
    for ci, (img, snap) in enumerate(((imgs[0], s1), (imgs[1], s2)), 1):
        tp = OA.pose_aligned_forward(sd, img, ov, None, depth=1, dino_depth=1, taps=taps)["taps"]
        d, dc = OF.dpt_head_forward(sd, "depth_head.", tp, (H, W), activation="exp")
        pp, pc = OF.dpt_head_forward(sd, "point_head.", tp, (H, W), activation="inv_log")
        o = OA.point_aligned_forward(sd, img, ov, ctx, raw_points=pp, raw_points_conf=pc, raw_depth=d, depth=1, dino_depth=1, taps=taps)
        ctx = {"world_points": o["world_points"], "world_points_conf": pc}
        for k, v in (("pose_enc", o["pose_enc"]), ("depth", o["depth"]), ("depth_conf", dc), ("world_points", o["world_points"]),
                     ("world_points_conf", pc)):
            check(f"point-aligned c{ci} {k}", v, snap[k], 2e-3)
            arrs[f"pt_c{ci}_{k}"] = snap[k] if k == "pose_enc" else snap[k][:, :, ::sub, ::sub].contiguous()
        arrs[f"pt_c{ci}_scale"] = o["scales"]
        arrs[f"pt_c{ci}_T"] = o["transform"]
    save("model_baselines_gt_dpt.npz", **arrs)


def case_model_ragged():
    """Chunks of different length through the REFERENCE FeatureAlignedVGGT with context: a full chunk, a shorter tail chunk
    (generate_chunks' last chunk, data.py:196-203) and a chunk no longer than num_overlap (overlap falls back to S-1, :93)."""
    print("[FeatureAlignedVGGT, ragged chunks S = 4, 3, 2 with overlap 2, depth 1/1, 56x84]")
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    os.environ["VGGT_SHIM_DEPTH"] = "1,1"
    with torch.device("meta"):
        model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False)
    sd = OW.fill_state_dict(OW.spec_of(model), seed=0)
    model = model.to_empty(device="cpu")
    model.load_state_dict(sd, strict=True)
    model.eval()
    taps = (0, 0, 0, 0)
    model.intermediate_layer_indices = list(taps)
    H, W, ov, lens = 56, 84, 2, (4, 3, 2)
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(700 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i, S in enumerate(lens)]
    arrs = {"H": H, "W": W, "ov": ov, "lens": np.array(lens), "wsum": OW.checksum(sd)}
    ref, ctx = None, None
    for ci, img in enumerate(imgs, 1):
        ref = model(img, ov, ref)
        snap = {"pose_enc": ref["pose_enc"][-1], "memory_tokens": ref["memory_tokens"][-1], "overlap_tokens": ref["overlap_tokens"],
                "chunk_sim3_alignment_enc": ref["chunk_sim3_alignment_enc"][:, -1:],
                "frame_se3_alignment_enc": ref["frame_se3_alignment_enc"][:, -(img.shape[1] - 1):]}
        o = OA.feature_aligned_forward(sd, img, ov, ctx, depth=1, dino_depth=1, taps=taps)
        ctx = {"overlap_tokens": o["overlap_tokens"], "memory_tokens": o["memory_tokens"], "pose_enc": o["pose_enc"]}
        for k, v in snap.items():
            check(f"ragged c{ci} {k}", o[k], v, 5e-4)
            arrs[f"c{ci}_{k}"] = v[..., ::8].contiguous() if k == "overlap_tokens" else v.clone()
    arrs["sample_stride"] = 8
    save("model_ragged_small.npz", **arrs)


def case_sim3_dict():
    """apply_sim3_alignment / apply_sim3_alignment_on_dict (alignment.py:428-489), real reference functions."""
    print("[apply_sim3_alignment_on_dict]")
    from aligned_vggt.utils.alignment import apply_sim3_alignment_on_dict
    B, S, H, W = 2, 3, 28, 42
    q = rnd(21, B, 4)
    T = torch.eye(4).repeat(B, 1, 1)
    T[:, :3, :3] = OF.quat_to_mat(q / q.norm(dim=-1, keepdim=True))
    T[:, :3, 3] = rnd(22, B, 3, scale=3.0)
    s = torch.tensor([0.6, 2.3])
    enc = torch.cat([rnd(23, B, S, 3), torch.nn.functional.normalize(rnd(24, B, S, 4), dim=-1),
                     0.5 + 0.3 * torch.rand(B, S, 2, generator=torch.Generator().manual_seed(25))], -1)
    pts, dep = rnd(26, B, S, H, W, 3, scale=5.0), rnd(27, B, S, H, W, 1).abs() + 0.1
    pred = {"pose_enc": enc.clone(), "world_points": pts.clone(), "depth": dep.clone()}
    apply_sim3_alignment_on_dict(pred, (H, W), T.numpy(), s.numpy())
    o = OA.apply_sim3_alignment(T, s, enc, (H, W), pts, dep)
    for k, v in zip(("pose_enc", "world_points", "depth"), o):
        check(f"sim3 dict {k}", v, pred[k], 1e-5)
    save("sim3_dict.npz", T=T, s=s, enc=enc, pts=pts, dep=dep, H=H, W=W, out_pose_enc=pred["pose_enc"], out_world_points=pred["world_points"],
         out_depth=pred["depth"])


def case_eval_geometry():
    """unproject_depth_map_to_point_map, scale_align_from_depths, convertDictListsToTensors of the reference."""
    print("[evaluation-side geometry]")
    from aligned_vggt.utils.geometry import unproject_depth_map_to_point_map
    from aligned_vggt.utils.alignment import scale_align_from_depths
    from aligned_vggt.utils.data import convertDictListsToTensors
    B, S, H, W = 2, 3, 10, 14
    depth = rnd(501, B, S, H, W, 1).abs() + 0.5
    q = torch.nn.functional.normalize(rnd(502, B, S, 4), dim=-1)
    extr = torch.cat([OF.quat_to_mat(q), rnd(503, B, S, 3, 1)], dim=-1)
    intr = torch.zeros(B, S, 3, 3)
    intr[..., 0, 0] = 20.0 + rnd(504, B, S).abs(); intr[..., 1, 1] = 22.0 + rnd(505, B, S).abs()
    intr[..., 0, 2] = W / 2; intr[..., 1, 2] = H / 2; intr[..., 2, 2] = 1.0; intr[..., 0, 1] = 0.1
    ref = unproject_depth_map_to_point_map(depth, extr, intr)
    check("unproject", OA.unproject_depth(depth, extr, intr), ref, 2e-6)
    # scale alignment: prediction = gt / true_scale * noise, with outliers, masked pixels, a zero and a negative prediction
    gt = rnd(506, B, S, H, W, 1).abs() * 3 + 0.2
    true_scale = torch.tensor([1.7, 0.6]).view(B, 1, 1, 1, 1)
    pred = gt / true_scale * (1 + 0.05 * rnd(507, B, S, H, W, 1))
    pred[0, 0, 0, :5] *= 8.0
    pred[1, 1, 2, 3] = 0.0
    pred[1, 2, 4, 5] *= -1
    mask = (rnd(508, B, S, H, W) > -1.0)
    conf = 1 + rnd(509, B, S, H, W).exp()
    pts = rnd(510, B, S, H, W, 3)
    pose = rnd(511, B, S, 9)
    preds = {"depth": pred.clone(), "depth_conf": conf.clone(), "world_points": pts.clone(), "pose_enc": pose.clone()}
    scale_align_from_depths(preds, {"depths": gt, "point_masks": mask})
    sc = torch.tensor(preds["alignment_scales"])
    mine = OA.depth_scale_align(pred.reshape(B, -1), gt.reshape(B, -1), mask.reshape(B, -1), conf.reshape(B, -1))
    check("depth scale", mine, sc, 1e-6)
    print("   scales", sc.tolist())
    # list merge
    chunks = {"depth": [rnd(520 + i, 1, 4, 2, 2, 1) for i in range(3)], "pose_enc": [rnd(530 + i, 1, 4, 9) for i in range(3)], "other": [1, 2, 3]}
    merged = {k: ([t.clone() for t in v] if torch.is_tensor(v[0]) else v) for k, v in chunks.items()}
    convertDictListsToTensors(merged, 1)
    mine = OA.convert_dict_lists(chunks, 1)
    check("merge depth", mine["depth"], merged["depth"], 0)
    check("merge pose", mine["pose_enc"], merged["pose_enc"], 0)
    save("eval_geometry.npz", depth=depth, extr=extr, intr=intr, unproj=ref, gt=gt, pred=pred, mask=mask.float(), conf=conf, pts=pts, pose=pose,
         scales=sc, depth_aligned=preds["depth"], pts_aligned=preds["world_points"], pose_aligned=preds["pose_enc"],
         merged_depth=merged["depth"], merged_pose=merged["pose_enc"])


def case_host_glue():
    """Host-side glue of data.py / geometry.py / alignment.py that training/{run_model,training_metrics,loss}.py import
    (tests/test_dropin_surface.py checks the drop-in's own versions against these reference outputs)."""
    print("[host glue: data.py / geometry.py / alignment.py]")
    import random
    from aligned_vggt.utils import alignment as RA
    from aligned_vggt.utils import data as RD
    from aligned_vggt.utils import geometry as RG
    out = {}
    B, S, H, W = 2, 5, 6, 8
    q = rnd(40, B, S, 4)
    extr = torch.cat([OF.quat_to_mat(q), rnd(41, B, S, 3, 1, scale=2.0)], dim=-1)          # (B,S,3,4) world-to-camera
    out["extr"] = extr
    out["rel_next"] = RG.compute_relative_poses(extr)
    out["rel_prev3"] = RG.compute_relative_poses(extr, 3, False)
    wp = rnd(42, B, S, H, W, 3, scale=4.0)
    K = torch.zeros(B, S, 3, 3)
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = 300.0, 320.0, W / 2, H / 2, 1.0
    pix, valid = RG.project_world_points_to_pixels(wp, extr, K)
    out.update(wp=wp, K=K, pix=pix, pix_valid=valid.float())
    # data.py: normalisation of a data-loader batch
    masks = (rnd(43, B, S, H, W) > -0.5)
    cam_pts, depths = rnd(44, B, S, H, W, 3, scale=3.0), rnd(45, B, S, H, W).abs() + 0.5
    ne, nc, nw, nd = RD.normalize_camera_extrinsics_and_points_batch(extr, cam_pts, wp, depths, True, masks)
    out.update(masks=masks.float(), cam_pts=cam_pts, depths=depths, norm_extr=ne, norm_cam=nc, norm_world=nw, norm_depths=nd)
    ne2, _, nw2, _ = RD.normalize_camera_extrinsics_and_points_batch(extr, cam_pts, wp, depths, False, masks)
    out.update(norm_extr_noscale=ne2, norm_world_noscale=nw2)
    # chunk_batch + two_chunks
    batch = {"images": rnd(46, B, 11, 3, 4, 4), "ids": torch.arange(B * 11).view(B, 11), "name": "not a tensor"}
    idx = RD.generate_chunks(11, "chunk_overlap", 5, 1)
    cb = RD.chunk_batch(batch, idx)
    assert sorted(cb.keys()) == ["ids", "images"]
    out["chunk_ids_last"] = cb["ids"][-1]
    random.seed(7)
    tc = [RD.generate_chunks(n, "two_chunks", 4, 1) for n in (2, 3, 9, 9)]
    out["two_chunks_flat"] = np.array([i for chunks in tc for c in chunks for i in c + [-1]])
    # alignment.py: closed-form solvers
    g = np.random.Generator(np.random.PCG64(47))
    x = g.standard_normal((3, 40))
    Rg = OF.quat_to_mat(rnd(48, 4)).double().numpy()
    y = 1.7 * Rg @ x + np.array([[0.3], [-1.0], [2.0]]) + 0.01 * g.standard_normal((3, 40))
    r, t, c = RA.umeyama(x, y)
    rh, th, sh = RA.methodOfHorn(x, y)
    rh1, th1, sh1 = RA.methodOfHorn(x, y, align_scale=False)
    out.update(um_x=x, um_y=y, um_r=r, um_t=t, um_c=c, horn_r=rh, horn_t=th, horn_s=sh, horn_t_noscale=th1,
               lse=RA.scale_lse_solver(x.T, y.T))
    # alignment.py: ground-truth scale aligners (in place on clones)
    def preds():
        return {"pose_enc": torch.cat([rnd(49, B, S, 3), rnd(50, B, S, 4), torch.full((B, S, 2), 0.8)], -1),
                "depth": rnd(51, B, S, H, W, 1).abs() + 0.3, "world_points": rnd(52, B, S, H, W, 3)}
    gtb = {"extrinsics": extr}
    for name, fn in (("sfp", lambda p: RA.scale_alignment_from_poses(p, gtb)), ("sfp3", lambda p: RA.scale_alignment_from_poses(p, gtb, 3)),
                     ("pfs", lambda p: RA.per_frame_scale_alignment_from_poses(p, gtb))):
        p = preds()
        fn(p)
        out.update({f"{name}_pose": p["pose_enc"], f"{name}_depth": p["depth"], f"{name}_points": p["world_points"],
                    f"{name}_scales": np.array(p["alignment_scales"], dtype=np.float64)})
    pc = {k: [v[:, :3].clone(), v[:, 2:].clone()] for k, v in preds().items()}
    RA.per_chunk_scale_alignment_from_poses(pc, {"extrinsics": [extr[:, :3], extr[:, 2:]]})
    out.update(pcs_pose1=pc["pose_enc"][1], pcs_depth0=pc["depth"][0], pcs_scales=torch.stack(pc["alignment_scales_per_chunk"]))
    conf = 1 + torch.exp(rnd(53, B, S, H, W))
    tgt = 0.9 * wp @ torch.from_numpy(Rg).float().T + 0.5
    Tp, cp = RA.umeyama_alignment_from_points(wp[:, :3], conf[:, :3], tgt[:, :3], masks[:, :3], confidence_threshold=50.0)
    out.update(ufp_conf=conf, ufp_tgt=tgt, ufp_T=Tp, ufp_c=cp)
    save("host_glue.npz", **out)


def case_model_full():
    print("[FeatureAlignedVGGT, full depth, config 1: S=4, 154x518, overlap 1]")
    model, sd = build_reference_model(None)
    run_two_chunks(model, sd, 4, 154, 518, 1, (4, 11, 17, 23), 24, 24, "full", sample=8)


def case_model_headline():
    """The benchmarked configuration (BASELINE configs[1]/[2]): S=32 frames of 154x518, overlap 8, full depth, THREE chained chunks
    (first chunk, context chunk, context + carried memory) through the reference's FeatureAlignedVGGT.  Token tensors are stored
    column-subsampled (tap: every 64th of 2048 columns, overlap tokens: every 32nd of 1024)."""
    print("[FeatureAlignedVGGT, full depth, headline config: S=32, 154x518, overlap 8, 3 chunks]")
    model, sd = build_reference_model(None)
    S, H, W, ov, taps, n = 32, 154, 518, 8, (4, 11, 17, 23), 3
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(100 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i in range(n)]
    captured = []
    hook = model.alignment_head.register_forward_pre_hook(lambda m, a: captured.append(a[0].detach().clone()))
    snaps, secs, pred = [], [], None
    with torch.no_grad():
        for i in range(n):
            t0 = time.time()
            pred = model(imgs[i], ov, pred)
            secs.append(time.time() - t0)
            snap = {k: (v[-1] if isinstance(v, list) else v).clone() for k, v in pred.items() if k != "images"}
            snap["chunk_sim3_alignment_enc"] = pred["chunk_sim3_alignment_enc"][:, -1:].clone()
            snap["frame_se3_alignment_enc"] = pred["frame_se3_alignment_enc"][:, -(S - 1):].clone()
            snaps.append(snap)
            for k in ("images",):           # the reference keeps every chunk's images in the context: drop them (memory)
                pred.pop(k, None)
            print(f"  reference chunk {i + 1}: {secs[-1]:.1f}s ({torch.get_num_threads()} threads)", flush=True)
    hook.remove()
    ctx = None
    with torch.no_grad():
        for i in range(n):
            o = OA.feature_aligned_forward(sd, imgs[i], ov, ctx, taps=taps)
            ctx = {"overlap_tokens": o["overlap_tokens"], "memory_tokens": o["memory_tokens"], "pose_enc": o["pose_enc"]}
            check(f"headline c{i + 1} last tap", o["taps"][-1], captured[i], 5e-4)
            for k in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "memory_tokens", "overlap_tokens", "pose_enc"):
                check(f"headline c{i + 1} {k}", o[k], snaps[i][k], 5e-4)
    arrs = {"S": S, "H": H, "W": W, "ov": ov, "taps": np.array(taps), "n_chunks": n, "wsum": OW.checksum(sd), "tap_stride": 64,
            "overlap_stride": 32, "secs_per_chunk": np.array(secs), "threads": torch.get_num_threads()}
    for i in range(n):
        c = f"c{i + 1}"
        arrs[c + "_tap_last"] = captured[i][..., ::64].contiguous()
        arrs[c + "_tap_last_norm"] = float(captured[i].norm())
        arrs[c + "_overlap_tokens"] = snaps[i]["overlap_tokens"][..., ::32].contiguous()
        for k in ("chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "memory_tokens", "pose_enc"):
            arrs[f"{c}_{k}"] = snaps[i][k]
    save("model_headline.npz", **arrs)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run the full-depth config-1 case (minutes, ~10 GB RAM)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.manual_seed(0)
    os.makedirs(GOLD, exist_ok=True)
    cases = {"spec": case_spec, "layers": case_layers, "geometry": case_geometry, "head": case_head, "model_small": case_model_small,
             "pose_aligned": case_pose_aligned_small, "model_dpt_small": case_model_dpt_small, "eval_geometry": case_eval_geometry,
             "baselines_gt_dpt": case_baselines_gt_dpt, "sim3_dict": case_sim3_dict, "model_ragged": case_model_ragged,
             "host_glue": case_host_glue}
    if args.full:
        cases["model_full"] = case_model_full
        cases["model_headline"] = case_model_headline
    for name, fn in cases.items():
        if args.only and name != args.only:
            continue
        fn()
    print("oracle restatement matches the reference on all cases")

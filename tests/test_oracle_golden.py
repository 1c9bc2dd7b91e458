"""CPU: the oracle restatement reproduces the golden vectors that oracle/make_golden.py generated from the
REAL reference code (/root/reference, which does not exist on the GPU box)."""
import re

import numpy as np
import pytest
import torch

from conftest import rnd, rand01
from oracle import aligned as OA
from oracle import functional as OF
from oracle import weights as OW

torch.set_grad_enabled(False)


def close(a, b, tol):
    a, b = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)
    d = float((a - b).abs().max())
    assert d <= tol * max(1.0, float(b.abs().max())), d


def test_layers_golden(golden):
    g = golden("layers.npz")
    close(OF.rope_apply_1d(g["rope_x"], g["rope_pos"], 100.0), g["rope_out"], 1e-6)
    shapes = [(f"delta_mlps.{i}.{j}.{w}", s) for i in range(8)
              for j, w, s in ((0, "weight", (512, 1536)), (0, "bias", (512,)), (2, "weight", (512, 512)), (2, "bias", (512,)))]
    shapes += [("gate_mlp.0.weight", (512, 1024)), ("gate_mlp.0.bias", (512,)), ("gate_mlp.2.weight", (1, 512)),
               ("gate_mlp.2.bias", (1,))]
    sd = OW.fill_state_dict(shapes, seed=3)
    out = OA.gated_update(sd, "", g["gu_mem"], g["gu_upd"])
    close(out, g["gu_out"], 1e-5)
    close(out.norm(dim=-1), torch.ones(2, 8), 1e-5)  # KAT: rows stay unit norm (gated_update.py:77)


def test_geometry_golden(golden):
    g = golden("geometry.npz")
    close(OA.apply_sim3_points(g["pts"], g["T"], g["s"]), g["sim3_points"], 1e-6)
    close(OA.apply_sim3_w2c(g["extr"], g["T"], g["s"]), g["sim3_w2c"], 1e-5)
    close(OA.apply_sim3_c2w(g["c2w"], g["T"], g["s"]), g["sim3_c2w"], 1e-6)
    avg = OA.average_pose_encodings(g["enc"])
    sgn = torch.sign((avg[..., 3:] * g["avg"][..., 3:]).sum(-1, keepdim=True))
    close(avg[..., :3], g["avg"][..., :3], 1e-6)
    close(avg[..., 3:] * sgn, g["avg"][..., 3:], 1e-5)
    close(OA.pose_encoding_to_extri(g["enc"]), g["enc_extr"], 1e-6)
    close(OA.extri_to_pose_encoding(g["enc_extr"]), g["enc_back"], 1e-6)
    R, t, s = OA.irls_umeyama(g["u_src"], g["u_dst"], g["u_cs"], g["u_cd"])
    close(R, g["irls_R"], 1e-5), close(t, g["irls_t"], 1e-4), close(s, g["irls_s"], 1e-5)
    # robust fit recovers the generating transform despite the outliers
    assert abs(float(s) - 1.3) < 5e-3


def test_geometry_kats():
    # identity Sim(3) leaves points unchanged (alignment.py:513-526)
    pts = rnd(1, 1, 2, 3, 4, 3)
    close(OA.apply_sim3_points(pts, torch.eye(4)[None], torch.ones(1)), pts, 0)
    # pose enc round trip (data.py:12-52)
    q = torch.nn.functional.normalize(rnd(2, 1, 5, 4), dim=-1)
    q = torch.where(q[..., 3:] < 0, -q, q)
    enc = torch.cat([rnd(3, 1, 5, 3), q], -1)
    close(OA.extri_to_pose_encoding(OA.pose_encoding_to_extri(enc)), enc, 1e-5)
    # average of N identical poses is that pose up to sign (geometry.py:4-37)
    same = enc[:, :1].expand(1, 6, 7)
    avg = OA.average_pose_encodings(same)
    assert min(float((avg[0, 0, 3:] - q[0, 0]).abs().max()), float((avg[0, 0, 3:] + q[0, 0]).abs().max())) < 1e-5
    # Umeyama recovers a noise-free (R,t,s) (pointAligned_wrapped_vggt.py:159-217)
    src = rnd(4, 200, 3)
    Rg = OF.quat_to_mat(q[0, 1])
    dst = 0.7 * src @ Rg.T + torch.tensor([1.0, -2.0, 0.5])
    R, t, s = OA.weighted_umeyama(src, dst, torch.ones(200))
    close(R, Rg, 1e-5), close(t, torch.tensor([1.0, -2.0, 0.5]), 1e-5)
    assert abs(float(s) - 0.7) < 1e-5
    with pytest.raises(ValueError):
        OA.weighted_umeyama(src, dst, torch.zeros(200))


def test_generate_chunks_golden(golden):
    g = golden("geometry.npz")
    n = 0
    for k, v in g.items():
        if not k.startswith("chunks_"):
            continue
        nf, mode, wd, ov = re.match(r"chunks_(\d+)_(.+)_(\d+)_(\d+)$", k).groups()
        wd, ov = int(wd), int(ov)
        got = OA.generate_chunks(int(nf), mode, wd, ov)
        assert [[c[0], c[-1], len(c)] for c in got] == v.tolist(), k
        n += 1
    assert n == 8
    ch = OA.generate_chunks(1000, "chunk_overlap", 32, 8)
    assert len(ch) == 42 and ch[-1][0] == 984 and ch[-1][-1] == 999  # SURVEY §4 KAT
    with pytest.raises(ValueError):
        OA.generate_chunks(10, "bogus", 4, 1)


HEAD_SPEC_CACHE = {}


def head_spec():
    """state_dict spec of the alignment head, written out (alignment_head.py:100-221)."""
    if HEAD_SPEC_CACHE:
        return HEAD_SPEC_CACHE["s"]
    s = [("per_frame_alignment_token", (1, 2, 1, 1024)), ("memory_token", (1, 8, 512)), ("alpha", ()),
         ("project_in.weight", (1024, 2048)), ("project_in.bias", (1024,)),
         ("project_dec.weight", (512, 1024)), ("project_dec.bias", (512,))]

    def lin(n, o, i):
        return [(n + ".weight", (o, i)), (n + ".bias", (o,))]

    def norm(n, d):
        return [(n + ".weight", (d,)), (n + ".bias", (d,))]

    for i in range(4):
        b = f"frame_blocks.{i}."
        s += norm(b + "norm1", 1024) + lin(b + "attn.qkv", 3072, 1024) + norm(b + "attn.q_norm", 128) + norm(b + "attn.k_norm", 128)
        s += lin(b + "attn.proj", 1024, 1024) + [(b + "ls1.gamma", (1024,))] + norm(b + "norm2", 1024)
        s += lin(b + "mlp.fc1", 4096, 1024) + lin(b + "mlp.fc2", 1024, 4096) + [(b + "ls2.gamma", (1024,))]
    for grp, d in (("temporal_blocks", 1024), ("chunk_cross_blocks", 512), ("frame_cross_blocks", 512)):
        for i in range(4 if d == 1024 else 2):
            b = f"{grp}.{i}."
            hd = d // 8
            s += norm(b + "norm1", d) + lin(b + "attn.q", d, d) + lin(b + "attn.k", d, d) + lin(b + "attn.v", d, d)
            s += norm(b + "attn.q_norm", hd) + norm(b + "attn.k_norm", hd) + lin(b + "attn.proj", d, d)
            s += [(b + "ls1.gamma", (d,))] + norm(b + "norm2", d) + lin(b + "mlp.fc1", 4 * d, d) + lin(b + "mlp.fc2", d, 4 * d)
            s += [(b + "ls2.gamma", (d,))] + norm(b + "norm3", d)
    s += lin("chunk_sim3_decoder.fc1", 256, 512) + lin("chunk_sim3_decoder.fc2", 8, 256)
    s += lin("frame_se3_decoder.fc1", 256, 512) + lin("frame_se3_decoder.fc2", 7, 256)
    s += norm("token_norm", 1024) + norm("dec_norm", 512) + norm("chunk_norm", 512) + norm("frame_norm", 512)
    s += lin("frame_proj", 4096, 512)
    for i in range(8):
        s += lin(f"gated_update.delta_mlps.{i}.0", 512, 1536) + lin(f"gated_update.delta_mlps.{i}.2", 512, 512)
    s += lin("gated_update.gate_mlp.0", 512, 1024) + lin("gated_update.gate_mlp.2", 1, 512)
    HEAD_SPEC_CACHE["s"] = s
    return s


def test_alignment_head_golden(golden):
    g = golden("head_temporal.npz")
    sd = OW.fill_state_dict(head_spec(), seed=7, ls_gamma=0.2)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"]), "synthetic weights differ from the golden run"
    S, gh, gw, ov = g["S"], g["gh"], g["gw"], g["ov"]
    P = 5 + gh * gw
    tok1, tok2 = rnd(30, 1, S, P, 2048), rnd(31, 1, S, P, 2048)
    o1 = OA.alignment_head_forward(sd, "", tok1, (gh * 14, gw * 14), ov)
    o2 = OA.alignment_head_forward(sd, "", tok2, (gh * 14, gw * 14), ov, o1[3], o1[2])
    for a, k in zip(o1, ("c1_sim3", "c1_se3", "c1_mem", "c1_overlap")):
        close(a, g[k], 2e-5)
    for a, k in zip(o2, ("c2_sim3", "c2_se3", "c2_mem", "c2_overlap")):
        close(a, g[k], 2e-5)
    assert o1[0].shape == (1, 1, 8) and o1[1].shape == (1, S - 1, 7) and o1[2].shape == (1, 8, 512)
    assert o1[3].shape == (1, 1 + ov, P + 1, 1024)
    with pytest.raises(AttributeError):  # the reference's own temporal_attention=False path is broken
        OA.alignment_head_forward(sd, "", tok1, (gh * 14, gw * 14), ov, temporal_attention=False)


def test_dpt_model_golden(golden):
    """oracle DPT restatement + Sim(3) application vs the reference forward's depth / world_points (model_dpt_small.npz:
    reference FeatureAlignedVGGT :166-207 over the shim DPTHead), two chained chunks."""
    import numpy as np
    g = golden("model_dpt_small.npz")
    S, H, W, ov, sub = g["S"], g["H"], g["W"], g["ov"], g["sub"]
    from lsvs_b200 import specs
    spec = [("aggregator." + n, s) for n, s in specs.aggregator_spec(1, 1)] + [("camera_head." + n, s) for n, s in specs.camera_head_spec()] + \
           [("alignment_head." + n, s) for n, s in specs.alignment_head_spec()] + \
           [("point_head." + n, s) for n, s in specs.dpt_head_spec(2048, 4)] + [("depth_head." + n, s) for n, s in specs.dpt_head_spec(2048, 2)]
    sd = OW.fill_state_dict(spec, seed=2)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"])
    ctx = None
    for ci in (1, 2):
        img = torch.from_numpy(np.random.Generator(np.random.PCG64(400 + ci - 1)).random((1, S, 3, H, W), dtype=np.float32))
        o = OA.feature_aligned_forward(sd, img, ov, ctx, depth=1, dino_depth=1, taps=(0, 0, 0, 0))
        d, dc = OF.dpt_head_forward(sd, "depth_head.", o["taps"], (H, W), activation="exp")
        p, pc = OF.dpt_head_forward(sd, "point_head.", o["taps"], (H, W), activation="inv_log")
        o = OA.feature_aligned_forward(sd, img, ov, ctx, depth=1, dino_depth=1, taps=(0, 0, 0, 0), raw_points=p, raw_depth=d)
        for k, v in (("depth", o["depth"]), ("depth_conf", dc), ("world_points", o["world_points"]), ("world_points_conf", pc)):
            close(v[:, :, ::sub, ::sub], g[f"c{ci}_{k}"], 5e-4)
        ctx = {"overlap_tokens": o["overlap_tokens"], "memory_tokens": o["memory_tokens"], "pose_enc": o["pose_enc"]}


def test_eval_geometry_golden(golden):
    """SURVEY §8f rank 3 restatements vs outputs of the reference's unproject_depth_map_to_point_map / scale_align_from_depths /
    convertDictListsToTensors (tests/golden/eval_geometry.npz)."""
    g = golden("eval_geometry.npz")
    t = lambda k: torch.as_tensor(g[k])
    close(OA.unproject_depth(t("depth"), t("extr"), t("intr")), g["unproj"], 2e-6)
    B = t("pred").shape[0]
    sc = OA.depth_scale_align(t("pred").reshape(B, -1), t("gt").reshape(B, -1), t("mask").reshape(B, -1), t("conf").reshape(B, -1))
    close(sc, g["scales"], 1e-6)
    close(t("pred") * sc.view(B, 1, 1, 1, 1), g["depth_aligned"], 1e-6)


def _gt_dpt_images(g):
    S, H, W, ov = g["S"], g["H"], g["W"], g["ov"]
    imgs = [torch.from_numpy(np.random.Generator(np.random.PCG64(500 + i)).random((1, S, 3, H, W), dtype=np.float32)) for i in range(2)]
    imgs[1][:, :ov] = imgs[0][:, -ov:]
    return imgs


def test_gt_poses_and_baselines_with_dpt_golden(golden):
    """gt_poses variants (featureAligned_vggt.py:123-124, poseAligned_wrapped_vggt.py:84-109,:144-169) and the baseline wrappers
    with their DPT heads (pose-aligned :132-195, point-aligned :69-138) vs the reference classes' own outputs."""
    g = golden("model_baselines_gt_dpt.npz")
    S, H, W, ov, sub = g["S"], g["H"], g["W"], g["ov"], g["sub"]
    imgs = _gt_dpt_images(g)
    from lsvs_b200 import specs
    base = [("aggregator." + n, s) for n, s in specs.aggregator_spec(1, 1)] + [("camera_head." + n, s) for n, s in specs.camera_head_spec()]
    kw = dict(depth=1, dino_depth=1, taps=(0, 0, 0, 0))
    # (a) feature-aligned with gt_poses (B,S,4,4)
    sd = OW.fill_state_dict(base + [("alignment_head." + n, s) for n, s in specs.alignment_head_spec()], seed=0)
    assert abs(OW.checksum(sd) - g["fa_wsum"]) < 1e-6 * abs(g["fa_wsum"])
    o1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, gt_poses=g["fa_gt1"], **kw)
    ctx = {"overlap_tokens": o1["overlap_tokens"], "memory_tokens": o1["memory_tokens"], "pose_enc": o1["pose_enc"]}
    o2 = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, gt_poses=g["fa_gt2"], **kw)
    close(o1["pose_enc"], g["fa_c1_pose_enc"], 5e-4)
    close(o2["pose_enc"], g["fa_c2_pose_enc"], 5e-4)
    o2_plain = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, **kw)
    assert float((o2_plain["pose_enc"] - o2["pose_enc"]).abs().max()) > 1e-2  # the gt transform really replaces the averaged one
    # (b) pose-aligned with DPT heads, without and with gt_poses (B,S,3,4); (c) point-aligned with DPT heads
    sd = OW.fill_state_dict(base + [("point_head." + n, s) for n, s in specs.dpt_head_spec(2048, 4)]
                            + [("depth_head." + n, s) for n, s in specs.dpt_head_spec(2048, 2)], seed=2)
    assert abs(OW.checksum(sd) - g["pa_wsum"]) < 1e-6 * abs(g["pa_wsum"])
    dpt = []
    for img in imgs:  # the heads depend only on the chunk's own images
        tp = OA.pose_aligned_forward(sd, img, ov, None, **kw)["taps"]
        dpt.append(OF.dpt_head_forward(sd, "depth_head.", tp, (H, W), activation="exp") + OF.dpt_head_forward(sd, "point_head.", tp, (H, W), activation="inv_log"))
    for tag in ("pa", "pagt"):
        ctx = None
        for ci in (1, 2):
            d, dc, pp, pc = dpt[ci - 1]
            gt = g[f"pagt_gt{ci}"] if tag == "pagt" else None
            o = OA.pose_aligned_forward(sd, imgs[ci - 1], ov, ctx, gt_poses=gt, raw_points=pp, raw_depth=d, **kw)
            ctx = {"pose_enc": o["pose_enc"]}
            close(o["pose_enc"], g[f"{tag}_c{ci}_pose_enc"], 5e-4)
            close(o["depth"][:, :, ::sub, ::sub], g[f"{tag}_c{ci}_depth"], 5e-4)
            close(o["world_points"][:, :, ::sub, ::sub], g[f"{tag}_c{ci}_world_points"], 5e-4)
            if gt is not None:
                close(o["batch_scales"], g[f"pagt_c{ci}_scale"], 1e-5)
    ctx = None
    for ci in (1, 2):
        d, dc, pp, pc = dpt[ci - 1]
        o = OA.point_aligned_forward(sd, imgs[ci - 1], ov, ctx, raw_points=pp, raw_points_conf=pc, raw_depth=d, **kw)
        ctx = {"world_points": o["world_points"], "world_points_conf": pc}
        close(o["pose_enc"], g[f"pt_c{ci}_pose_enc"], 2e-3)
        close(o["world_points"][:, :, ::sub, ::sub], g[f"pt_c{ci}_world_points"], 2e-3)
        close(o["depth"][:, :, ::sub, ::sub], g[f"pt_c{ci}_depth"], 2e-3)
        close(o["scales"], g[f"pt_c{ci}_scale"], 1e-4)


def test_sim3_dict_golden(golden):
    """apply_sim3_alignment (alignment.py:449-489) restatement vs the reference function's outputs."""
    g = golden("sim3_dict.npz")
    pose, pts, dep = OA.apply_sim3_alignment(g["T"], g["s"], g["enc"], (g["H"], g["W"]), g["pts"], g["dep"])
    close(pose, g["out_pose_enc"], 1e-5)
    close(pts, g["out_world_points"], 1e-5)
    close(dep, g["out_depth"], 1e-6)


def test_ragged_chunks_golden(golden):
    """Chunks of 4, 3 and 2 frames chained with overlap 2 (short tail chunk; S <= num_overlap -> overlap S-1,
    featureAligned_vggt.py:93) vs the reference model's outputs."""
    g = golden("model_ragged_small.npz")
    from lsvs_b200 import specs
    spec = [("aggregator." + n, s) for n, s in specs.aggregator_spec(1, 1)] + [("camera_head." + n, s) for n, s in specs.camera_head_spec()] + \
           [("alignment_head." + n, s) for n, s in specs.alignment_head_spec()]
    sd = OW.fill_state_dict(spec, seed=0)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"])
    H, W, ov, st = g["H"], g["W"], g["ov"], g["sample_stride"]
    ctx = None
    for ci, S in enumerate(g["lens"].tolist(), 1):
        img = torch.from_numpy(np.random.Generator(np.random.PCG64(700 + ci - 1)).random((1, S, 3, H, W), dtype=np.float32))
        o = OA.feature_aligned_forward(sd, img, ov, ctx, depth=1, dino_depth=1, taps=(0, 0, 0, 0))
        ctx = {"overlap_tokens": o["overlap_tokens"], "memory_tokens": o["memory_tokens"], "pose_enc": o["pose_enc"]}
        assert o["overlap_tokens"].shape[1] == 1 + min(ov, S - 1) and o["frame_se3_alignment_enc"].shape == (1, S - 1, 7)
        for k in ("pose_enc", "memory_tokens", "chunk_sim3_alignment_enc", "frame_se3_alignment_enc"):
            close(o[k], g[f"c{ci}_{k}"], 5e-4)
        close(o["overlap_tokens"][..., ::st], g[f"c{ci}_overlap_tokens"], 5e-4)


def test_sliced_forward_equals_monolithic():
    """oracle.aligned.feature_aligned_forward_sliced (what bench.py's CPU reference arm steps through) computes exactly what
    feature_aligned_forward does, first chunk and chunk with context."""
    import numpy as np
    from lsvs_b200 import specs
    from oracle import aligned as OA
    from oracle import weights as OW
    spec = [("aggregator." + n, s) for n, s in specs.aggregator_spec(2, 2)] + [("camera_head." + n, s) for n, s in specs.camera_head_spec()] \
        + [("alignment_head." + n, s) for n, s in specs.alignment_head_spec()]
    sd = OW.fill_state_dict(spec, seed=0)
    img = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).random((1, 3, 3, 28, 42), dtype=np.float32))
    pts = torch.randn(1, 3, 28, 42, 3)
    ctx = None
    with torch.no_grad():
        for _ in range(2):
            a = OA.feature_aligned_forward(sd, img, 1, ctx, depth=2, dino_depth=2, taps=(0, 0, 1, 1), raw_points=pts)
            gen = OA.feature_aligned_forward_sliced(sd, img, 1, ctx, depth=2, dino_depth=2, taps=(0, 0, 1, 1), raw_points=pts)
            units = []
            while True:
                try:
                    units.append(next(gen))
                except StopIteration as fin:
                    b = fin.value
                    break
            assert [u[0] for u in units] == ["patch_embed", "dino.0", "dino.1", "frame.0", "global.0", "frame.1", "global.1", "tail"]
            for k in ("pose_enc", "overlap_tokens", "memory_tokens", "chunk_sim3_alignment_enc", "frame_se3_alignment_enc", "world_points"):
                assert torch.equal(a[k], b[k]), k
            ctx = {"overlap_tokens": a["overlap_tokens"], "memory_tokens": a["memory_tokens"], "pose_enc": a["pose_enc"]}

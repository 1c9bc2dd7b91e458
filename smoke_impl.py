"""Body of __graft_entry__.smoke(): one small invocation of the hot path on cuda:0 checked against the CPU oracle:
a reduced-depth FeatureAlignedVGGT (1 DINO block, 1 frame/global pair, full-width alignment head + camera head) over
two chained 3-frame chunks, plus the Sim(3) application on a synthetic point map, plus the DPT depth head on the same taps."""
import numpy as np
import torch


def rnd(seed, *shape, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(g.standard_normal(shape, dtype=np.float32) * np.float32(scale))


def run():
    from oracle import aligned as OA
    from oracle import functional as OF
    from oracle import weights as OW
    from lsvs_b200.modules import DPTHead
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    taps = (0, 0, 0, 0)
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False, depth=1, patch_embed_depth=1,
                               intermediate_layer_indices=taps)
    sd = OW.fill_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()], seed=0)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    S, H, W, ov = 3, 28, 42, 1
    g = np.random.Generator(np.random.PCG64(5))
    imgs = [torch.from_numpy(g.random((1, S, 3, H, W), dtype=np.float32)) for _ in range(2)]
    pts = rnd(1, 1, S, H, W, 3, scale=5.0)
    p1 = model(imgs[0].to(dev), ov, None, raw_points=pts.to(dev))
    p2 = model(imgs[1].to(dev), ov, p1, raw_points=pts.to(dev))
    o1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, raw_points=pts, depth=1, dino_depth=1, taps=taps)
    ctx = {"overlap_tokens": o1["overlap_tokens"], "memory_tokens": o1["memory_tokens"], "pose_enc": o1["pose_enc"]}
    o2 = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, raw_points=pts, depth=1, dino_depth=1, taps=taps)

    def rel(a, b):
        return float((a.float().cpu() - b).norm() / b.norm())
    errs = {"overlap_tokens": rel(p2["overlap_tokens"], o2["overlap_tokens"]), "memory": rel(p2["memory_tokens"][-1], o2["memory_tokens"]),
            "sim3": rel(p2["chunk_sim3_alignment_enc"][:, -1:], o2["chunk_sim3_alignment_enc"]), "pose_enc": rel(p2["pose_enc"][-1], o2["pose_enc"]),
            "world_points": rel(p2["world_points"][-1], o2["world_points"])}
    # DPT depth head on the second chunk's oracle taps: bf16 tensor-core convolutions vs the fp32 restatement
    head = DPTHead(dim_in=2048, output_dim=2, activation="exp", conf_activation="expp1", prefix="depth_head.")
    hsd = OW.fill_state_dict([(k, tuple(v.shape)) for k, v in head.state_dict().items()], seed=1)
    head.load_state_dict(hsd, strict=True)
    head = head.to(dev).eval()
    depth, conf = head([t.to(dev) for t in o2["taps"]], images=imgs[1].to(dev), patch_start_idx=5)
    d_ref, c_ref = OF.dpt_head_forward(hsd, "", o2["taps"], (H, W), activation="exp")
    errs["dpt_log_depth"] = rel(torch.log(depth), torch.log(d_ref))
    errs["dpt_conf"] = rel(conf, c_ref)
    print("smoke rel-L2 vs CPU oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["dpt_log_depth"] < 3e-2 and errs["dpt_conf"] < 3e-2, errs
    assert errs["overlap_tokens"] < 1e-2 and errs["memory"] < 1e-2, errs
    assert errs["sim3"] < 3e-2 and errs["pose_enc"] < 3e-2 and errs["world_points"] < 3e-2, errs

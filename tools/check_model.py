"""Run the drop-in model on the GPU against the golden fixtures / oracle and print the parity metrics."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from parity_util import *

torch.set_grad_enabled(False)
def gold(name):
    with np.load(os.path.join(ROOT, "tests", "golden", name)) as z:
        return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}

def head_case():
    from aligned_vggt.heads.alignment_head import AlignmentHead
    from conftest import rnd
    g = gold("head_temporal.npz")
    head = AlignmentHead()
    sd = load_synth_weights(head, seed=7, ls_gamma=0.2)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"])
    head = head.cuda()
    S, gh, gw, ov = g["S"], g["gh"], g["gw"], g["ov"]
    P = 5 + gh * gw
    tok1, tok2 = rnd(30, 1, S, P, 2048), rnd(31, 1, S, P, 2048)
    r1 = head(tok1.cuda(), (gh * 14, gw * 14), ov)
    r2 = head(tok2.cuda(), (gh * 14, gw * 14), ov, overlap_tokens=r1[3], memory_tokens=r1[2])
    for c, r in (("c1", r1), ("c2", r2)):
        print(f"head {c}: overlap_tokens rel_l2={rel_l2(r[3], g[c+'_overlap']):.3e} memory rel_l2={rel_l2(r[2], g[c+'_mem']):.3e} "
              f"sim3 {pose_metrics(r[0], g[c+'_sim3'])} scale_rel={scalar_rel(r[0][..., 7], g[c+'_sim3'][..., 7]):.3e} se3 {pose_metrics(r[1], g[c+'_se3'])}")

def model_case(tag, depth, dino, taps, full=False):
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    g = gold(f"model_{tag}.npz")
    t0 = time.time()
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False, depth=depth, patch_embed_depth=dino,
                               intermediate_layer_indices=taps)
    sd = load_synth_weights(model, seed=0)
    assert abs(OW.checksum(sd) - g["wsum"]) < 1e-6 * abs(g["wsum"]), "weights differ from golden run"
    model = model.cuda().eval()
    print(f"model {tag}: built+loaded in {time.time()-t0:.1f}s")
    S, H, W, ov, st = g["S"], g["H"], g["W"], g["ov"], g["sample_stride"]
    imgs = [synth_images(100 + i, 1, S, H, W).cuda() for i in range(2)]
    p1 = model(imgs[0], ov)
    snap1 = {k: (v[-1] if isinstance(v, list) else v).clone() for k, v in p1.items() if k != "images"}
    tap1 = model.aggregator(imgs[0])[0][taps[-1]]
    p2 = model(imgs[1], ov, p1)
    snap2 = {k: (v[-1] if isinstance(v, list) else v).clone() for k, v in p2.items() if k != "images"}
    snap2["chunk_sim3_alignment_enc"] = p2["chunk_sim3_alignment_enc"][:, -1:]
    snap2["frame_se3_alignment_enc"] = p2["frame_se3_alignment_enc"][:, -(S - 1):]
    tap2 = model.aggregator(imgs[1])[0][taps[-1]]
    for c, snap, tap in (("c1", snap1, tap1), ("c2", snap2, tap2)):
        print(f"model {tag} {c}: tap_last rel_l2={rel_l2(tap[..., ::st], g[c+'_tap_last']):.3e} overlap rel_l2={rel_l2(snap['overlap_tokens'][..., ::st], g[c+'_overlap_tokens']):.3e} "
              f"memory rel_l2={rel_l2(snap['memory_tokens'], g[c+'_memory_tokens']):.3e}")
        print(f"   sim3 {pose_metrics(snap['chunk_sim3_alignment_enc'], g[c+'_chunk_sim3_alignment_enc'])} scale_rel={scalar_rel(snap['chunk_sim3_alignment_enc'][..., 7], g[c+'_chunk_sim3_alignment_enc'][..., 7]):.3e}")
        print(f"   se3 {pose_metrics(snap['frame_se3_alignment_enc'], g[c+'_frame_se3_alignment_enc'])}  pose_enc {pose_metrics(snap['pose_enc'], g[c+'_pose_enc'])} fov_rel={scalar_rel(snap['pose_enc'][..., 7:], g[c+'_pose_enc'][..., 7:]):.3e}")
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3): model(imgs[0], ov)
    torch.cuda.synchronize()
    print(f"model {tag}: {(time.time()-t0)/3*1e3:.2f} ms per first-chunk forward (S={S}, {H}x{W})")

if __name__ == "__main__":
    head_case()
    model_case("small", 2, 2, (0, 0, 1, 1))
    if "--full" in sys.argv:
        model_case("full", 24, 24, (4, 11, 17, 23), full=True)

"""How far does the reference's OWN shipping precision (bf16 autocast, emulated on the CPU oracle) sit from its fp32
result?  Calibrates the end-to-end Sim(3) tolerances: a bf16 tensor-core path cannot be closer to fp32 than this."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200"), os.path.join(ROOT, "tests")]
import torch
from parity_util import *
from oracle import aligned as OA
from lsvs_b200 import specs
torch.set_grad_enabled(False)

def run(depth, dino, taps, S, H, W, ov):
    spec = [("aggregator." + n, s) for n, s in specs.aggregator_spec(depth, dino)] + [("camera_head." + n, s) for n, s in specs.camera_head_spec()] + \
           [("alignment_head." + n, s) for n, s in specs.alignment_head_spec()]
    sd = OW.fill_state_dict(spec, seed=0)
    imgs = [synth_images(100 + i, 1, S, H, W) for i in range(2)]
    outs = {}
    for amp in (False, True):
        o1 = OA.feature_aligned_forward(sd, imgs[0], ov, None, depth=depth, dino_depth=dino, taps=taps, amp=amp)
        ctx = {"overlap_tokens": o1["overlap_tokens"], "memory_tokens": o1["memory_tokens"], "pose_enc": o1["pose_enc"]}
        o2 = OA.feature_aligned_forward(sd, imgs[1], ov, ctx, depth=depth, dino_depth=dino, taps=taps, amp=amp)
        outs[amp] = (o1, o2)
    for c in (0, 1):
        a, r = outs[True][c], outs[False][c]
        print(f"chunk{c+1}: tap rel_l2={rel_l2(a['taps'][-1], r['taps'][-1]):.3e} overlap={rel_l2(a['overlap_tokens'], r['overlap_tokens']):.3e} "
              f"sim3={pose_metrics(a['chunk_sim3_alignment_enc'], r['chunk_sim3_alignment_enc'])} se3={pose_metrics(a['frame_se3_alignment_enc'], r['frame_se3_alignment_enc'])} "
              f"pose={pose_metrics(a['pose_enc'], r['pose_enc'])}")

if __name__ == "__main__":
    run(2, 2, (0, 0, 1, 1), 4, 56, 84, 2)

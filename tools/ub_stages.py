"""Time of the pipeline stages of one 32-frame chunk on one GPU: encode (Aggregator + camera head, runs on the chunk's owner)
vs align (alignment head + pose chain, sequential chain on rank 0) -> the scheduler's head_cost = align / encode."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
from lsvs_b200.scheduler import ModelStages

torch.set_grad_enabled(False)
S, H, W, ov = 32, 154, 518, 8
model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False).cuda().eval()
st = ModelStages(model, ov, S, H, W, torch.device("cuda"))
imgs = torch.rand(1, S, 3, H, W, device="cuda")
tok, cam = st.encode((imgs, None, None))
packet, ctx = st.align(tok, cam, None)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
enc = timeit(lambda: st.encode((imgs, None, None)))
ali = timeit(lambda: st.align(tok, cam, ctx))
print(json.dumps({"encode_ms": round(enc, 2), "align_ms": round(ali, 2), "head_cost": round(ali / enc, 4)}))

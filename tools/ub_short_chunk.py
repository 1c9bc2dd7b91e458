"""Latency of short chunks (S = 4 / 5, the reference's shipped feature-aligned configuration) with the kernel-class split:
shows whether the step is launch-bound (sum of kernel times << step time)."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import native
from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT

lib = native.lib()
model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False).cuda().eval()
for S, ov in ((4, 1), (5, 1), (8, 2)):
    imgs = torch.rand(1, S, 3, 154, 518, device="cuda")
    with torch.no_grad():
        p = model(imgs, ov)
        for _ in range(3):
            p2 = model(imgs, ov, {k: (list(v) if isinstance(v, list) else v) for k, v in p.items()})
        torch.cuda.synchronize()
        n = 10
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            p2 = model(imgs, ov, {k: (list(v) if isinstance(v, list) else v) for k, v in p.items()})
        b.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
        ms = a.elapsed_time(b) / n
        lib.lsvs_profile_enable(1)
        p2 = model(imgs, ov, {k: (list(v) if isinstance(v, list) else v) for k, v in p.items()})
        arr = lambda t: (t * 6)()
        pms, pfl, pby, pln = arr(ctypes.c_double), arr(ctypes.c_double), arr(ctypes.c_double), arr(ctypes.c_longlong)
        lib.lsvs_profile_read(pms, pfl, pby, pln)
        lib.lsvs_profile_enable(0)
    print(json.dumps({"S": S, "ms_per_chunk": round(ms, 2), "wall_ms": round(wall, 2), "frames_per_s": round((S - ov) / ms * 1e3, 1),
                      "kernel_ms_sum": round(sum(pms), 2), "launches": sum(pln), "by_class_ms": [round(x, 2) for x in pms]}))

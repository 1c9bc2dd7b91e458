"""DPT head timing at the BASELINE chunk shape (32 frames of 154x518), with the native profiler's GEMM / element-wise split."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200 import native
from lsvs_b200.modules import DPTHead

FC = int(os.environ.get("FC", 8))
frames, H, W = int(os.environ.get("FRAMES", 32)), int(os.environ.get("H", 154)), int(os.environ.get("W", 518))
P = 5 + (H // 14) * (W // 14)
lib = native.lib()
for od, act, pre in ((2, "exp", "depth_head."), (4, "inv_log", "point_head.")):
    head = DPTHead(dim_in=2048, output_dim=od, activation=act, prefix=pre).cuda().eval()
    taps = [torch.randn(1, frames, P, 2048, device="cuda") for _ in range(4)]
    images = torch.zeros(1, frames, 3, H, W, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            head(taps, images=images, patch_start_idx=5, frames_chunk_size=FC)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 3
        for _ in range(n):
            head(taps, images=images, patch_start_idx=5, frames_chunk_size=FC)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        lib.lsvs_profile_enable(1)
        head(taps, images=images, patch_start_idx=5, frames_chunk_size=FC)
        arr = lambda t: (t * 6)()
        pms, pfl, pby, pln = arr(ctypes.c_double), arr(ctypes.c_double), arr(ctypes.c_double), arr(ctypes.c_longlong)
        lib.lsvs_profile_read(pms, pfl, pby, pln)
        lib.lsvs_profile_enable(0)
    print(json.dumps({"head": pre, "frames": frames, "hw": [H, W], "ms": round(ms, 3), "frames_per_s": round(frames / ms * 1e3, 1),
                      "gemm_ms": round(pms[0], 3), "gemm_tflops": round(pfl[0] / pms[0] / 1e9, 1), "gemm_launches": pln[0],
                      "elementwise_ms": round(pms[2], 3), "elementwise_gbps": round(pby[2] / pms[2] / 1e6, 1), "elementwise_launches": pln[2]}))

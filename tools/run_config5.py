"""BASELINE config 5: global-attention stress, one 64-frame 518x518 chunk (87 936 tokens) through the Aggregator."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch
from lsvs_b200.modules import Aggregator
torch.set_grad_enabled(False)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
with torch.device("cuda"):
    agg = Aggregator(keep_layers=[4, 11, 17, 23])
img = torch.rand(1, S, 3, 518, 518, device="cuda")
out, _ = agg(img); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out, _ = agg(img); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
P = 5 + 37 * 37
D = 1024
flops = S * (2 * 588 * D * 1369 + 72 * P * 24 * D * D + 48 * 4 * P * P * D + 24 * 4 * S * P * P * D)
t = out[23]
print(json.dumps({"config": f"{S} x 518x518 ({S * P} tokens)", "ms": ms, "frames_per_s": S / ms * 1e3, "TFLOPs": flops / ms / 1e9,
                  "finite": bool(torch.isfinite(t).all()), "tap_shape": list(t.shape), "mem_GB": torch.cuda.max_memory_allocated() / 2**30}))

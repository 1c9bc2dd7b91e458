"""Drop-in for the step right after the path in the reference's aligned_vggt/utils/geometry.py:
`unproject_depth_map_to_point_map` (:39-75) and `generate_3D_pixel_grid` (:142-158).  CUDA tensors only."""
import ctypes

import torch

from lsvs_b200 import native as _n


def generate_3D_pixel_grid(H: int, W: int, device) -> torch.Tensor:
    """(H, W, 3) homogeneous pixel coordinates (u, v, 1) — reference :142-158 (kept for callers; the kernel needs no grid)."""
    u, v = torch.meshgrid(torch.arange(W, device=device), torch.arange(H, device=device), indexing="xy")
    return torch.stack((u, v, torch.ones_like(u)), dim=-1).float()


def unproject_depth_map_to_point_map(depth_map: torch.Tensor, extrinsics: torch.Tensor, intrinsics: torch.Tensor) -> torch.Tensor:
    """depth_map (B,S,H,W,1), extrinsics (B,S,3,4) world-to-camera, intrinsics (B,S,3,3) -> world coordinates (B,S,H,W,3)."""
    B, S, H, W, _ = depth_map.shape
    for t, name in ((depth_map, "depth_map"), (extrinsics, "extrinsics"), (intrinsics, "intrinsics")):
        if not t.is_cuda:
            raise _n.NativeError(f"{name} must be a CUDA tensor (no CPU fallback on this path)")
    d = depth_map.detach().float().contiguous()
    E = extrinsics.detach().float().reshape(B * S, 3, 4).contiguous()
    K = intrinsics.detach().float().reshape(B * S, 3, 3).contiguous()
    out = torch.empty(B, S, H, W, 3, dtype=torch.float32, device=d.device)
    _n.check(_n.lib().lsvs_unproject_depth(_n.ptr(d), _n.ptr(E), _n.ptr(K), _n.ptr(out), ctypes.c_int(B * S), ctypes.c_int(H), ctypes.c_int(W),
                                          _n.stream_ptr()), "unproject_depth")
    return out

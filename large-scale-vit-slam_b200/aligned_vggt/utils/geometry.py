"""Drop-in for the reference's aligned_vggt/utils/geometry.py: every public name of that module.  `unproject_depth_map_to_point_map`
(:39-75, the per-pixel step right after the path) runs in liblsvs_b200.so on CUDA tensors only; the
pose-sized helpers (`averagePoseEncodings` :4-37, `project_world_points_to_pixels` :77-105, `compute_relative_poses` :107-140,
`generate_3D_pixel_grid` :142-158) are differentiable torch glue — `training/loss.py:8` back-propagates through
`compute_relative_poses`."""
import ctypes

import torch

from lsvs_b200 import native as _n
from lsvs_b200 import posemath as _pm


def averagePoseEncodings(pose_encodings: torch.Tensor) -> torch.Tensor:
    """reference :4-37.  (B,N,7) -> (B,1,7): mean translation and the Markley quaternion mean (principal eigenvector of
    sum q q^T / N over the normalised quaternions).  Inside the chunk chain this runs in csrc/pose.cu (4x4 Jacobi)."""
    q = pose_encodings[..., 3:7]
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    M = (q.unsqueeze(-1) * q.unsqueeze(-2)).mean(dim=1)
    qm = torch.linalg.eigh(M)[1][..., -1]
    qm = qm / qm.norm(dim=-1, keepdim=True)
    return torch.cat([pose_encodings[..., :3].mean(dim=1, keepdim=True), qm.unsqueeze(1)], dim=-1).float()


def project_world_points_to_pixels(world_points: torch.Tensor, extrinsics_cam: torch.Tensor, intrinsics_cam: torch.Tensor):
    """reference :77-105.  world_points (B,S,H,W,3), extrinsics (B,S,3,4), intrinsics (B,S,3,3) -> homogeneous pixels
    (B,S,H,W,3) = (u, v, w) divided by |w| where 1e-8 < |w| < 100 (w keeps its sign), and that validity mask (B,S,H,W)."""
    B, S, H, W, _ = world_points.shape
    with torch.amp.autocast("cuda", enabled=False):
        pts = world_points.reshape(B, S, H * W, 3)
        cam = pts @ extrinsics_cam[..., :3, :3].transpose(-1, -2) + extrinsics_cam[..., None, :3, 3]
        pix = cam @ intrinsics_cam.transpose(-1, -2)
        w = pix[..., 2].abs()
        valid = (w > 1e-8) & (w < 100.0)
        pix = torch.where(valid[..., None], pix / w.clamp(min=1e-30)[..., None], pix)
    return pix.view(B, S, H, W, 3), valid.view(B, S, H, W)


def compute_relative_poses(extrinsics: torch.Tensor, offset: int = 1, toNext: bool = True) -> torch.Tensor:
    """reference :107-140.  (B,S,3,4) world-to-camera -> (B,S-offset,3,4): pose of frame s+offset relative to frame s
    (toNext) or the other way round.  Differentiable (training/loss.py:8)."""
    S = extrinsics.shape[1]
    if S <= offset:
        raise Exception("To small sequence for offset")
    with torch.amp.autocast("cuda", enabled=False):
        w2c = _pm.to_homogeneous(extrinsics.float())
        c2w = torch.linalg.inv(w2c)   # general inverse, as the reference (:127)
        rel = w2c[:, offset:] @ c2w[:, :-offset] if toNext else w2c[:, :-offset] @ c2w[:, offset:]
    return rel[:, :, :3, :4]


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

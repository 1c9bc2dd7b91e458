"""Small differentiable pose helpers for the host-side glue around the path (evaluation aligners, loss-side geometry):
plain torch on a handful of (B,S,.) poses — device agnostic, autograd friendly, never on the per-pixel hot path (that is csrc/).

Conventions are UPSTREAM vggt's: quaternions scalar-last (x,y,z,w), `mat_to_quat` returns w >= 0,
pose encoding "absT_quaR_FoV" = [t(3), quat(4), fov_h, fov_w], extrinsics world-to-camera.
"""
import torch


def quat_to_mat(q: torch.Tensor) -> torch.Tensor:
    """(...,4) xyzw (any norm) -> (...,3,3)."""
    x, y, z, w = q.unbind(-1)
    k = 2.0 / (q * q).sum(-1)
    rows = [1 - k * (y * y + z * z), k * (x * y - z * w), k * (x * z + y * w),
            k * (x * y + z * w), 1 - k * (x * x + z * z), k * (y * z - x * w),
            k * (x * z - y * w), k * (y * z + x * w), 1 - k * (x * x + y * y)]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def mat_to_quat(R: torch.Tensor) -> torch.Tensor:
    """(...,3,3) -> (...,4) xyzw with w >= 0: the numerically best of the four trace-based candidates per matrix."""
    lead = R.shape[:-2]
    m = R.reshape(lead + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    t = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    mag = torch.sqrt(t.clamp(min=0.0))                                    # 2|w|, 2|x|, 2|y|, 2|z|
    cand = torch.stack([torch.stack([t[..., 0], m21 - m12, m02 - m20, m10 - m01], -1),
                        torch.stack([m21 - m12, t[..., 1], m10 + m01, m02 + m20], -1),
                        torch.stack([m02 - m20, m10 + m01, t[..., 2], m12 + m21], -1),
                        torch.stack([m10 - m01, m20 + m02, m21 + m12, t[..., 3]], -1)], -2)   # rows: (w,x,y,z) * 2*component
    cand = cand / (2.0 * mag.clamp(min=0.1)[..., None])
    best = mag.argmax(-1)
    wxyz = torch.gather(cand, -2, best[..., None, None].expand(lead + (1, 4))).squeeze(-2)
    xyzw = torch.cat([wxyz[..., 1:], wxyz[..., :1]], -1)
    return torch.where(xyzw[..., 3:4] < 0, -xyzw, xyzw)


def to_homogeneous(extr: torch.Tensor) -> torch.Tensor:
    """(...,3,4) -> (...,4,4) (a (...,4,4) input is returned as is)."""
    if extr.shape[-2] == 4:
        return extr
    last = torch.zeros(extr.shape[:-2] + (1, 4), dtype=extr.dtype, device=extr.device)
    last[..., 0, 3] = 1.0
    return torch.cat([extr, last], -2)


def inverse_se3(m: torch.Tensor) -> torch.Tensor:
    """[R t; 0 1]^-1 = [R^T, -R^T t; 0 1] over any leading dims; (...,3,4) or (...,4,4) -> (...,4,4)."""
    Rt = m[..., :3, :3].transpose(-1, -2)
    t = -(Rt @ m[..., :3, 3:])
    return to_homogeneous(torch.cat([Rt, t], -1))


def pose_encoding_to_extri_intri(pose_encoding: torch.Tensor, image_size_hw):
    """(B,S,9) -> extrinsics (B,S,3,4), intrinsics (B,S,3,3) (principal point at the image centre)."""
    H, W = image_size_hw
    t, q = pose_encoding[..., :3], pose_encoding[..., 3:7]
    extr = torch.cat([quat_to_mat(q), t[..., None]], -1)
    fy = (H / 2.0) / torch.tan(pose_encoding[..., 7] / 2.0)
    fx = (W / 2.0) / torch.tan(pose_encoding[..., 8] / 2.0)
    K = torch.zeros(pose_encoding.shape[:-1] + (3, 3), dtype=pose_encoding.dtype, device=pose_encoding.device)
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = fx, fy, W / 2.0, H / 2.0, 1.0
    return extr, K


def extri_intri_to_pose_encoding(extrinsics: torch.Tensor, intrinsics: torch.Tensor, image_size_hw) -> torch.Tensor:
    H, W = image_size_hw
    fov_h = 2 * torch.atan((H / 2.0) / intrinsics[..., 1, 1])
    fov_w = 2 * torch.atan((W / 2.0) / intrinsics[..., 0, 0])
    return torch.cat([extrinsics[..., :3, 3], mat_to_quat(extrinsics[..., :3, :3]), fov_h[..., None], fov_w[..., None]], -1).float()

from oracle.functional import closed_form_inverse_se3  # noqa: F401


def unproject_depth_map_to_point_map(depth_map, extrinsics_cam, intrinsics_cam):
    """UPSTREAM utils/geometry.py (numpy, per frame): depth (S,H,W[,1]), extrinsics (S,3,4) world-to-camera, intrinsics (S,3,3)
    -> world points (S,H,W,3).  Only imported by the reference's visualisation (aligned_vggt/utils/visualization.py:17)."""
    import numpy as np
    import torch
    to_np = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    d, E, K = to_np(depth_map), to_np(extrinsics_cam), to_np(intrinsics_cam)
    if d.ndim == 4:
        d = d[..., 0]
    S, H, W = d.shape
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    out = np.empty((S, H, W, 3), dtype=np.float32)
    for s in range(S):
        cam = np.stack([(u - K[s, 0, 2]) * d[s] / K[s, 0, 0], (v - K[s, 1, 2]) * d[s] / K[s, 1, 1], d[s]], axis=-1)
        R, t = E[s, :3, :3], E[s, :3, 3]
        out[s] = (cam - t) @ R      # R^T (x_cam - t), row-vector form
    return out

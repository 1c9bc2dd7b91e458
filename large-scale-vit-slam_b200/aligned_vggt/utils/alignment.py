"""Drop-in for the hot-path part of the reference's aligned_vggt/utils/alignment.py (:244-323, :428-594).

Same function names, argument meaning and error behaviour; the arithmetic runs in liblsvs_b200.so
(csrc/sim3.cu) on the tensors' CUDA device.  No CPU fallback: CPU tensors raise.
"""
import ctypes

import torch

from lsvs_b200 import native as _n


def _prep(x: torch.Tensor, name: str) -> torch.Tensor:
    if not x.is_cuda:
        raise _n.NativeError(f"{name} must be a CUDA tensor (no CPU fallback on this path)")
    return x.detach().to(torch.float32).contiguous()


def apply_sim3_alignment_on_point_maps(point_maps: torch.Tensor, alignment_transforms: torch.Tensor,
                                       alignment_scales: torch.Tensor) -> torch.Tensor:
    """reference alignment.py:491-526.  (B,S,H,W,3)|(S,H,W,3), (B,4,4)|(4,4), (B,)|() -> (B,S,H,W,3)."""
    if point_maps.dim() == 4:
        point_maps = point_maps.unsqueeze(0)
        alignment_transforms = alignment_transforms.unsqueeze(0)
        alignment_scales = alignment_scales.unsqueeze(0)
    assert point_maps.shape[0] == alignment_transforms.shape[0] == alignment_scales.shape[0], \
        "Inputs must have matching batch dimension"
    B, S, H, W, _ = point_maps.shape
    pts = _prep(point_maps, "point_maps")
    T = _prep(alignment_transforms, "alignment_transforms")
    s = _prep(alignment_scales, "alignment_scales").reshape(B)
    out = torch.empty_like(pts)
    _n.check(_n.lib().lsvs_sim3_apply_points(_n.ptr(pts), _n.ptr(T), _n.ptr(s), _n.ptr(out), ctypes.c_int(B),
                                             ctypes.c_longlong(S * H * W), _n.stream_ptr()), "sim3_apply_points")
    return out


def apply_sim3_alignment_on_c2w(poses: torch.Tensor, alignment_transform: torch.Tensor,
                                alignment_scales: torch.Tensor) -> torch.Tensor:
    """reference alignment.py:558-594.  (B,S,4,4) -> (B,S,4,4)."""
    if poses.dim() == 3:
        poses = poses.unsqueeze(0)
        alignment_transform = alignment_transform.unsqueeze(0)
        alignment_scales = alignment_scales.unsqueeze(0)
    assert poses.shape[0] == alignment_transform.shape[0] == alignment_scales.shape[0], \
        "Inputs must have matching batch dimension"
    B, S = poses.shape[:2]
    p = _prep(poses, "poses")
    out = torch.empty((B, S, 4, 4), dtype=torch.float32, device=p.device)
    _n.check(_n.lib().lsvs_sim3_apply_c2w(_n.ptr(p), _n.ptr(_prep(alignment_transform, "alignment_transform")),
                                          _n.ptr(_prep(alignment_scales, "alignment_scales").reshape(B)), _n.ptr(out),
                                          ctypes.c_int(B), ctypes.c_int(S), _n.stream_ptr()), "sim3_apply_c2w")
    return out


def apply_sim3_alignment_on_w2c(extr: torch.Tensor, alignment_transform: torch.Tensor,
                                alignment_scales: torch.Tensor) -> torch.Tensor:
    """reference alignment.py:528-556.  (B,S,3,4)|(B,S,4,4) -> (B,S,4,4)."""
    if extr.dim() == 3:
        extr = extr.unsqueeze(0)
        alignment_transform = alignment_transform.unsqueeze(0)
        alignment_scales = alignment_scales.unsqueeze(0)
    assert extr.shape[0] == alignment_transform.shape[0] == alignment_scales.shape[0], \
        "Inputs must have matching batch dimension"
    B, S, rows = extr.shape[:3]
    e = _prep(extr, "extr")
    out = torch.empty((B, S, 4, 4), dtype=torch.float32, device=e.device)
    _n.check(_n.lib().lsvs_sim3_apply_w2c(_n.ptr(e), ctypes.c_int(rows), _n.ptr(_prep(alignment_transform, "alignment_transform")),
                                          _n.ptr(_prep(alignment_scales, "alignment_scales").reshape(B)), _n.ptr(out),
                                          ctypes.c_int(B), ctypes.c_int(S), _n.stream_ptr()), "sim3_apply_w2c")
    return out


def scale_depth(depth: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """`depth *= chunk_scale.view(B,1,1,1,1)` (featureAligned_vggt.py:171), out of place."""
    B = depth.shape[0]
    d = _prep(depth, "depth")
    out = torch.empty_like(d)
    _n.check(_n.lib().lsvs_scale_rows(_n.ptr(d), _n.ptr(_prep(scales, "scales").reshape(B)), _n.ptr(out), ctypes.c_int(B),
                                      ctypes.c_longlong(d.numel() // B), _n.stream_ptr()), "scale_rows")
    return out


def _as_device_f32(x, device) -> torch.Tensor:
    """numpy array / python scalars / tensor -> fp32 tensor on `device` (the reference does torch.from_numpy(...).float().to(device))."""
    return torch.as_tensor(x).to(device=device, dtype=torch.float32)


def apply_sim3_alignment(alignment_transforms, alignment_scales, pose_encodings: torch.Tensor, images_size: tuple,
                         points: torch.Tensor = None, depths: torch.Tensor = None) -> tuple:
    """reference alignment.py:449-489.  alignment_transforms (B,4,4) and alignment_scales (B,) numpy arrays (tensors accepted);
    pose_encodings (B,S,9) -> aligned (w2c -> Sim(3) in camera-to-world space -> encoding, one kernel); points (B,S,H,W,3)
    transformed; depths (B,S,H,W,1) scaled IN PLACE like the reference's `depths *= ...`."""
    dev = pose_encodings.device
    T = _as_device_f32(alignment_transforms, dev)
    s = _as_device_f32(alignment_scales, dev)
    B = T.shape[0]
    from lsvs_b200.engine import pose_enc_apply_sim3
    pose_encodings = pose_enc_apply_sim3(pose_encodings, T, s.reshape(B), images_size)
    if points is not None:
        points = apply_sim3_alignment_on_point_maps(points, T, s.reshape(B))
    if depths is not None:
        if depths.dtype == torch.float32 and depths.is_contiguous() and depths.is_cuda:
            _n.check(_n.lib().lsvs_scale_rows(_n.ptr(depths), _n.ptr(s.reshape(B).contiguous()), _n.ptr(depths), ctypes.c_int(B),
                                              ctypes.c_longlong(depths.numel() // B), _n.stream_ptr()), "scale_rows")
        else:
            depths.copy_(scale_depth(depths, s))
    return pose_encodings, points, depths


def apply_sim3_alignment_on_dict(pred: dict, images_size: tuple, alignment_poses, alignment_scales) -> None:
    """reference alignment.py:428-447: Sim(3) applied to pred["pose_enc"], and to "world_points" / "depth" when present."""
    pose, pts, dep = apply_sim3_alignment(alignment_poses, alignment_scales, pred["pose_enc"], images_size,
                                          pred["world_points"] if "world_points" in pred else None,
                                          pred["depth"] if "depth" in pred else None)
    pred["pose_enc"] = pose
    if "world_points" in pred:
        pred["world_points"] = pts
    if "depth" in pred:
        pred["depth"] = dep


def scale_align_from_depths(predictions: dict, batch: dict) -> None:
    """reference alignment.py:244-323 — one robust L1-optimal scale per batch element from predicted vs ground-truth depth
    (confidence- and inverse-depth-weighted median of the ratios), applied in place to depth, world points and the pose
    translations; `predictions["alignment_scales"]` = list of python floats (the reference's `.item()` per batch)."""
    d_pred, conf = predictions["depth"], predictions["depth_conf"]
    d_gt, mask = batch["depths"], batch["point_masks"]
    B, S, H, W, _ = d_pred.shape
    N = S * H * W
    x, y = _prep(d_pred, "depth").reshape(B, N), _prep(d_gt, "depths").reshape(B, N)
    m, c = _prep(mask, "point_masks").reshape(B, N), _prep(conf, "depth_conf").reshape(B, N)
    lib = _n.lib()
    lib.lsvs_depth_scale_align_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(lib.lsvs_depth_scale_align_workspace_bytes(ctypes.c_int(B)), dtype=torch.uint8, device=x.device)
    scales = torch.empty(B, dtype=torch.float32, device=x.device)
    _n.check(lib.lsvs_depth_scale_align(_n.ptr(x), _n.ptr(y), _n.ptr(m), _n.ptr(c), ctypes.c_int(B), ctypes.c_longlong(N), _n.ptr(scales),
                                        _n.ptr(ws), _n.stream_ptr()), "depth_scale_align")
    predictions["depth"] *= scales[:, None, None, None, None]
    if "world_points" in predictions:
        predictions["world_points"] *= scales[:, None, None, None, None]
    if "pose_enc" in predictions:
        predictions["pose_enc"][..., :3] *= scales[:, None, None]
    predictions["alignment_scales"] = [scales[b].item() for b in range(B)]

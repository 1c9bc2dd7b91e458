"""Drop-in for the reference's aligned_vggt/utils/alignment.py: every public name of that module.

Hot-path part (:244-323 weighted-median depth scale, :428-594 Sim(3) application): same function names, argument meaning and
error behaviour; the arithmetic runs in liblsvs_b200.so (csrc/sim3.cu, csrc/evalgeom.cu) on the tensors' CUDA device.  No CPU
fallback there: CPU tensors raise.
Evaluation-side solvers on a few hundred camera positions / a masked point subset (:6-129 `umeyama`, `methodOfHorn`,
`scale_lse_solver`; :131-242, :325-426 the ground-truth aligners that `alignAndConvertOutputs` dispatches to): host glue in
numpy / torch like the reference's, their results applied through the kernels above.
"""
import ctypes

import numpy as np
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


# ------------------------------------------------------------------------------------------------ evaluation-side solvers
def _kabsch(cov: np.ndarray):
    """SVD of a 3x3 cross-covariance -> (proper rotation U S V^T, singular values, S diagonal)."""
    u, d, vt = np.linalg.svd(cov)
    sign = np.ones(cov.shape[0])
    if np.linalg.det(u) * np.linalg.det(vt) < 0.0:
        sign[-1] = -1.0
    return (u * sign) @ vt, d, sign


def umeyama(x: np.ndarray, y: np.ndarray) -> tuple:
    """reference :6-59.  Least-squares Sim(m) with y ~ c r x + t (Umeyama 1991).  x, y (m,n) -> r (m,m), t (m,), c."""
    assert x.shape == y.shape, "x shape not equal to y shape"
    n = x.shape[1]
    mx, my = x.mean(axis=1), y.mean(axis=1)
    xc, yc = x - mx[:, None], y - my[:, None]
    r, d, sign = _kabsch(yc @ xc.T / n)
    c = float((d * sign).sum() / ((xc ** 2).sum() / n))
    return r, my - c * (r @ mx), c


def methodOfHorn(model: np.ndarray, data: np.ndarray, align_scale: bool = True) -> tuple:
    """reference :61-111 (closed-form trajectory alignment, evaluate_ate_scale).  model, data (3,n) -> rot (3,3), trans (3,), s."""
    assert model.shape == data.shape, "model shape not equal to data shape"
    model, data = np.asarray(model, dtype=np.float64), np.asarray(data, dtype=np.float64)
    mc, dc = model - model.mean(1, keepdims=True), data - data.mean(1, keepdims=True)
    rot, _, _ = _kabsch(dc @ mc.T)
    s = float(((rot @ mc) * dc).sum() / (mc ** 2).sum()) if align_scale else 1.0
    trans = data.mean(1) - s * (rot @ model.mean(1))
    return rot, trans, np.asarray(s)


def scale_lse_solver(x: np.ndarray, y: np.ndarray) -> float:
    """reference :113-129.  argmin_s |s x - y|^2, made positive.  x, y (n,3)."""
    assert x.shape == y.shape, "x shape not equal to y shape"
    return np.abs(np.sum(x * y) / np.sum(x ** 2))


def _positions(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def _scale_maps_(predictions: dict, index, factor) -> None:
    for key in ("depth", "world_points"):
        if key in predictions:
            predictions[key][index] *= factor


def per_frame_scale_alignment_from_poses(predictions: dict, batch: dict) -> None:
    """reference :131-165: one least-squares scale per frame from the camera translations (frame 0 keeps 1.0), applied in place to
    pose translations, depth and world points.  `alignment_scales` holds the last batch element's list, as in the reference."""
    B, S = batch["extrinsics"].shape[:2]
    gt, pr = _positions(batch["extrinsics"][..., :3, 3]), _positions(predictions["pose_enc"][..., :3])
    for b in range(B):
        frame_scales = [1.0 if s == 0 else scale_lse_solver(pr[b, s], gt[b, s]) for s in range(S)]
        for s, f in enumerate(frame_scales):
            predictions["pose_enc"][b, s, :3] *= f
            _scale_maps_(predictions, (b, s), f)
    predictions["alignment_scales"] = frame_scales


def per_chunk_scale_alignment_from_poses(predictions: dict, batch: dict) -> None:
    """reference :167-203: chunked (list-valued) predictions / batch; one scale per (chunk, batch element)."""
    out = []
    for c in range(len(batch["extrinsics"])):
        gt, pr = _positions(batch["extrinsics"][c][..., :3, 3]), _positions(predictions["pose_enc"][c][..., :3])
        scales = [scale_lse_solver(pr[b], gt[b]) for b in range(gt.shape[0])]
        for b, f in enumerate(scales):
            predictions["pose_enc"][c][b, :, :3] *= f
            for key in ("depth", "world_points"):
                if key in predictions:
                    predictions[key][c][b, ...] *= f
        out.append(torch.tensor(scales))
    predictions["alignment_scales_per_chunk"] = out


def scale_alignment_from_poses(predictions: dict, batch: dict, seq_width: int = -1) -> None:
    """reference :206-242: one scale per batch element from the first `seq_width` frames' translations (-1 = all)."""
    B = batch["extrinsics"].shape[0]
    if seq_width == -1:
        seq_width = batch["extrinsics"].shape[1]
    gt, pr = _positions(batch["extrinsics"][:, :seq_width, :3, 3]), _positions(predictions["pose_enc"][:, :seq_width, :3])
    scales = [scale_lse_solver(pr[b], gt[b]) for b in range(B)]
    for b, f in enumerate(scales):
        predictions["pose_enc"][b, :, :3] *= f
        _scale_maps_(predictions, (b, Ellipsis), f)
    predictions["alignment_scales"] = scales


def _sim3_matrices(solutions):
    T = np.tile(np.eye(4), (len(solutions), 1, 1))
    for b, (r, t, _) in enumerate(solutions):
        T[b, :3, :3], T[b, :3, 3] = r, t
    return T, np.array([c for _, _, c in solutions])


def umeyama_alignment_from_poses(predictions: dict, batch: dict, seq_width: int) -> None:
    """reference :325-370: Sim(3) between predicted and ground-truth camera centres of the first `seq_width` frames (Umeyama),
    applied to pose encodings, world points and depth through apply_sim3_alignment."""
    from lsvs_b200 import posemath as pm
    B = batch["extrinsics"].shape[0]
    hw = tuple(batch["images"].shape[-2:])
    gt_c = _positions(pm.inverse_se3(batch["extrinsics"][:, :seq_width].float())[..., :3, 3])
    pred_extr, _ = pm.pose_encoding_to_extri_intri(predictions["pose_enc"][:, :seq_width].float(), hw)
    pr_c = _positions(pm.inverse_se3(pred_extr)[..., :3, 3])
    T, c = _sim3_matrices([umeyama(pr_c[b].T, gt_c[b].T) for b in range(B)])
    pose, pts, dep = apply_sim3_alignment(T, c, predictions["pose_enc"], hw, predictions.get("world_points"), predictions.get("depth"))
    predictions["pose_enc"] = pose
    if "world_points" in predictions:
        predictions["world_points"] = pts
    if "depth" in predictions:
        predictions["depth"] = dep


def umeyama_alignment_from_points(pred_points, pred_confidence, target_points, target_point_mask, confidence_threshold: int) -> tuple:
    """reference :372-426: per batch element, Umeyama between the predicted and target points that are valid in the target mask and
    at least at the `confidence_threshold` percentile of predicted confidence.  Returns (poses (B,4,4), scales (B,)) numpy arrays.
    (Like the reference, the selected (n,3) arrays are reshaped — not transposed — to (3,n) before the solve.)"""
    as_np = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    pred_points, pred_confidence, target_points, target_point_mask = map(as_np, (pred_points, pred_confidence, target_points, target_point_mask))
    sols = []
    for b in range(pred_points.shape[0]):
        conf = pred_confidence[b]
        keep = (target_point_mask[b] > 0) & (conf >= np.percentile(conf, confidence_threshold)) & (conf > 1e-5)
        sols.append(umeyama(pred_points[b][keep].reshape(3, -1), target_points[b][keep].reshape(3, -1)))
    return _sim3_matrices(sols)

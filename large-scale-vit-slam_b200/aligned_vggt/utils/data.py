"""Drop-in for the reference's aligned_vggt/utils/data.py (host logic around the path): every public name of that module —
`extri_to_pose_encoding` / `pose_encoding_to_extri` (:12-52), `convertDictListsToTensors` (:54-87), `moveDictListItemToCPU`
(:88-105), `alignAndConvertOutputs` (:107-153), `generate_chunks` (:155-207), `chunk_batch` (:209-225), `check_valid_tensor`
(:228-238), `normalize_camera_extrinsics_and_points_batch` (:241-334) — plus, like the reference (`from ...alignment import *`,
:8), everything aligned_vggt.utils.alignment exports."""
import logging
from typing import Optional, Tuple

import torch

from aligned_vggt.utils.alignment import *  # noqa: F401,F403  (reference data.py:8 re-exports the alignment module the same way)
from aligned_vggt.utils import alignment as _al
from lsvs_b200 import posemath as _pm
from lsvs_b200.scheduler import generate_chunks  # noqa: F401  (same signature and ValueError as the reference)


def extri_to_pose_encoding(extrinsics: torch.Tensor) -> torch.Tensor:
    """reference :12-30.  (B,S,3|4,4) -> (B,S,7) = [t, unit quaternion xyzw]."""
    q = _pm.mat_to_quat(extrinsics[:, :, :3, :3])
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    return torch.cat([extrinsics[:, :, :3, 3], q], dim=-1).float()


def pose_encoding_to_extri(pose_encoding: torch.Tensor) -> torch.Tensor:
    """reference :33-52.  (B,S,>=7) -> homogeneous (B,S,4,4); only [..., :3] and [..., 3:7] are read, the quaternion is normalised."""
    q = pose_encoding[..., 3:7]
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    return _pm.to_homogeneous(torch.cat([_pm.quat_to_mat(q), pose_encoding[..., :3, None]], dim=-1))

_KEYS_TO_MERGE = ["pose_enc", "pose_enc_list", "world_points", "world_points_conf", "depth", "depth_conf", "extrinsics", "intrinsics",
                  "scales", "cam_points", "depths", "point_masks", "images", "ids"]


def convertDictListsToTensors(chunked_dict: dict, overlap: int, out_dict: dict = None) -> None:
    """Concatenate the per-chunk lists along dim 1, dropping the first `overlap` frames of every chunk but the first; results go
    to `out_dict` (or back into `chunked_dict`).  As in the reference the per-chunk lists are trimmed in place."""
    if out_dict is None:
        out_dict = chunked_dict
    for key in list(chunked_dict.keys()):
        if key not in _KEYS_TO_MERGE:
            continue
        items = chunked_dict[key]
        if isinstance(items[0], list):  # list of pose-encoding lists
            if overlap > 0:
                for i in range(1, len(items)):
                    items[i] = [t[:, overlap:] for t in items[i]]
            out_dict[key] = [torch.cat(ts, dim=1) for ts in zip(*items)]
        else:
            if overlap > 0:
                for i in range(1, len(items)):
                    items[i] = items[i][:, overlap:]
            out_dict[key] = torch.cat(items, dim=1)


def moveDictListItemToCPU(chunked_dict: dict, itemIndex: int) -> None:
    """Move one chunk's entries of every list to the CPU (reference :88-105; called after each chunk, training_metrics.py:650)."""
    for key in chunked_dict.keys():
        v = chunked_dict[key]
        if isinstance(v, list) and len(v) >= (abs(itemIndex) if itemIndex < 0 else itemIndex + 1):
            if isinstance(v[0], list):
                v[itemIndex] = [(t.cpu() if isinstance(t, torch.Tensor) else t) for t in v[itemIndex]]
            elif isinstance(v[itemIndex], torch.Tensor):
                v[itemIndex] = v[itemIndex].cpu()


def alignAndConvertOutputs(predictions: dict, batch: dict, chunked_batch: dict, alignment_type: str, seq_width: int, overlap: int) -> None:
    """reference :107-153: merge the per-chunk lists (predictions in place, chunked_batch into batch) and run the requested
    ground-truth alignment.  "per_chunk_scale_from_poses" works on the chunk lists before merging; unknown types align nothing."""
    if alignment_type == "per_chunk_scale_from_poses":
        _al.per_chunk_scale_alignment_from_poses(predictions, chunked_batch)
    convertDictListsToTensors(chunked_batch, overlap, batch)
    convertDictListsToTensors(predictions, overlap)
    if alignment_type == "scale_from_fc_poses":
        _al.scale_alignment_from_poses(predictions, batch, seq_width)
    elif alignment_type == "scale_from_poses":
        _al.scale_alignment_from_poses(predictions, batch)
    elif alignment_type == "per_frame_scale_from_poses":
        _al.per_frame_scale_alignment_from_poses(predictions, batch)
    elif alignment_type == "scale_from_depths":
        if "depth" not in predictions:
            raise ValueError("scale_from_depths alignment requires depth head to be enabled.")
        _al.scale_align_from_depths(predictions, batch)
    elif alignment_type == "sim3_from_poses":
        _al.umeyama_alignment_from_poses(predictions, batch, seq_width)
    elif alignment_type == "sim3_from_points":
        if "world_points" not in predictions:
            raise ValueError("sim3_from_points alignment requires point head to be enabled.")
        T, c = _al.umeyama_alignment_from_points(predictions["world_points"][:, :seq_width], predictions["world_points_conf"][:, :seq_width],
                                                batch["world_points"][:, :seq_width], batch["point_masks"][:, :seq_width],
                                                confidence_threshold=50.0)
        _al.apply_sim3_alignment_on_dict(predictions, batch["images"].shape[-2:], T, c)


def chunk_batch(batch: dict, indices: list) -> dict:
    """reference :209-225: every tensor entry (B,N,...) of `batch` -> list of per-chunk tensors batch[key][:, chunk_ids]."""
    tensors = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
    return {k: [v[:, ids] for ids in indices] for k, v in tensors.items()} if indices else {}


def check_valid_tensor(input_tensor: Optional[torch.Tensor], name: str = "tensor") -> None:
    """reference :228-238: log a warning if the tensor holds NaN / Inf."""
    if input_tensor is not None and not bool(torch.isfinite(input_tensor).all()):
        logging.warning(f"NaN or Inf found in tensor: {name}")


def _zero_non_finite(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    """UPSTREAM check_and_fix_inf_nan(..., hard_max=None): NaN / Inf entries become 0 (with a warning)."""
    if t is None:
        return None
    bad = ~torch.isfinite(t)
    if bool(bad.any()):
        logging.warning(f"Inf or NaN found in {name}; replaced with 0")
        t = torch.where(bad, torch.zeros_like(t), t)
    return t


def normalize_camera_extrinsics_and_points_batch(
        extrinsics: torch.Tensor, cam_points: Optional[torch.Tensor] = None, world_points: Optional[torch.Tensor] = None,
        depths: Optional[torch.Tensor] = None, scale_by_points: bool = True, point_masks: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """reference :241-334 (data-loader side, CPU tensors as there): re-base extrinsics (B,S,3,4) and world points on the first camera
    and, with `scale_by_points`, divide translations / points / depths by the mean distance of the valid points."""
    for t, n in ((extrinsics, "extrinsics"), (cam_points, "cam_points"), (world_points, "world_points"), (depths, "depths")):
        check_valid_tensor(t, n)
    assert extrinsics.device == torch.device("cpu")
    B = extrinsics.shape[0]
    first = extrinsics[:, 0]
    new_extr = _pm.to_homogeneous(extrinsics) @ _pm.inverse_se3(first).unsqueeze(1)                 # (B,S,4,4)
    new_world = None
    if world_points is not None:  # x_cam0 = R0 x_world + t0
        new_world = world_points @ first[:, None, None, :3, :3].transpose(-1, -2) + first[:, None, None, None, :3, 3]
    if not scale_by_points:
        return new_extr[:, :, :3], cam_points, new_world, depths
    dist = new_world.norm(dim=-1)
    avg = ((dist * point_masks).sum(dim=[1, 2, 3]) / (point_masks.sum(dim=[1, 2, 3]) + 1e-3)).clamp(min=1e-6, max=1e6)
    new_world = new_world / avg.view(B, 1, 1, 1, 1)
    new_extr = new_extr[:, :, :3].clone()
    new_extr[:, :, :3, 3] = new_extr[:, :, :3, 3] / avg.view(B, 1, 1)
    new_depths = depths.clone() / avg.view(B, 1, 1, 1)
    new_cam = cam_points.clone() / avg.view(B, 1, 1, 1, 1)
    return (_zero_non_finite(new_extr, "new_extrinsics"), _zero_non_finite(new_cam, "new_cam_points"),
            _zero_non_finite(new_world, "new_world_points"), _zero_non_finite(new_depths, "new_depths"))

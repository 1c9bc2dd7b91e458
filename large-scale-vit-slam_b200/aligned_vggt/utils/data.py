"""Drop-in for the chunk bookkeeping of the reference's aligned_vggt/utils/data.py (host logic around the path):
`generate_chunks` (:155-207), `convertDictListsToTensors` (:54-87), `moveDictListItemToCPU` (:88-105)."""
import torch

from lsvs_b200.scheduler import generate_chunks  # noqa: F401  (same signature and ValueError as the reference)

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

"""Drop-in for the reference's aligned_vggt/models/poseAligned_wrapped_vggt.py: the pose-aligned baseline `VGGT`
(:16-204).  Chunk-to-chunk SE(3) = Markley mean of the relative overlap poses (:107-126), applied to the camera poses
(:129) and, inverted, to the point maps (:171-187).  Same kernels as the feature-aligned path: the pose chain
(csrc/pose.cu) degenerates to this when the learned chunk Sim(3) / per-frame SE(3) are the identity.

DPT point / depth heads are outside this path: `raw_points`, `raw_points_conf`, `raw_depth`,
`raw_depth_conf` stand in for their outputs.  The gt_poses scale alignment (:84-104, chunk_gt sampling) is a
training/evaluation-time path and raises NotImplementedError."""
import torch
import torch.nn as nn

from aligned_vggt.utils.alignment import apply_sim3_alignment_on_point_maps
from lsvs_b200.engine import Engine, pose_chain
from lsvs_b200.modules import Aggregator, CameraHead

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass


class VGGT(nn.Module, PyTorchModelHubMixin):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, enable_camera=True, enable_point=True, enable_depth=True,
                 enable_track=True, depth=24, patch_embed_depth=24, intermediate_layer_indices=(4, 11, 17, 23)):
        super().__init__()
        self.intermediate_layer_indices = list(intermediate_layer_indices)
        self.aggregator = Aggregator(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, depth=depth,
                                     patch_embed_depth=patch_embed_depth, keep_layers=self.intermediate_layer_indices)
        self.camera_head = CameraHead(dim_in=2 * embed_dim) if enable_camera else None
        self.point_head = self.depth_head = self.track_head = None  # DPT / track heads: SURVEY §8f, not on this path yet
        for child in (self.aggregator, self.camera_head):
            if child is not None:
                child._bind(self)

    def set_config(self, cfg):
        self.camera_head = self.camera_head if cfg.enable_camera else None

    def _engine(self) -> Engine:
        eng = self.__dict__.get("_native_engine")
        if eng is None:
            eng = Engine(self.aggregator.depth, self.aggregator.dino_depth, 0, 8, False, self.camera_head is not None, self.aggregator.rope_freq)
            self.__dict__["_native_engine"] = eng
        eng.sync(self.named_parameters())
        return eng

    def forward(self, images: torch.Tensor, num_overlap: int, context: dict = None, gt_poses: torch.Tensor = None,
                raw_points=None, raw_points_conf=None, raw_depth=None, raw_depth_conf=None) -> dict:
        """reference :36-204."""
        if gt_poses is not None:
            raise NotImplementedError("gt_poses (sample_mode chunk_gt) is a training-time path outside this build")
        B, S, C, H, W = images.shape
        predictions = {}
        tokens_list, _ = self.aggregator(images)
        taps = [tokens_list[i] for i in self.intermediate_layer_indices]
        del tokens_list
        point_T = None
        if self.camera_head is not None:
            cam_enc = self.camera_head(taps)[-1]
            dev = cam_enc.device
            ident_sim3 = torch.tensor([0, 0, 0, 0, 0, 0, 1, 1], dtype=torch.float32, device=dev).view(1, 1, 8).expand(B, -1, -1).contiguous()
            ident_se3 = torch.tensor([0, 0, 0, 0, 0, 0, 1], dtype=torch.float32, device=dev).view(1, 1, 7).expand(B, S - 1, -1).contiguous()
            prev = context["pose_enc"][-1] if context is not None else None
            aligned_pose_enc, point_T, _ = pose_chain(ident_sim3, ident_se3, cam_enc, prev, num_overlap, (H, W))  # :107-130
            _append(predictions, context, "pose_enc", aligned_pose_enc)
        if raw_depth is not None:  # :139-157 (no scale without gt_poses)
            _append(predictions, context, "depth", raw_depth)
            _append(predictions, context, "depth_conf", raw_depth_conf)
        if raw_points is not None:  # :159-195
            pts = raw_points
            if point_T is not None:
                pts = apply_sim3_alignment_on_point_maps(raw_points, point_T, torch.ones(B, device=raw_points.device))
            _append(predictions, context, "world_points", pts)
            _append(predictions, context, "world_points_conf", raw_points_conf)
        if not self.training:
            _append(predictions, context, "images", images)
        return predictions


def _append(predictions, context, key, value):
    if context is None:
        predictions[key] = [value]
    else:
        context.setdefault(key, []).append(value)
        predictions[key] = context[key]

"""Drop-in for the reference's aligned_vggt/models/poseAligned_wrapped_vggt.py: the pose-aligned baseline `VGGT`
(:16-204).  Chunk-to-chunk SE(3) = Markley mean of the relative overlap poses (:107-126), applied to the camera poses
(:129) and, inverted, to the point maps (:171-187).  Same kernels as the feature-aligned path: the pose chain
(csrc/pose.cu) degenerates to this when the learned chunk Sim(3) / per-frame SE(3) are the identity.

The DPT point / depth heads (:22-23, :132-137, :156-158) run on the engine like the feature-aligned model's;
`raw_points`, `raw_points_conf`, `raw_depth`, `raw_depth_conf` (optional) stand in for their outputs.  With `gt_poses`
(B,S,3,4) (sample modes chunk_gt / two_chunks) the chain takes its chunk transform from gt_poses[:,0] and scales
translations, depth and points by the least-squares position scale (:84-109, :144-147, :166-169) inside the same kernel
instead of the reference's numpy round trip."""
import torch
import torch.nn as nn

from aligned_vggt.utils.alignment import apply_sim3_alignment_on_point_maps
from aligned_vggt.utils.alignment import scale_depth
from lsvs_b200.engine import GT_MEAN, GT_SCALE, Engine, pose_chain
from lsvs_b200.modules import Aggregator, CameraHead, DPTHead

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
        self.point_head = DPTHead(dim_in=2 * embed_dim, output_dim=4, activation="inv_log", conf_activation="expp1",
                                  prefix="point_head.") if enable_point else None  # :22
        self.depth_head = DPTHead(dim_in=2 * embed_dim, output_dim=2, activation="exp", conf_activation="expp1",
                                  prefix="depth_head.") if enable_depth else None  # :23
        self.track_head = None  # never called by the reference's forward; not built
        self._bind_children()

    def _bind_children(self):
        for child in (self.aggregator, self.camera_head, self.point_head, self.depth_head):
            if child is not None:
                child._bind(self)
        self.__dict__.pop("_native_engine", None)

    def set_config(self, cfg):
        """reference :27-34."""
        self.camera_head = self.camera_head if cfg.enable_camera else None
        self.point_head = self.point_head if cfg.enable_point else None
        self.depth_head = self.depth_head if cfg.enable_depth else None
        self._bind_children()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        from lsvs_b200.modules import load_state_dict_without_track_head
        return load_state_dict_without_track_head(self, state_dict, strict, assign)

    def __getstate__(self):
        # copy.deepcopy / pickle (EMA copies, checkpoint cloning): the native engine handle is process-local and is rebuilt lazily
        state = self.__dict__.copy()
        state.pop("_native_engine", None)
        return state

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
        B, S, C, H, W = images.shape
        predictions = {}
        tokens_list, patch_start_idx = self.aggregator(images)
        taps = [tokens_list[i] for i in self.intermediate_layer_indices]
        del tokens_list
        point_T = batch_scales = None
        if self.camera_head is not None:
            cam_enc = self.camera_head(taps)[-1]
            dev = cam_enc.device
            ident_sim3 = torch.tensor([0, 0, 0, 0, 0, 0, 1, 1], dtype=torch.float32, device=dev).view(1, 1, 8).expand(B, -1, -1).contiguous()
            ident_se3 = torch.tensor([0, 0, 0, 0, 0, 0, 1], dtype=torch.float32, device=dev).view(1, 1, 7).expand(B, S - 1, -1).contiguous()
            prev = context["pose_enc"][-1] if context is not None else None
            gt_mode = 0
            if gt_poses is not None:  # :84-109
                if tuple(gt_poses.shape[-2:]) != (3, 4):  # the reference pads one row (:86) and would build 5x4 matrices otherwise
                    raise RuntimeError(f"gt_poses must be (B,S,3,4) world-to-camera matrices, got {tuple(gt_poses.shape)}")
                gt_mode = GT_MEAN | GT_SCALE
            aligned_pose_enc, point_T, scale = pose_chain(ident_sim3, ident_se3, cam_enc, prev, num_overlap, (H, W),
                                                          gt_poses=gt_poses, gt_mode=gt_mode)  # :107-130
            if gt_poses is not None and S > 1:
                batch_scales = scale
            _append(predictions, context, "pose_enc", aligned_pose_enc)
        if raw_depth is None and self.depth_head is not None:  # :132-135
            raw_depth, raw_depth_conf = self.depth_head(taps, images=images, patch_start_idx=patch_start_idx)
        if raw_points is None and self.point_head is not None:  # :156-158
            raw_points, raw_points_conf = self.point_head(taps, images=images, patch_start_idx=patch_start_idx)
        if raw_depth is not None:  # :137-153 (gt scale only)
            depth = scale_depth(raw_depth, batch_scales) if batch_scales is not None else raw_depth
            _append(predictions, context, "depth", depth)
            _append(predictions, context, "depth_conf", raw_depth_conf)
        if raw_points is not None:  # :159-195
            pts = raw_points
            if point_T is not None:
                s = batch_scales if batch_scales is not None else torch.ones(B, device=raw_points.device)
                pts = apply_sim3_alignment_on_point_maps(raw_points, point_T, s)
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

"""Drop-in for the reference's aligned_vggt/models/featureAligned_vggt.py (FeatureAlignedVGGT :16, forward :48,
merge_results :227): same constructor, forward signature, returned keys and state_dict names.  The Aggregator,
alignment head, camera head, pose composition and Sim(3) application run as sm_100a kernels behind the C ABI.

The DPT depth / point heads (reference :28-29, :166-207) run on the same engine (bf16 tensor-core convolutions).
Not on this path: the track head (`enable_track` is accepted and ignored, `track_head` stays None —
the reference's forward never calls it either).  `raw_depth` / `raw_points` (optional) stand in for the DPT outputs,
e.g. to exercise the Sim(3) application (:171, :187-207) on chosen inputs."""
import torch
import torch.nn as nn

from aligned_vggt.heads.alignment_head import AlignmentHead
from aligned_vggt.utils import alignment as _al
from lsvs_b200 import train as _train
from lsvs_b200.engine import GT_MEAN, Engine, pose_chain
from lsvs_b200.modules import Aggregator, CameraHead, DPTHead

try:  # the reference mixes in the HF hub loader; keep it when the package is present
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass


class FeatureAlignedVGGT(nn.Module, PyTorchModelHubMixin):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, enable_camera=True, enable_point=True,
                 enable_depth=True, enable_track=True, num_memory_tokens=8, temporal_attention=True,
                 depth=24, patch_embed_depth=24, intermediate_layer_indices=(4, 11, 17, 23), precision=None):
        """precision (not a reference argument): None / 0 = bf16 tensor-core operands like the reference's bf16-mixed inference,
        1 = fp32-class alignment head + camera-head trunk, 2 = fp32-class everywhere (include/lsvs_b200.h)."""
        super().__init__()
        self.precision = precision
        self.embed_dim = embed_dim
        self.enable_memory = num_memory_tokens > 0
        self.intermediate_layer_indices = list(intermediate_layer_indices)
        self.aggregator = Aggregator(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, depth=depth,
                                     patch_embed_depth=patch_embed_depth, keep_layers=self.intermediate_layer_indices)
        self.camera_head = CameraHead(dim_in=2 * embed_dim) if enable_camera else None
        self.point_head = DPTHead(dim_in=2 * embed_dim, output_dim=4, activation="inv_log", conf_activation="expp1",
                                  prefix="point_head.") if enable_point else None  # :28
        self.depth_head = DPTHead(dim_in=2 * embed_dim, output_dim=2, activation="exp", conf_activation="expp1",
                                  prefix="depth_head.") if enable_depth else None  # :29
        self.track_head = None  # not on this path (never called by the reference's forward)
        self._requested_heads = dict(point=enable_point, depth=enable_depth, track=enable_track)
        self.alignment_head = AlignmentHead(in_dim=2 * embed_dim, patch_size=patch_size, num_memory_tokens=num_memory_tokens,
                                            temporal_attention=temporal_attention)
        self._bind_children()

    def _bind_children(self):
        for child in (self.aggregator, self.camera_head, self.alignment_head, self.point_head, self.depth_head):
            if child is not None:
                child._bind(self)
        self.__dict__.pop("_native_engine", None)

    def set_config(self, cfg):
        """reference :34-46 — called after from_pretrained; re-creates the alignment head (random init)."""
        self.camera_head = self.camera_head if cfg.enable_camera else None
        self.point_head = self.point_head if cfg.enable_point else None
        self.depth_head = self.depth_head if cfg.enable_depth else None
        self._requested_heads = dict(point=cfg.enable_point, depth=cfg.enable_depth, track=cfg.enable_track)
        self.enable_memory = cfg.num_memory_tokens > 0
        dev = next(self.aggregator.parameters()).device
        self.alignment_head = AlignmentHead(in_dim=2 * self.embed_dim, patch_size=cfg.patch_size,
                                            num_memory_tokens=cfg.num_memory_tokens,
                                            temporal_attention=cfg.temporal_attention).to(dev)
        self._bind_children()

    def set_precision(self, precision):
        """Switch the arithmetic mode (0 / 1 / 2, see __init__); the native engine is rebuilt and re-packs the weights on the next forward."""
        self.precision = precision
        self.__dict__.pop("_native_engine", None)

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
            eng = Engine(self.aggregator.depth, self.aggregator.dino_depth, self.alignment_head.depth_aa,
                         self.alignment_head.num_memory_tokens, True, self.camera_head is not None, self.aggregator.rope_freq,
                         precision=self.precision)
            self.__dict__["_native_engine"] = eng
        eng.sync(self.named_parameters())
        return eng

    def forward(self, images: torch.Tensor, num_overlap: int, context: dict = None, gt_poses: torch.Tensor = None,
                raw_depth=None, raw_points=None) -> dict:
        """images (B,S,3,H,W) in [0,1] -> predictions dict with the reference's keys (:60-71)."""
        B, S, C, H, W = images.shape
        predictions = {}
        tokens_list, patch_start_idx = self.aggregator(images)
        taps = [tokens_list[i] for i in self.intermediate_layer_indices]
        del tokens_list

        ctx_overlap = ctx_memory = None
        if context is not None:
            ctx_overlap = context["overlap_tokens"]
            if self.enable_memory:
                ctx_memory = context["memory_tokens"][-1]
        overlap = num_overlap if S > num_overlap else S - 1  # :93
        train_path = _train.wants_training_path(self.alignment_head)   # train() + grad enabled + trainable head: autograd path
        chunk_sim3_enc, frame_se3_enc, memory_tokens, overlap_tokens = self.alignment_head(
            taps[-1], (H, W), overlap, overlap_tokens=ctx_overlap, memory_tokens=ctx_memory)

        point_T = chunk_scale = None
        if self.camera_head is not None:
            cam_enc = self.camera_head(taps)[-1]
            prev = context["pose_enc"][-1] if context is not None else None
            gt, gt_mode = None, 0
            if gt_poses is not None and context is not None:  # :123-124 mean_camera_transform = gt_poses[:, :1]
                if tuple(gt_poses.shape[-2:]) != (4, 4):     # the reference's matmul with the (B,S,4,4) chain fails on (3,4) here
                    raise RuntimeError(f"gt_poses must be (B,S,4,4) homogeneous world-to-camera matrices, got {tuple(gt_poses.shape)}")
                gt, gt_mode = gt_poses, GT_MEAN
            if train_path:  # differentiable twin of the kernel: the training losses on pose_enc reach the head through it
                aligned_pose_enc, point_T, chunk_scale = _train.pose_chain_train(chunk_sim3_enc, frame_se3_enc, cam_enc, prev, overlap, (H, W),
                                                                                 gt_mean=None if gt is None else gt[:, :1])
            else:
                aligned_pose_enc, point_T, chunk_scale = pose_chain(chunk_sim3_enc, frame_se3_enc, cam_enc, prev, overlap, (H, W),
                                                                    gt_poses=gt, gt_mode=gt_mode)
            predictions["overlap_tokens"] = overlap_tokens
            if context is None:
                predictions["pose_enc"] = [aligned_pose_enc]
                predictions["chunk_sim3_alignment_enc"] = chunk_sim3_enc
                predictions["frame_se3_alignment_enc"] = frame_se3_enc
                if self.enable_memory:
                    predictions["memory_tokens"] = [memory_tokens]
            else:
                context.setdefault("pose_enc", []).append(aligned_pose_enc)
                predictions["pose_enc"] = context["pose_enc"]
                predictions["chunk_sim3_alignment_enc"] = merge_results(context["chunk_sim3_alignment_enc"], chunk_sim3_enc, 0, 1)
                predictions["frame_se3_alignment_enc"] = merge_results(context["frame_se3_alignment_enc"], frame_se3_enc, 0, 1)
                if self.enable_memory:
                    context.setdefault("memory_tokens", []).append(memory_tokens)
                    predictions["memory_tokens"] = context["memory_tokens"]
        if chunk_scale is None:
            chunk_scale = chunk_sim3_enc[..., -1].reshape(B)

        depth_conf = pts_conf = None
        if raw_depth is None and self.depth_head is not None:  # :166-168
            raw_depth, depth_conf = self.depth_head(taps, images=images, patch_start_idx=patch_start_idx)
        if raw_points is None and self.point_head is not None:  # :183-185
            raw_points, pts_conf = self.point_head(taps, images=images, patch_start_idx=patch_start_idx)
        if raw_depth is not None:  # :171 depth *= chunk_scale
            depth = raw_depth * chunk_scale.view(B, 1, 1, 1, 1) if train_path else _al.scale_depth(raw_depth, chunk_scale)
            _append(predictions, context, "depth", depth)
            if depth_conf is not None:
                _append(predictions, context, "depth_conf", depth_conf)
        if raw_points is not None:  # :187-207 scale, then the chunk's point transform
            if point_T is None:
                pts = raw_points
            elif train_path:
                pts = _train.apply_sim3_points_train(raw_points, point_T, chunk_scale)
            else:
                pts = _al.apply_sim3_alignment_on_point_maps(raw_points, point_T, chunk_scale)
            _append(predictions, context, "world_points", pts)
            if pts_conf is not None:
                _append(predictions, context, "world_points_conf", pts_conf)
        if not self.training:
            _append(predictions, context, "images", images)
        return predictions


def _append(predictions, context, key, value):
    if context is None:
        predictions[key] = [value]
    else:
        context.setdefault(key, []).append(value)
        predictions[key] = context[key]


def merge_results(first_chunk, second_chunk, num_overlap: int = 0, mergeDim: int = 1):
    """reference :227-254."""
    if isinstance(first_chunk, list) and isinstance(second_chunk, list):
        if num_overlap > 0:
            second_chunk = [item[:, num_overlap:] for item in second_chunk]
        return [torch.cat((a, b), dim=mergeDim) for a, b in zip(first_chunk, second_chunk)]
    if num_overlap > 0:
        second_chunk = second_chunk[:, num_overlap:]
    return torch.cat((first_chunk, second_chunk), dim=mergeDim)

"""Drop-in for the reference's aligned_vggt/models/pointAligned_wrapped_vggt.py: the point-aligned baseline `VGGT`
(:14-157) and its IRLS weighted-Umeyama solver (:159-305).  Aggregator / camera head run in the native engine, the
Sim(3) estimation in csrc/umeyama.cu (no host round trips inside the IRLS loop), its application in csrc/sim3.cu.

The DPT point / depth heads (:19-20, :69-72, :130-133) run on the engine like the feature-aligned model's; `raw_points`,
`raw_points_conf`, `raw_depth`, `raw_depth_conf` (optional) stand in for their outputs.  `gt_poses` is accepted and
unused, as in the reference."""
import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from aligned_vggt.utils.alignment import apply_sim3_alignment_on_point_maps, scale_depth
from lsvs_b200 import native as _n
from lsvs_b200.engine import Engine, pose_enc_apply_sim3
from lsvs_b200.modules import Aggregator, CameraHead, DPTHead

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass


def _f(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise _n.NativeError("irls_sim3_umeyama needs CUDA tensors (no CPU fallback on this path)")
    return x.detach().to(torch.float32).contiguous()


def irls_sim3_umeyama(src: torch.Tensor, dst: torch.Tensor, conf_src: Optional[torch.Tensor], conf_dst: Optional[torch.Tensor],
                      conf_threshold_factor: float = 0.5, delta: float = 0.1, max_iters: int = 20, tol: float = 1e-9
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """reference :219-305.  src, dst (N,H,W,3); conf (N,H,W) -> R (3,3), t (3,), s ()."""
    assert src.shape[0] == dst.shape[0]
    x, y = _f(src).reshape(-1, 3), _f(dst).reshape(-1, 3)
    cs, cd = _f(conf_src).reshape(-1), _f(conf_dst).reshape(-1)
    assert x.shape == y.shape and cs.numel() == x.shape[0] == cd.numel()
    dev = x.device
    R, t, s = torch.empty(3, 3, device=dev), torch.empty(3, device=dev), torch.empty((), device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _n.lib()
    lib.lsvs_irls_umeyama_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(int(lib.lsvs_irls_umeyama_workspace_bytes()), dtype=torch.uint8, device=dev)
    _n.check(lib.lsvs_irls_umeyama(_n.ptr(x), _n.ptr(y), _n.ptr(cs), _n.ptr(cd), ctypes.c_longlong(x.shape[0]),
                                   ctypes.c_float(conf_threshold_factor), ctypes.c_float(delta), ctypes.c_int(max_iters),
                                   ctypes.c_float(tol), _n.ptr(R), _n.ptr(t), _n.ptr(s), _n.ptr(status), _n.ptr(ws), _n.stream_ptr()),
             "irls_umeyama")
    if int(status.item()) != 0:  # the reference raises here too (:184-185); this is the only host sync of the solver
        raise ValueError("Total weight too small for meaningful estimation")
    return R, t, s


def weighted_umeyama_sim3(src: torch.Tensor, dst: torch.Tensor, weights: torch.Tensor):
    """reference :159-217.  One weighted solve: src, dst (M,3), weights (M,)."""
    assert src.ndim == 2 and src.shape[1] == 3
    assert dst.shape == src.shape
    w = _f(weights)
    return irls_sim3_umeyama(src.reshape(1, -1, 1, 3), dst.reshape(1, -1, 1, 3), (w * w).reshape(1, -1, 1), torch.ones_like(w).reshape(1, -1, 1),
                             conf_threshold_factor=0.0, max_iters=0)


class VGGT(nn.Module, PyTorchModelHubMixin):
    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, enable_camera=True, enable_point=True, enable_depth=True,
                 enable_track=True, depth=24, patch_embed_depth=24, intermediate_layer_indices=(4, 11, 17, 23)):
        super().__init__()
        self.intermediate_layer_indices = list(intermediate_layer_indices)
        self.aggregator = Aggregator(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, depth=depth,
                                     patch_embed_depth=patch_embed_depth, keep_layers=self.intermediate_layer_indices)
        self.camera_head = CameraHead(dim_in=2 * embed_dim) if enable_camera else None
        self.point_head = DPTHead(dim_in=2 * embed_dim, output_dim=4, activation="inv_log", conf_activation="expp1",
                                  prefix="point_head.") if enable_point else None  # :19
        self.depth_head = DPTHead(dim_in=2 * embed_dim, output_dim=2, activation="exp", conf_activation="expp1",
                                  prefix="depth_head.") if enable_depth else None  # :20
        self.track_head = None  # never called by the reference's forward; not built
        self._bind_children()

    def _bind_children(self):
        for child in (self.aggregator, self.camera_head, self.point_head, self.depth_head):
            if child is not None:
                child._bind(self)
        self.__dict__.pop("_native_engine", None)

    def set_config(self, cfg):
        """reference :25-32."""
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
        """reference :34-157."""
        B, S, C, H, W = images.shape
        predictions = {}
        tokens_list, patch_start_idx = self.aggregator(images)
        taps = [tokens_list[i] for i in self.intermediate_layer_indices]
        del tokens_list
        alignment_transform = batch_scales = None
        if raw_points is None and self.point_head is not None:  # :69-72
            raw_points, raw_points_conf = self.point_head(taps, images=images, patch_start_idx=patch_start_idx)
        if raw_depth is None and self.depth_head is not None:  # :130-133
            raw_depth, raw_depth_conf = self.depth_head(taps, images=images, patch_start_idx=patch_start_idx)
        if raw_points is not None:
            pts3d, pts3d_conf = raw_points, raw_points_conf
            if context is not None:
                ctx_pts = context["world_points"][-1][:, -num_overlap:].to(pts3d.device)
                ctx_conf = context["world_points_conf"][-1][:, -num_overlap:].to(pts3d.device)
                Ts, ss = [], []
                for b in range(B):  # :82-92
                    r, t, s = irls_sim3_umeyama(pts3d[b, :num_overlap], ctx_pts[b], pts3d_conf[b, :num_overlap], ctx_conf[b])
                    pose = torch.eye(4, device=pts3d.device, dtype=torch.float32)
                    pose[:3, :3] = r
                    pose[:3, 3] = t
                    Ts.append(pose)
                    ss.append(s)
                alignment_transform, batch_scales = torch.stack(Ts), torch.stack(ss)
            else:
                alignment_transform = torch.eye(4, device=images.device, dtype=torch.float32).view(1, 4, 4).expand(B, -1, -1).contiguous()
                batch_scales = torch.ones(B, device=pts3d.device, dtype=torch.float32)
            pts3d_final = apply_sim3_alignment_on_point_maps(pts3d, alignment_transform, batch_scales)
            _append(predictions, context, "world_points", pts3d_final)
            _append(predictions, context, "world_points_conf", pts3d_conf)
        if self.camera_head is not None:
            pose_enc = self.camera_head(taps)[-1]
            if alignment_transform is not None:  # :113-122
                pose_enc = pose_enc_apply_sim3(pose_enc, alignment_transform, batch_scales, (H, W))
            _append(predictions, context, "pose_enc", pose_enc)
        if raw_depth is not None:  # :135-138
            depth = scale_depth(raw_depth, batch_scales) if batch_scales is not None else raw_depth
            _append(predictions, context, "depth", depth)
            _append(predictions, context, "depth_conf", raw_depth_conf)
        if not self.training:
            _append(predictions, context, "images", images)
        return predictions


def _append(predictions, context, key, value):
    if context is None:
        predictions[key] = [value]
    else:
        context.setdefault(key, []).append(value)
        predictions[key] = context[key]

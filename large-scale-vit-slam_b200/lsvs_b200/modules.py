"""B200-native stand-ins for the UPSTREAM vggt modules on the path (Aggregator, CameraHead): parameter containers
with the upstream state_dict names whose forward runs in the native engine."""
from typing import List, Optional

import torch
import torch.nn as nn

from . import specs
from .engine import Engine


def load_state_dict_without_track_head(self, state_dict, strict: bool = True, assign: bool = False):
    """nn.Module.load_state_dict for the model mirrors: `track_head.*` entries of a facebook/VGGT-1B style checkpoint are skipped
    while the model has no track head (the reference's forward never calls it), so strict loading keeps working."""
    if getattr(self, "track_head", None) is None:
        state_dict = {k: v for k, v in state_dict.items() if not k.startswith("track_head.")}
    return nn.Module.load_state_dict(self, state_dict, strict=strict, assign=assign)


class _EngineBound(specs.ParamTree):
    """A parameter tree bound to a (possibly shared) native engine under a name prefix."""
    _prefix = ""

    def _bind(self, owner):
        object.__setattr__(self, "_owner", owner)  # not a submodule: avoid a reference cycle in the module tree

    def __getstate__(self):
        # copy.deepcopy / pickle (EMA copies, checkpoint cloning): the native engine handle is process-local and is rebuilt lazily
        state = self.__dict__.copy()
        state.pop("_own_engine", None)
        return state

    def _engine(self) -> Engine:
        owner = getattr(self, "_owner", None)
        if owner is not None:
            return owner._engine()
        eng = getattr(self, "_own_engine", None)
        if eng is None:
            eng = self._make_engine()
            object.__setattr__(self, "_own_engine", eng)
        eng.sync((self._prefix + n, p) for n, p in self.named_parameters())
        return eng


class Aggregator(_EngineBound):
    """UPSTREAM vggt.models.aggregator.Aggregator.  forward(images) -> (list of depth
    tensors (B,S,P,2048) — only the layers in `keep_layers` are materialised, the rest are None —, patch_start_idx)."""
    _prefix = "aggregator."

    def __init__(self, img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                 num_register_tokens=4, qkv_bias=True, proj_bias=True, ffn_bias=True, patch_embed="dinov2_vitl14_reg",
                 aa_order=("frame", "global"), aa_block_size=1, qk_norm=True, rope_freq=100, init_values=0.01,
                 patch_embed_depth=24, keep_layers=None):
        if not (embed_dim == 1024 and num_heads == 16 and patch_size == 14 and num_register_tokens == 4 and mlp_ratio == 4.0
                and patch_embed == "dinov2_vitl14_reg" and list(aa_order) == ["frame", "global"] and aa_block_size == 1
                and qk_norm and rope_freq > 0):
            raise ValueError("only the VGGT-1B Aggregator geometry is built on this path")
        super().__init__(specs.aggregator_spec(depth, patch_embed_depth, img_size, patch_size, embed_dim))
        self.depth, self.dino_depth, self.rope_freq = depth, patch_embed_depth, float(rope_freq)
        self.patch_size, self.patch_start_idx = patch_size, 1 + num_register_tokens
        self.keep_layers = None if keep_layers is None else [int(i) for i in keep_layers]
        self.patch_embed.mask_token.requires_grad_(False)
        specs.init_default_(self)

    def _make_engine(self):
        return Engine(self.depth, self.dino_depth, 0, 8, False, False, self.rope_freq)

    def forward(self, images: torch.Tensor):
        keep = list(range(self.depth)) if self.keep_layers is None else self.keep_layers
        taps = self._engine().aggregator_forward(images, keep)
        out: List[Optional[torch.Tensor]] = [None] * self.depth
        for i, t in zip(keep, taps):
            out[i] = t
        return out, self.patch_start_idx


class CameraHead(_EngineBound):
    """UPSTREAM vggt.heads.camera_head.CameraHead (A.5): forward(list of tapped layers, num_iterations=4) -> list of
    `num_iterations` activated pose encodings (B,S,9), one per refinement iteration (the reference reads [-1],
    featureAligned_vggt.py:109)."""
    _prefix = "camera_head."

    def __init__(self, dim_in=2048, trunk_depth=4, pose_encoding_type="absT_quaR_FoV", num_heads=16, mlp_ratio=4,
                 init_values=0.01, trans_act="linear", quat_act="linear", fl_act="relu"):
        if not (dim_in == 2048 and trunk_depth == 4 and num_heads == 16 and pose_encoding_type == "absT_quaR_FoV"
                and (trans_act, quat_act, fl_act) == ("linear", "linear", "relu")):
            raise ValueError("only the VGGT-1B CameraHead geometry is built on this path")
        super().__init__(specs.camera_head_spec(dim_in, trunk_depth))
        specs.init_default_(self)

    def _make_engine(self):
        return Engine(0, 0, 0, 8, False, True)

    def forward(self, aggregated_tokens_list, num_iterations: int = 4):
        return self._engine().camera_head_forward(aggregated_tokens_list[-1], num_iterations, all_iterations=True)


class DPTHead(_EngineBound):
    """UPSTREAM vggt.heads.dpt_head.DPTHead (A.6): forward(aggregated_tokens_list, images, patch_start_idx) ->
    (pred (B,S,H,W,output_dim-1), conf (B,S,H,W)).  Construction as in featureAligned_vggt.py:28-29; the state_dict keys
    are facebook/VGGT-1B's `depth_head.*` / `point_head.*` (torch conv layouts; re-laid out when pushed to the engine)."""

    def __init__(self, dim_in=2048, patch_size=14, output_dim=4, activation="inv_log", conf_activation="expp1", features=256,
                 out_channels=(256, 512, 1024, 1024), intermediate_layer_idx=(4, 11, 17, 23), pos_embed=True,
                 feature_only=False, down_ratio=1, prefix="point_head."):
        if not (dim_in == 2048 and patch_size == 14 and features == 256 and tuple(out_channels) == (256, 512, 1024, 1024)
                and pos_embed and not feature_only and down_ratio == 1 and 2 <= output_dim <= 4):
            raise ValueError("only the VGGT-1B depth / point DPTHead geometry is built on this path")
        if activation not in ("exp", "inv_log"):
            raise ValueError(f"Unknown activation: {activation}")
        if conf_activation != "expp1":
            raise ValueError(f"Unknown conf_activation: {conf_activation}")
        super().__init__(specs.dpt_head_spec(dim_in, output_dim, features, out_channels))
        self.output_dim, self.activation = output_dim, activation
        self.intermediate_layer_idx = list(intermediate_layer_idx)
        self._prefix = prefix
        specs.init_default_(self)

    def _make_engine(self):
        return Engine(0, 0, 0, 8, False, False)

    def forward(self, aggregated_tokens_list, images, patch_start_idx: int = 5, frames_chunk_size: int = None):
        """frames_chunk_size: frames per pass (upstream default 8; None = chosen by the engine from a memory budget — the
        computation is per frame, so the result does not depend on it)."""
        if patch_start_idx != 5:
            raise ValueError("DPTHead: patch_start_idx must be 5 (camera token + 4 register tokens)")
        taps = [aggregated_tokens_list[i] for i in self.intermediate_layer_idx] if len(aggregated_tokens_list) > 4 \
            else list(aggregated_tokens_list)
        H, W = images.shape[-2:]
        return self._engine().dpt_head_forward(self._prefix, taps, (H, W), self.output_dim, self.activation, frames_chunk_size)

"""DPTHead as a parameter container with the upstream module tree / state_dict names
(facebook/VGGT-1B `depth_head.*`, `point_head.*`); math in oracle.functional.dpt_head_forward.
TEST INFRASTRUCTURE — recalled from the public upstream repository (not in the container)."""
import torch.nn as nn
from oracle import functional as OF
from .._p import params_of


class _ResidualConvUnit(nn.Module):
    def __init__(self, features):
        super().__init__()
        self.conv1 = nn.Conv2d(features, features, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(features, features, 3, 1, 1, bias=True)


class _FeatureFusionBlock(nn.Module):
    def __init__(self, features, has_residual=True):
        super().__init__()
        self.out_conv = nn.Conv2d(features, features, 1, 1, 0, bias=True)
        if has_residual:
            self.resConfUnit1 = _ResidualConvUnit(features)
        self.resConfUnit2 = _ResidualConvUnit(features)


class DPTHead(nn.Module):
    def __init__(self, dim_in, patch_size=14, output_dim=4, activation="inv_log", conf_activation="expp1", features=256,
                 out_channels=(256, 512, 1024, 1024), intermediate_layer_idx=(4, 11, 17, 23), pos_embed=True,
                 feature_only=False, down_ratio=1):
        super().__init__()
        assert pos_embed and not feature_only and down_ratio == 1 and features == OF.DPT_FEATURES
        assert tuple(out_channels) == OF.DPT_OUT_CHANNELS
        self.patch_size, self.activation, self.conf_activation = patch_size, activation, conf_activation
        self.intermediate_layer_idx = list(intermediate_layer_idx)
        oc = list(out_channels)
        self.norm = nn.LayerNorm(dim_in)
        self.projects = nn.ModuleList([nn.Conv2d(dim_in, c, 1, 1, 0) for c in oc])
        self.resize_layers = nn.ModuleList([nn.ConvTranspose2d(oc[0], oc[0], 4, 4, 0), nn.ConvTranspose2d(oc[1], oc[1], 2, 2, 0),
                                            nn.Identity(), nn.Conv2d(oc[3], oc[3], 3, 2, 1)])
        scratch = nn.Module()
        for i, c in enumerate(oc):
            setattr(scratch, f"layer{i + 1}_rn", nn.Conv2d(c, features, 3, 1, 1, bias=False))
        scratch.refinenet1 = _FeatureFusionBlock(features)
        scratch.refinenet2 = _FeatureFusionBlock(features)
        scratch.refinenet3 = _FeatureFusionBlock(features)
        scratch.refinenet4 = _FeatureFusionBlock(features, has_residual=False)
        scratch.output_conv1 = nn.Conv2d(features, features // 2, 3, 1, 1)
        scratch.output_conv2 = nn.Sequential(nn.Conv2d(features // 2, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv2d(32, output_dim, 1, 1, 0))
        self.scratch = scratch

    def forward(self, aggregated_tokens_list, images, patch_start_idx, frames_chunk_size=8):
        taps = [aggregated_tokens_list[i] for i in self.intermediate_layer_idx]
        return OF.dpt_head_forward(params_of(self), "", taps, tuple(images.shape[-2:]), patch_start_idx, self.activation,
                                   self.conf_activation, self.patch_size)

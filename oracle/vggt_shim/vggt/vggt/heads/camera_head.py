"""CameraHead (A.5) as a parameter container; math in oracle.functional.camera_head_forward."""
import torch
import torch.nn as nn
from oracle import functional as OF
from .._p import params_of
from ..layers import Mlp
from ..layers.block import Block


class CameraHead(nn.Module):
    def __init__(self, dim_in=2048, trunk_depth=4, pose_encoding_type="absT_quaR_FoV", num_heads=16, mlp_ratio=4,
                 init_values=0.01, trans_act="linear", quat_act="linear", fl_act="relu"):
        super().__init__()
        assert pose_encoding_type == "absT_quaR_FoV" and (trans_act, quat_act, fl_act) == ("linear", "linear", "relu")
        self.target_dim = 9
        self.trunk_depth = trunk_depth
        self.num_heads = num_heads
        self.trunk = nn.Sequential(*[Block(dim=dim_in, num_heads=num_heads, mlp_ratio=mlp_ratio, init_values=init_values)
                                     for _ in range(trunk_depth)])
        self.token_norm = nn.LayerNorm(dim_in)
        self.trunk_norm = nn.LayerNorm(dim_in)
        self.empty_pose_tokens = nn.Parameter(torch.zeros(1, 1, self.target_dim))
        self.embed_pose = nn.Linear(self.target_dim, dim_in)
        self.poseLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim_in, 3 * dim_in, bias=True))
        self.adaln_norm = nn.LayerNorm(dim_in, elementwise_affine=False, eps=1e-6)
        self.pose_branch = Mlp(in_features=dim_in, hidden_features=dim_in // 2, out_features=self.target_dim, drop=0)

    def forward(self, aggregated_tokens_list, num_iterations=4):
        return OF.camera_head_forward(params_of(self), "", aggregated_tokens_list[-1], num_iterations,
                                      self.trunk_depth, self.num_heads)

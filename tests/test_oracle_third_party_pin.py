"""Mechanical pins of the UPSTREAM restatements in oracle/functional.py against an INDEPENDENT implementation of the same published
architectures: HuggingFace `transformers` (installed in this image; library code, never on the product path).

UPSTREAM facebookresearch/vggt is absent from the container, so oracle/functional.py restates it from its
published description.  Two of its building blocks are published models in their own right and ship in `transformers`:

  * DINOv2 ViT with register tokens (`Dinov2WithRegistersModel`) — VGGT's patch embedder (aggregator.patch_embed.*): patch
    convolution, [cls | registers | patches] ordering, bicubic anti-aliased position-embedding interpolation, pre-LN blocks with
    fused-qkv attention, exact-erf GELU MLP, LayerScale, final LayerNorm (eps 1e-6);
  * the DPT refinement stage (`DPTFeatureFusionLayer` / `DPTPreActResidualLayer`) — the fusion blocks of VGGT's DPTHead
    (ResidualConvUnit, skip add, bilinear x2 with align_corners, 1x1 out_conv).

The same random weights are loaded into both (key mapping below); outputs must agree to fp32 rounding.  What this does NOT pin:
the alternating frame / global attention with q/k norm + 2-D RoPE of the Aggregator, the camera head, and the `nn.ReLU(inplace=True)`
quirk of VGGT's fusion blocks (the `relu_inplace=False` reading is the one `transformers` implements and the one pinned here).
"""
import pytest
import torch

transformers = pytest.importorskip("transformers")

from oracle import functional as OF  # noqa: E402


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 if p.dim() == 1 else 1.0 / max(1.0, float(p[0].numel()) ** 0.5)))


def _dinov2_params(hf_sd, prefix, depth):
    """HF Dinov2WithRegisters keys -> UPSTREAM (facebookresearch/dinov2) keys as VGGT's state_dict carries them."""
    p = {prefix + "patch_embed.proj.weight": hf_sd["embeddings.patch_embeddings.projection.weight"],
         prefix + "patch_embed.proj.bias": hf_sd["embeddings.patch_embeddings.projection.bias"],
         prefix + "cls_token": hf_sd["embeddings.cls_token"], prefix + "register_tokens": hf_sd["embeddings.register_tokens"],
         prefix + "pos_embed": hf_sd["embeddings.position_embeddings"],
         prefix + "norm.weight": hf_sd["layernorm.weight"], prefix + "norm.bias": hf_sd["layernorm.bias"]}
    for i in range(depth):
        h, o = f"encoder.layer.{i}.", f"{prefix}blocks.{i}."
        for n in ("norm1", "norm2"):
            p[o + n + ".weight"], p[o + n + ".bias"] = hf_sd[h + n + ".weight"], hf_sd[h + n + ".bias"]
        a = h + "attention.attention."
        p[o + "attn.qkv.weight"] = torch.cat([hf_sd[a + "query.weight"], hf_sd[a + "key.weight"], hf_sd[a + "value.weight"]], 0)
        p[o + "attn.qkv.bias"] = torch.cat([hf_sd[a + "query.bias"], hf_sd[a + "key.bias"], hf_sd[a + "value.bias"]], 0)
        p[o + "attn.proj.weight"], p[o + "attn.proj.bias"] = hf_sd[h + "attention.output.dense.weight"], hf_sd[h + "attention.output.dense.bias"]
        p[o + "ls1.gamma"], p[o + "ls2.gamma"] = hf_sd[h + "layer_scale1.lambda1"], hf_sd[h + "layer_scale2.lambda1"]
        for n in ("fc1", "fc2"):
            p[o + "mlp." + n + ".weight"], p[o + "mlp." + n + ".bias"] = hf_sd[h + "mlp." + n + ".weight"], hf_sd[h + "mlp." + n + ".bias"]
    return p


@pytest.mark.parametrize("hidden,heads,depth,native,hw", [
    (256, 4, 3, 6, (84, 84)),        # native grid: no position-embedding interpolation
    (256, 4, 3, 6, (42, 70)),        # 3 x 5 grid from a 6 x 6 table: anti-aliased bicubic down-sampling (the 518x154 case in small)
    (1024, 16, 2, 4, (70, 98)),      # ViT-L width / heads (VGGT-1B geometry), 5 x 7 grid from a 4 x 4 table: up-sampling
])
def test_dinov2_restatement_matches_transformers(hidden, heads, depth, native, hw):
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    cfg = Dinov2WithRegistersConfig(hidden_size=hidden, num_hidden_layers=depth, num_attention_heads=heads, image_size=14 * native,
                                    patch_size=14, num_register_tokens=4, mlp_ratio=4, hidden_act="gelu", layer_norm_eps=1e-6,
                                    qkv_bias=True, layerscale_value=1.0, use_swiglu_ffn=False)
    try:
        cfg._attn_implementation = "eager"
    except Exception:
        pass
    model = Dinov2WithRegistersModel(cfg).eval()
    _randomize(model, 7 + hidden)
    images = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = model(pixel_values=images).last_hidden_state[:, 1 + 4:]          # x_norm_patchtokens
        p = _dinov2_params(model.state_dict(), "aggregator.patch_embed.", depth)
        got = OF.dinov2_patch_tokens(p, "aggregator.patch_embed.", images, depth=depth, num_heads=heads)
    assert got.shape == ref.shape
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize("with_residual", [False, True])
def test_dpt_fusion_block_restatement_matches_transformers(with_residual):
    from transformers import DPTConfig
    from transformers.models.dpt.modeling_dpt import DPTFeatureFusionLayer
    cfg = DPTConfig(fusion_hidden_size=64, use_batch_norm_in_fusion_residual=False)
    layer = DPTFeatureFusionLayer(cfg, align_corners=True).eval()
    _randomize(layer, 11)
    sd = layer.state_dict()
    pre = "scratch.refinenet1."
    p = {pre + "out_conv.weight": sd["projection.weight"], pre + "out_conv.bias": sd["projection.bias"]}
    for u in (1, 2):
        for c in (1, 2):
            p[f"{pre}resConfUnit{u}.conv{c}.weight"] = sd[f"residual_layer{u}.convolution{c}.weight"]
            p[f"{pre}resConfUnit{u}.conv{c}.bias"] = sd[f"residual_layer{u}.convolution{c}.bias"]
    g = torch.Generator().manual_seed(5)
    x0, x1 = torch.randn(2, 64, 9, 13, generator=g), torch.randn(2, 64, 9, 13, generator=g)
    with torch.no_grad():
        ref = layer(x0, x1 if with_residual else None)
        got = OF._fusion_block(p, pre, x0, x1 if with_residual else None, None, relu_inplace=False)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5


def test_rotation_and_pose_encoding_helpers_match_scipy():
    """UPSTREAM vggt.utils.rotation quat_to_mat / mat_to_quat (scalar-last x,y,z,w; canonical w >= 0), closed_form_inverse_se3 and
    the absT_quaR_FoV pose encoding against scipy.spatial.transform.Rotation (scalar-last by default) and numpy.linalg."""
    import numpy as np
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(21)
    q = torch.randn(64, 4, generator=g, dtype=torch.float64)
    ref_R = Rotation.from_quat(q.numpy()).as_matrix()                       # normalises, like the 2 / |q|^2 factor
    assert np.abs(OF.quat_to_mat(q).numpy() - ref_R).max() < 1e-12
    # matrices near every branch of matrix_to_quaternion (largest of w, x, y, z), incl. rotations by ~pi
    axes = torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [1.0, 1.0, 1.0]], dtype=torch.float64)
    rv = torch.cat([torch.randn(32, 3, generator=g, dtype=torch.float64), axes * (np.pi - 1e-3), axes * 1e-4], 0)
    R = torch.from_numpy(Rotation.from_rotvec(rv.numpy()).as_matrix())
    got = OF.mat_to_quat(R).numpy()
    ref_q = Rotation.from_matrix(R.numpy()).as_quat()
    ref_q = np.where(ref_q[:, 3:4] < 0, -ref_q, ref_q)
    assert np.abs(got - ref_q).max() < 1e-9 and (got[:, 3] >= 0).all()
    # SE(3) inverse and the pose-encoding round trip (T, quaternion, FoV from a pinhole intrinsic matrix)
    se3 = torch.eye(4, dtype=torch.float64).repeat(len(R), 1, 1)
    se3[:, :3, :3], se3[:, :3, 3] = R, torch.randn(len(R), 3, generator=g, dtype=torch.float64)
    assert np.abs(OF.closed_form_inverse_se3(se3).numpy() - np.linalg.inv(se3.numpy())).max() < 1e-12
    H, W = 154, 518
    K = torch.zeros(1, len(R), 3, 3, dtype=torch.float64)
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = 400.0, 380.0, W / 2, H / 2, 1.0
    enc = OF.extri_intri_to_pose_encoding(se3[None, :, :3].float(), K.float(), (H, W))
    assert enc.shape == (1, len(R), 9)
    assert abs(float(enc[0, 0, 7]) - 2 * np.arctan(H / 2 / 380.0)) < 1e-6 and abs(float(enc[0, 0, 8]) - 2 * np.arctan(W / 2 / 400.0)) < 1e-6
    ext, intr = OF.pose_encoding_to_extri_intri(enc, (H, W))
    assert float((ext - se3[None, :, :3].float()).abs().max()) < 2e-6 and float((intr - K.float()).abs().max()) < 2e-3


def test_rope_2d_is_the_reference_pinned_1d_rope_on_each_half():
    """UPSTREAM RotaryPositionEmbedding2D ("1-D sibling in-repo layers/rope.py has identical math"): the 2-D
    restatement must equal the 1-D one — which tests/test_oracle_golden.py pins against the reference's own rope.py output — applied
    to the first half of the head dimension with the y positions and to the second half with the x positions."""
    g = torch.Generator().manual_seed(33)
    x = torch.randn(2, 3, 5 + 11 * 37, 64, generator=g)
    pos = OF.token_positions(2, 11, 37, 5, x.device)
    got = OF.rope_apply_2d(x, pos, 100.0)
    ref = torch.cat([OF.rope_apply_1d(x[..., :32], pos[..., 0], 100.0), OF.rope_apply_1d(x[..., 32:], pos[..., 1], 100.0)], dim=-1)
    assert torch.equal(got, ref)
    assert torch.equal(got[:, :, :5], x[:, :, :5])                      # special tokens sit at (0, 0): identity rotation
    assert float((got.norm(dim=-1) - x.norm(dim=-1)).abs().max()) < 1e-4   # a rotation

"""CPU: the drop-in modules expose exactly the reference's state_dict contract (keys and shapes generated from the
reference's own classes by oracle/make_golden.py -> tests/golden/state_dict_spec.json)."""
import json
import os

import torch


def test_state_dict_matches_reference():
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_spec.json")))
    with torch.device("meta"):
        model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False)
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert set(mine) == set(ref)
    assert all(mine[k] == ref[k] for k in ref)
    names = {n for n, _ in model.named_modules()}
    assert {"aggregator", "camera_head", "alignment_head", "aggregator.frame_blocks.23.attn.qkv"} <= names  # freeze globs (*aggregator*)


def test_default_init_statistics():
    from aligned_vggt.heads.alignment_head import AlignmentHead
    h = AlignmentHead()
    sd = h.state_dict()
    m = sd["memory_token"][0]
    assert torch.allclose(m @ m.T, torch.eye(8), atol=1e-5)            # orthonormal rows (alignment_head.py:211-214)
    assert abs(float(sd["alpha"]) - 0.1) < 1e-7 and float(sd["gated_update.gate_mlp.2.bias"]) == 0.0
    assert abs(float(sd["frame_blocks.0.ls1.gamma"].mean()) - 0.01) < 1e-8
    assert float(sd["per_frame_alignment_token"].abs().max()) < 1e-4


def test_dpt_state_dict_matches_reference():
    """depth_head.* / point_head.* keys of the reference model (built on the oracle shim's recalled upstream DPTHead tree)."""
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_spec_dpt.json")))
    with torch.device("meta"):
        model = FeatureAlignedVGGT(enable_point=True, enable_depth=True, enable_track=False, depth=1, patch_embed_depth=1)
    mine = {k: list(v.shape) for k, v in model.state_dict().items() if k.startswith(("depth_head.", "point_head."))}
    assert len(ref) == 124 and set(mine) == set(ref)
    assert all(mine[k] == ref[k] for k in ref)


def test_baseline_wrappers_state_dict_matches_reference():
    """pose-aligned / point-aligned `VGGT` with their DPT heads: keys and shapes of the reference classes
    (poseAligned_wrapped_vggt.py:16-25, pointAligned_wrapped_vggt.py:14-22; enable_track=False)."""
    from aligned_vggt.models.pointAligned_wrapped_vggt import VGGT as PointVGGT
    from aligned_vggt.models.poseAligned_wrapped_vggt import VGGT as PoseVGGT
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_spec_baselines.json")))
    for cls in (PoseVGGT, PointVGGT):
        with torch.device("meta"):
            model = cls(enable_track=False, depth=1, patch_embed_depth=1)
        mine = {k: list(v.shape) for k, v in model.state_dict().items()}
        assert set(mine) == set(ref), (cls, sorted(set(mine) ^ set(ref))[:5])
        assert all(mine[k] == ref[k] for k in ref)

        class Cfg:
            enable_camera, enable_point, enable_depth, enable_track = True, False, True, False
        model.set_config(Cfg)
        assert model.point_head is None and model.depth_head is not None and model.camera_head is not None

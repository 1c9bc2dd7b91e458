"""GPU parity of the evaluation-side geometry right after the path through the C ABI, against golden
outputs of the REFERENCE functions (tests/golden/eval_geometry.npz) and against the oracle at BASELINE size.
Tolerances: unprojection 1e-5 relative (fp32, different summation order of the 3x3 inverse); the weighted-median scale must pick
the same ratio as the reference's sort + cumsum up to neighbouring order statistics: 1e-4 relative."""
import pytest
import torch

from parity_util import rel_l2

pytestmark = pytest.mark.gpu


def test_unproject_and_scale_align_golden(golden):
    from aligned_vggt.utils.alignment import scale_align_from_depths
    from aligned_vggt.utils.geometry import unproject_depth_map_to_point_map
    g = golden("eval_geometry.npz")
    t = lambda k: torch.as_tensor(g[k]).cuda()
    out = unproject_depth_map_to_point_map(t("depth"), t("extr"), t("intr"))
    assert float((out.cpu() - torch.as_tensor(g["unproj"])).abs().max()) < 1e-5 * float(abs(g["unproj"]).max())
    preds = {"depth": t("pred").clone(), "depth_conf": t("conf"), "world_points": t("pts").clone(), "pose_enc": t("pose").clone()}
    scale_align_from_depths(preds, {"depths": t("gt"), "point_masks": t("mask") > 0.5})
    sc = torch.tensor(preds["alignment_scales"])
    assert float(((sc - torch.as_tensor(g["scales"])) / torch.as_tensor(g["scales"])).abs().max()) < 1e-4, (sc, g["scales"])
    assert rel_l2(preds["depth"], torch.as_tensor(g["depth_aligned"])) < 1e-4
    assert rel_l2(preds["world_points"], torch.as_tensor(g["pts_aligned"])) < 1e-4
    assert rel_l2(preds["pose_enc"], torch.as_tensor(g["pose_aligned"])) < 1e-4


def test_scale_align_baseline_size_vs_oracle():
    """32 frames of 154x518 (2.55 M pixels per batch element), 2 batch elements, heavy-tailed ratios, 30 % masked."""
    from lsvs_b200 import native
    from oracle import aligned as OA
    import ctypes
    B, N = 2, 32 * 154 * 518
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(B, N, generator=g) * 30 + 0.3
    pred = gt / torch.tensor([[2.5], [0.4]]) * torch.exp(0.2 * torch.randn(B, N, generator=g))
    pred[:, ::97] *= 20.0
    mask = (torch.rand(B, N, generator=g) > 0.3).float()
    conf = 1 + torch.exp(torch.randn(B, N, generator=g))
    ref = OA.depth_scale_align(pred, gt, mask, conf)
    lib = native.lib()
    lib.lsvs_depth_scale_align_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(lib.lsvs_depth_scale_align_workspace_bytes(ctypes.c_int(B)), dtype=torch.uint8, device="cuda")
    out = torch.empty(B, device="cuda")
    args = [x.cuda().contiguous() for x in (pred, gt, mask, conf)]
    native.check(lib.lsvs_depth_scale_align(*[native.ptr(a) for a in args], ctypes.c_int(B), ctypes.c_longlong(N), native.ptr(out), native.ptr(ws),
                                            native.stream_ptr()), "depth_scale_align")
    assert float(((out.cpu() - ref) / ref).abs().max()) < 1e-4, (out.cpu(), ref)


def test_unproject_baseline_size_roundtrip():
    """size-independent property at BASELINE size: projecting the unprojected points back gives the pixel grid and the depth."""
    from aligned_vggt.utils.geometry import unproject_depth_map_to_point_map
    from oracle import functional as OF
    B, S, H, W = 1, 32, 154, 518
    g = torch.Generator().manual_seed(5)
    depth = torch.rand(B, S, H, W, 1, generator=g) * 20 + 0.5
    q = torch.nn.functional.normalize(torch.randn(B, S, 4, generator=g), dim=-1)
    extr = torch.cat([OF.quat_to_mat(q), torch.randn(B, S, 3, 1, generator=g)], dim=-1)
    K = torch.zeros(B, S, 3, 3)
    K[..., 0, 0] = 400.0; K[..., 1, 1] = 410.0; K[..., 0, 2] = W / 2; K[..., 1, 2] = H / 2; K[..., 2, 2] = 1.0
    pts = unproject_depth_map_to_point_map(depth.cuda(), extr.cuda(), K.cuda()).cpu().double()
    cam = pts @ extr[..., :3, :3].double().transpose(-1, -2)[:, :, None] + extr[..., :3, 3].double()[:, :, None, None]
    assert float((cam[..., 2] - depth[..., 0].double()).abs().max()) < 1e-4 * 20
    uvw = cam @ K.double().transpose(-1, -2)[:, :, None]
    u = uvw[..., 0] / uvw[..., 2]
    assert float((u - torch.arange(W).double().view(1, 1, 1, W)).abs().max()) < 1e-2


def test_convert_dict_lists_matches_reference(golden):
    from aligned_vggt.utils.data import convertDictListsToTensors, moveDictListItemToCPU
    from conftest import rnd
    g = golden("eval_geometry.npz")
    chunks = {"depth": [rnd(520 + i, 1, 4, 2, 2, 1).cuda() for i in range(3)], "pose_enc": [rnd(530 + i, 1, 4, 9).cuda() for i in range(3)], "other": [1, 2, 3]}
    moveDictListItemToCPU(chunks, 0)
    assert not chunks["depth"][0].is_cuda and chunks["depth"][1].is_cuda
    chunks["depth"][0] = chunks["depth"][0].cuda(); chunks["pose_enc"][0] = chunks["pose_enc"][0].cuda()
    convertDictListsToTensors(chunks, 1)
    assert torch.equal(chunks["depth"].cpu(), torch.as_tensor(g["merged_depth"])) and torch.equal(chunks["pose_enc"].cpu(), torch.as_tensor(g["merged_pose"]))
    assert chunks["other"] == [1, 2, 3]


@pytest.mark.parametrize("H,W", [(7, 9), (12, 20), (33, 128)])
def test_unproject_ragged_sizes_vs_oracle(H, W):
    """H*W not a multiple of 4 (pixel-per-thread kernel), partial 128-pixel chunks, and whole chunks."""
    from aligned_vggt.utils.geometry import unproject_depth_map_to_point_map
    from oracle import aligned as OA
    from oracle import functional as OF
    B, S = 2, 3
    g = torch.Generator().manual_seed(H * W)
    depth = torch.rand(B, S, H, W, 1, generator=g) * 10 + 0.1
    q = torch.nn.functional.normalize(torch.randn(B, S, 4, generator=g), dim=-1)
    extr = torch.cat([OF.quat_to_mat(q), torch.randn(B, S, 3, 1, generator=g)], dim=-1)
    K = torch.zeros(B, S, 3, 3)
    K[..., 0, 0] = 30.0; K[..., 1, 1] = 31.0; K[..., 0, 2] = W / 2; K[..., 1, 2] = H / 2; K[..., 2, 2] = 1.0; K[..., 0, 1] = 0.2
    ref = OA.unproject_depth(depth, extr, K)
    out = unproject_depth_map_to_point_map(depth.cuda(), extr.cuda(), K.cuda()).cpu()
    assert float((out - ref).abs().max()) < 1e-5 * float(ref.abs().max())

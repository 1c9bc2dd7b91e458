"""GPU parity: tcgen05 GEMM + fused epilogues through the C ABI vs the fp32 CPU oracle on bf16-rounded inputs.
Tolerance: inputs are identical bf16 values, accumulation is fp32 on both sides, so only summation order and the
final bf16 rounding of the output differ -> rel-L2 <= 4e-3 (bf16 eps = 3.9e-3) for bf16 outputs, 1e-4 for fp32."""
import pytest
import torch

from conftest import rnd
from oracle import functional as OF

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def bf(x):
    return x.to(torch.bfloat16)


def mk(M, N, K, seed=0):
    a, w = bf(rnd(seed, M, K)), bf(rnd(seed + 1, N, K, scale=0.05))
    bias = rnd(seed + 2, N, scale=0.1)
    return a, w, bias


@pytest.mark.parametrize("M,N,K", [(1, 128, 64), (127, 128, 128), (128, 256, 64), (129, 384, 192), (300, 512, 1024),
                                   (1648, 3072, 1024), (1000, 1024, 4096), (77, 640, 640)])
def test_gemm_bias_bf16(M, N, K):
    from lsvs_b200 import ops
    a, w, bias = mk(M, N, K)
    ref = a.float() @ w.float().T + bias
    out = ops.gemm(a.cuda(), w.cuda(), ops.EPI_BIAS_BF16, bias=bias.cuda()).cpu()
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert rel_l2(out, ref) < 4e-3
    assert float((out.float() - ref).abs().max()) < 0.02 * float(ref.abs().max()) + 1e-3


def test_gemm_f32_gelu_resid():
    from lsvs_b200 import ops
    M, N, K = 333, 1024, 1024
    a, w, bias = mk(M, N, K, seed=5)
    lin = a.float() @ w.float().T + bias
    out = ops.gemm(a.cuda(), w.cuda(), ops.EPI_BIAS_F32, bias=bias.cuda()).cpu()
    assert rel_l2(out, lin) < 1e-4
    out = ops.gemm(a.cuda(), w.cuda(), ops.EPI_BIAS_GELU_BF16, bias=bias.cuda()).cpu()
    assert rel_l2(out, torch.nn.functional.gelu(lin)) < 4e-3
    resid, gamma = rnd(9, M, N), 0.1 * (1 + 0.1 * rnd(10, N))
    r_dev = resid.cuda()
    tap = torch.zeros(M, 2 * N, device="cuda")
    ops.gemm(a.cuda(), w.cuda(), ops.EPI_RESID_F32, bias=bias.cuda(), gamma=gamma.cuda(), resid=r_dev, out2=tap[:, N:])
    ref = resid + gamma * lin
    assert rel_l2(r_dev.cpu(), ref) < 1e-4
    assert torch.equal(tap[:, N:].cpu(), r_dev.cpu()) and float(tap[:, :N].abs().max()) == 0.0


def test_gemm_strided_a_and_repeat():
    """A given as a column slice of a wider buffer (lda > K); two launches give identical bits (determinism)."""
    from lsvs_b200 import ops
    M, N, K = 260, 256, 128
    big = bf(rnd(3, M, 3 * K)).cuda()
    a = big[:, K:2 * K]
    w, bias = bf(rnd(4, N, K, scale=0.05)), rnd(5, N)
    o1 = ops.gemm(a, w.cuda(), ops.EPI_BIAS_BF16, bias=bias.cuda())
    o2 = ops.gemm(a, w.cuda(), ops.EPI_BIAS_BF16, bias=bias.cuda())
    assert torch.equal(o1, o2)
    assert rel_l2(o1.cpu(), a.cpu().float() @ w.float().T + bias) < 4e-3


def test_rope_table():
    from lsvs_b200 import ops
    tab = ops.rope_table(40, 16, 100.0).cpu()
    cos, sin = OF.rope_angles(32, 40, 100.0, "cpu", torch.float32)
    assert float((tab[..., 0] - cos[:, :16]).abs().max()) < 2e-6
    assert float((tab[..., 1] - sin[:, :16]).abs().max()) < 2e-6


@pytest.mark.parametrize("S,gh,gw", [(2, 3, 5), (3, 11, 37)])
def test_gemm_qkv_headnorm64_rope2d(S, gh, gw):
    """Fused qkv epilogue == upstream Attention's qkv Linear -> q_norm/k_norm -> 2-D RoPE (oracle.functional.attention)."""
    from lsvs_b200 import ops
    D, H, hd, nsp = 1024, 16, 64, 5
    P = nsp + gh * gw
    M = S * P
    x, w, bias = bf(rnd(20, M, D)), bf(rnd(21, 3 * D, D, scale=0.03)), rnd(22, 3 * D, scale=0.1)
    qn = (1 + 0.1 * rnd(23, hd), 0.1 * rnd(24, hd))
    kn = (1 + 0.1 * rnd(25, hd), 0.1 * rnd(26, hd))
    qkv = (x.float() @ w.float().T + bias).reshape(S, P, 3, H, hd).permute(2, 0, 3, 1, 4)
    pos = OF.token_positions(S, gh, gw, nsp, "cpu")
    q = OF.rope_apply_2d(torch.nn.functional.layer_norm(qkv[0], (hd,), qn[0], qn[1], 1e-5), pos)
    k = OF.rope_apply_2d(torch.nn.functional.layer_norm(qkv[1], (hd,), kn[0], kn[1], 1e-5), pos)
    ref = torch.stack([q, k, qkv[2]]).permute(1, 3, 0, 2, 4).reshape(M, 3 * D)
    tab = ops.rope_table(max(gh, gw) + 2, hd // 4)
    out = ops.gemm(x.cuda(), w.cuda(), ops.EPI_HEADNORM64_BF16, bias=bias.cuda(), qn=[t.cuda() for t in qn],
                   kn=[t.cuda() for t in kn], n_q_cols=D, n_k_cols=D, rope_mode=ops.ROPE_2D, rope_tab=tab,
                   tokens_per_frame=P, n_special=nsp, grid_w=gw).cpu()
    for sec, name in enumerate("qkv"):
        assert rel_l2(out[:, sec * D:(sec + 1) * D], ref[:, sec * D:(sec + 1) * D]) < 5e-3, name


def test_gemm_headnorm128_rope1d_and_2d():
    from lsvs_b200 import ops
    D, H, hd = 1024, 8, 128
    S, G = 4, 30  # G groups of S rows; position = ids[row % S]
    M = S * G
    ids = torch.tensor([2, 3, 4, 9], dtype=torch.int32)
    x, w, bias = bf(rnd(30, M, D)), bf(rnd(31, 2 * D, D, scale=0.03)), rnd(32, 2 * D, scale=0.1)
    kn = (1 + 0.1 * rnd(33, hd), 0.1 * rnd(34, hd))
    kv = x.float() @ w.float().T + bias
    k = kv[:, :D].reshape(G, S, H, hd).transpose(1, 2)
    k = OF.rope_apply_1d(torch.nn.functional.layer_norm(k, (hd,), kn[0], kn[1], 1e-5), ids.long().view(1, S).expand(G, -1))
    ref = torch.cat([k.transpose(1, 2).reshape(M, D), kv[:, D:]], dim=1)
    tab = ops.rope_table(16, hd // 2)
    out = ops.gemm(x.cuda(), w.cuda(), ops.EPI_HEADNORM128_BF16, bias=bias.cuda(), kn=[t.cuda() for t in kn],
                   n_q_cols=0, n_k_cols=D, rope_mode=ops.ROPE_1D, rope_tab=tab, pos_ids=ids.cuda()).cpu()
    assert rel_l2(out[:, :D], ref[:, :D]) < 5e-3 and rel_l2(out[:, D:], ref[:, D:]) < 4e-3
    # 2-D variant at head dim 128 (alignment-head frame blocks: 6 special tokens)
    gh, gw, nsp, Sf = 3, 4, 6, 3
    P = nsp + gh * gw
    M2 = Sf * P
    x2, w2, b2 = bf(rnd(40, M2, D)), bf(rnd(41, 3 * D, D, scale=0.03)), rnd(42, 3 * D, scale=0.1)
    qkv = (x2.float() @ w2.float().T + b2).reshape(Sf, P, 3, H, hd).permute(2, 0, 3, 1, 4)
    pos = OF.token_positions(Sf, gh, gw, nsp, "cpu")
    q = OF.rope_apply_2d(torch.nn.functional.layer_norm(qkv[0], (hd,), kn[0], kn[1], 1e-5), pos)
    refq = q.permute(0, 2, 1, 3).reshape(M2, D)
    tab2 = ops.rope_table(8, hd // 4)
    out2 = ops.gemm(x2.cuda(), w2.cuda(), ops.EPI_HEADNORM128_BF16, bias=b2.cuda(), qn=[t.cuda() for t in kn],
                    kn=[t.cuda() for t in kn], n_q_cols=D, n_k_cols=D, rope_mode=ops.ROPE_2D, rope_tab=tab2,
                    tokens_per_frame=P, n_special=nsp, grid_w=gw).cpu()
    assert rel_l2(out2[:, :D], refq) < 5e-3


def test_gemm_argument_errors():
    from lsvs_b200 import ops
    a, w = torch.zeros(4, 100, dtype=torch.bfloat16, device="cuda"), torch.zeros(128, 100, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        ops.gemm(a, w, ops.EPI_BIAS_BF16)  # K % 64 != 0


@pytest.mark.parametrize("M,N,K", [(1648, 1024, 4096), (2060, 1024, 1024), (13184, 1024, 1024), (700, 1024, 2048)])
def test_gemm_residual_tma_reduce_and_split_k(M, N, K):
    """LayerScale + residual epilogue of the pair kernel (TMA reduce-add into the fp32 stream), including the short-chunk
    shapes where K is sliced over CTA pairs (28 / 36 tiles on 74 pairs) and every slice adds its partial; bias counted once."""
    from lsvs_b200 import ops
    a, w, bias = mk(M, N, K, seed=M % 97)
    resid, gamma = rnd(21, M, N), 0.1 * (1 + 0.1 * rnd(22, N))
    r_dev = resid.cuda()
    ops.gemm(a.cuda(), w.cuda(), ops.EPI_RESID_F32, bias=bias.cuda(), gamma=gamma.cuda(), resid=r_dev)
    ref = resid.double() + gamma.double() * (a.double() @ w.double().T + bias.double())
    assert rel_l2(r_dev.cpu().double(), ref) < 1e-5
    # the update itself (not hidden behind the residual's magnitude)
    assert rel_l2(r_dev.cpu().double() - resid.double(), ref - resid.double()) < 1e-4


@pytest.mark.parametrize("M,N,K", [(32, 2048, 2048), (32, 6144, 2048), (32, 8192, 2048), (32, 2048, 8192), (5, 2048, 2048), (1, 1024, 6144),
                                   (64, 1024, 512), (100, 256, 1024), (128, 384, 192), (33, 128, 64)])
def test_gemm_few_rows_weight_streaming(M, N, K):
    """M <= 128 (camera-head trunk: the S frames of a chunk x 2048 channels): gemm_fewrows_tcgen05 — swapped operands, the K slices
    of a 128-row weight tile summed over a thread-block cluster in slice order.  Every epilogue of that path vs the fp32 oracle,
    rows past M untouched, and two launches bit-identical (no atomics)."""
    from lsvs_b200 import ops
    a, w, bias = mk(M, N, K, seed=M + N)
    lin = a.float() @ w.float().T + bias
    ad, wd, bd = a.cuda(), w.cuda(), bias.cuda()
    guard = torch.full((M + 3, N), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ad, wd, ops.EPI_BIAS_BF16, bias=bd, out=guard[:M])
    assert rel_l2(guard[:M].cpu(), lin) < 4e-3 and float((guard[M:].float() - 7.0).abs().max()) == 0.0
    o2 = ops.gemm(ad, wd, ops.EPI_BIAS_BF16, bias=bd)
    assert torch.equal(o2, guard[:M])
    out = ops.gemm(ad, wd, ops.EPI_BIAS_F32, bias=bd).cpu()
    assert rel_l2(out, lin) < 1e-4
    out = ops.gemm(ad, wd, ops.EPI_BIAS_GELU_BF16, bias=bd).cpu()
    assert rel_l2(out, torch.nn.functional.gelu(lin)) < 4e-3
    resid, gamma = rnd(9, M, N), 0.1 * (1 + 0.1 * rnd(10, N))
    r1, r2 = resid.cuda(), resid.cuda()
    tap = torch.zeros(M, 2 * N, device="cuda")
    ops.gemm(ad, wd, ops.EPI_RESID_F32, bias=bd, gamma=gamma.cuda(), resid=r1, out2=tap[:, N:])
    ops.gemm(ad, wd, ops.EPI_RESID_F32, bias=bd, gamma=gamma.cuda(), resid=r2)
    assert rel_l2(r1.cpu(), resid + gamma * lin) < 1e-4
    assert torch.equal(r1, r2) and torch.equal(tap[:, N:], r1) and float(tap[:, :N].abs().max()) == 0.0

"""GPU parity: tcgen05 flash attention through the C ABI vs fp32 softmax attention on the CPU (bf16-rounded inputs).
P is rounded to bf16 before the PV product and O to bf16 on output -> rel-L2 <= 6e-3."""
import pytest
import torch

from conftest import rnd

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def ref_attention(qkv, B, H, hd, L):
    D = H * hd
    q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(B, L, H, hd).transpose(1, 2) for i in range(3))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    return o.transpose(1, 2).reshape(B * L, D)


@pytest.mark.parametrize("B,H,hd,L", [(1, 1, 64, 128), (1, 2, 64, 256), (2, 3, 64, 412), (1, 16, 64, 29), (3, 2, 64, 1),
                                      (1, 2, 64, 1648), (1, 1, 64, 700), (1, 1, 128, 64), (2, 2, 128, 413), (1, 16, 128, 32),
                                      (1, 2, 128, 300)])
def test_self_attention(B, H, hd, L):
    from lsvs_b200 import ops
    D = H * hd
    qkv = rnd(B * 1000 + L, B * L, 3 * D, scale=1.5).to(torch.bfloat16)
    ref = ref_attention(qkv, B, H, hd, L)
    dev = qkv.cuda()
    out = ops.attention(dev[:, :D], dev[:, D:2 * D], dev[:, 2 * D:], B, H, hd, L, L).cpu()
    assert out.shape == (B * L, D)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 6e-3, rel_l2(out, ref)


def test_cross_lengths_and_peaked_softmax():
    """Lq != Lk, and logits with a large dynamic range (running-max rescale path)."""
    from lsvs_b200 import ops
    B, H, hd, Lq, Lk = 2, 2, 64, 200, 333
    D = H * hd
    q = (rnd(1, B * Lq, D, scale=4.0)).to(torch.bfloat16)
    k = (rnd(2, B * Lk, D, scale=4.0)).to(torch.bfloat16)
    v = rnd(3, B * Lk, D).to(torch.bfloat16)
    qf, kf, vf = (t.float().reshape(B, -1, H, hd).transpose(1, 2) for t in (q, k, v))
    ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * Lq, D)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), B, H, hd, Lq, Lk).cpu()
    assert rel_l2(out, ref) < 8e-3, rel_l2(out, ref)


def test_global_size_properties():
    """BASELINE config-2 global attention size (1 x 13184 tokens, 16 heads): rows of softmax sum to one, so with
    V = const the output equals that constant; and a permutation of the keys leaves the output unchanged."""
    from lsvs_b200 import ops
    H, hd, L = 16, 64, 13184
    D = H * hd
    g = torch.Generator("cuda").manual_seed(0)
    q = torch.randn(L, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(L, D, device="cuda", generator=g).bfloat16()
    v = torch.full((L, D), 0.75, device="cuda").bfloat16()
    out = ops.attention(q, k, v, 1, H, hd, L, L)
    assert float((out.float() - 0.75).abs().max()) < 0.01
    v2 = torch.randn(L, D, device="cuda", generator=g).bfloat16()
    o1 = ops.attention(q, k, v2, 1, H, hd, L, L)
    perm = torch.randperm(L, device="cuda", generator=g)
    o2 = ops.attention(q, k[perm].contiguous(), v2[perm].contiguous(), 1, H, hd, L, L)
    assert rel_l2(o2, o1) < 6e-3
    # spot-check 64 query rows against fp32 math on the GPU box's CPU
    rows = torch.arange(0, L, L // 64)[:64]
    qf = q[rows].float().cpu().reshape(64, H, hd).transpose(0, 1)
    kf, vf = (t.float().cpu().reshape(L, H, hd).transpose(0, 1) for t in (k, v2))
    ref = torch.softmax(qf @ kf.transpose(1, 2) / 8.0, dim=-1) @ vf
    assert rel_l2(o1[rows].cpu(), ref.transpose(0, 1).reshape(64, D)) < 6e-3


@pytest.mark.parametrize("Lk", [700, 2600])  # one-tile (Lk <= 1024) and two-tile kernels
def test_rising_and_plateau_scores(Lk):
    """the reference maximum is set by the first key block and only raised when a row sum signals an element 2^8 above it:
    (a) scores that rise by ~50 per key block (a rescale on every block), (b) a plateau 1.5 above the first block (row sums
    above 2^8 without any jump: the max pass runs, nothing is rescaled)."""
    from lsvs_b200 import ops
    H, hd, Lq = 2, 64, 300
    D = H * hd
    g = torch.Generator().manual_seed(Lk)
    u = torch.nn.functional.normalize(torch.randn(H, hd, generator=g), dim=-1) * 8.0
    q = (u[None] + 0.05 * torch.randn(Lq, H, hd, generator=g)).reshape(Lq, D).bfloat16()
    ramp = (torch.arange(Lk).float() * 0.05).view(Lk, 1, 1)
    k_rise = (u[None] / 8.0 * ramp + 0.05 * torch.randn(Lk, H, hd, generator=g)).reshape(Lk, D).bfloat16()
    plateau = torch.where(torch.arange(Lk) < 128, 0.0, 1.5).view(Lk, 1, 1)
    k_plat = (u[None] / 8.0 * plateau + 0.02 * torch.randn(Lk, H, hd, generator=g)).reshape(Lk, D).bfloat16()
    v = torch.randn(Lk, D, generator=g).bfloat16()
    for k in (k_rise, k_plat):
        qf, kf, vf = (t.float().reshape(1, -1, H, hd).transpose(1, 2) for t in (q, k, v))
        ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(Lq, D)
        out = ops.attention(q.cuda(), k.cuda(), v.cuda(), 1, H, hd, Lq, Lk).cpu()
        assert torch.isfinite(out.float()).all()
        assert rel_l2(out, ref) < 8e-3, rel_l2(out, ref)


@pytest.mark.parametrize("L,rising", [(2560, False), (2500, True), (2500, False)])
def test_split_key_ranges_are_merged(L, rising):
    """320 (head, query tile) units on 296 CTA slots: the 24 units of the partial last round are cut into key ranges, written as
    unnormalised partials and merged (attention_combine_kernel).  With scores that rise along the keys every range has its own
    reference maximum and the last range carries almost all of the weight."""
    from lsvs_b200 import ops
    H, hd = 16, 64
    D = H * hd
    g = torch.Generator().manual_seed(L + rising)
    if rising:
        u = torch.nn.functional.normalize(torch.randn(H, hd, generator=g), dim=-1) * 8.0
        q = (u[None] + 0.05 * torch.randn(L, H, hd, generator=g)).reshape(L, D).bfloat16()
        ramp = (torch.arange(L).float() * 0.02).view(L, 1, 1)
        k = (u[None] / 8.0 * ramp + 0.05 * torch.randn(L, H, hd, generator=g)).reshape(L, D).bfloat16()
    else:
        q = (1.5 * torch.randn(L, D, generator=g)).bfloat16()
        k = (1.5 * torch.randn(L, D, generator=g)).bfloat16()
    v = torch.randn(L, D, generator=g).bfloat16()
    qf, kf, vf = (t.float().reshape(1, L, H, hd).transpose(1, 2) for t in (q, k, v))
    ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(L, D)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), 1, H, hd, L, L).cpu()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 8e-3, rel_l2(out, ref)
    # the units of the last round (head 15, query tiles 16..19 at least) on their own
    tail = slice(16 * 128, L)
    assert rel_l2(out[tail, 15 * hd:], ref[tail, 15 * hd:]) < 8e-3


@pytest.mark.parametrize("B,H,hd,Lq,Lk", [(413, 8, 128, 32, 9), (40, 8, 128, 32, 32), (7, 8, 128, 5, 2), (9, 8, 128, 64, 17), (3, 16, 64, 33, 64)])
def test_temporal_cross_attention_shapes(B, H, hd, Lq, Lk):
    """The alignment head's temporal cross attention (cross_attention.py:65-73): many (group, head) pairs of a few queries against
    a few keys (mostly padding on the 128-row tile: inactive warps, narrow last key block)."""
    from lsvs_b200 import ops
    D = H * hd
    q = rnd(11 * Lq + Lk, B * Lq, D, scale=1.5).to(torch.bfloat16)
    kv = rnd(13 * Lq + Lk, B * Lk, 2 * D, scale=1.5).to(torch.bfloat16)
    sh = lambda t, L: t.float().reshape(B, L, H, hd).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(sh(q, Lq), sh(kv[:, :D], Lk), sh(kv[:, D:], Lk)).transpose(1, 2).reshape(B * Lq, D)
    kvd = kv.cuda()
    out = ops.attention(q.cuda(), kvd[:, :D], kvd[:, D:], B, H, hd, Lq, Lk).cpu()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 4e-3, rel_l2(out, ref)

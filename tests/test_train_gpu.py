"""Training path of the alignment head (lsvs_b200/train.py): gradients of the native autograd nodes and of
the whole head + pose chain against the oracle's own autograd (fp32 torch on the CPU).

Tolerances: with fp32-class GEMM operands (precision 1) gradients agree with the fp32 oracle to ~1e-4 relative; with bf16 operands
(precision 0, the reference's bf16-mixed training arithmetic) to the bf16 level."""
import pytest
import torch

from conftest import rnd
from oracle import aligned as OA
from parity_util import load_synth_weights, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_on():
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(True)     # other test modules switch autograd off process-wide
    yield
    torch.set_grad_enabled(prev)


@pytest.mark.parametrize("precision,tol", [(1, 3e-5), (0, 1e-2)])
def test_linear_node_gradients(precision, tol):
    from lsvs_b200 import train
    M, K, N = 413, 1024, 384
    x = rnd(1, 3, M, K).cuda().requires_grad_(True)
    W = (rnd(2, N, K) * 0.05).cuda().requires_grad_(True)
    b = rnd(3, N).cuda().requires_grad_(True)
    g = rnd(4, 3, M, N).cuda()
    y = train.linear(x, W, b, precision)
    y.backward(g)
    xr, Wr, br = (t.detach().double().requires_grad_(True) for t in (x, W, b))
    yr = torch.nn.functional.linear(xr, Wr, br)
    yr.backward(g.double())
    assert rel_l2(y.detach(), yr.detach().float()) < tol
    for got, ref in ((x.grad, xr.grad), (W.grad, Wr.grad), (b.grad, br.grad)):
        assert rel_l2(got, ref.float()) < tol, (precision, rel_l2(got, ref.float()))


@pytest.mark.parametrize("hd,heads,B,Lq,Lk", [(128, 8, 2, 413, 413), (128, 8, 7, 32, 9), (64, 4, 1, 100, 37), (128, 2, 3, 5, 5)])
def test_attention_node_gradients(hd, heads, B, Lq, Lk):
    from lsvs_b200 import train
    D = heads * hd
    q, k, v = (rnd(10 + i, B * L, D).cuda().requires_grad_(True) for i, L in enumerate((Lq, Lk, Lk)))
    g = rnd(20, B * Lq, D).cuda()
    o = train.attention(q, k * 1.5, v, B, heads, hd, Lq, Lk)
    o.backward(g)
    qr, kr, vr = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    sh = lambda t, L: t.view(B, L, heads, hd).transpose(1, 2)
    orf = torch.nn.functional.scaled_dot_product_attention(sh(qr, Lq), sh(kr * 1.5, Lk), sh(vr, Lk)).transpose(1, 2).reshape(B * Lq, D)
    orf.backward(g.double())
    assert rel_l2(o.detach(), orf.detach().float()) < 1e-5
    for name, got, ref in (("dq", q.grad, qr.grad), ("dk", k.grad, kr.grad), ("dv", v.grad, vr.grad)):
        assert rel_l2(got, ref.float()) < 2e-5, (name, rel_l2(got, ref.float()))


def _oracle_loss(sd, toks, hw, ov, weights):
    """Two chained chunks through the oracle head (fp32, CPU, autograd) -> scalar loss, outputs."""
    ov_t = mem = None
    loss = 0.0
    outs = []
    for t, (w0, w1, w2, w3) in zip(toks, weights):
        sim3, se3, mem, ov_t = OA.alignment_head_forward(sd, "", t, hw, ov, ov_t, mem)
        loss = loss + (sim3 * w0).sum() + (se3 * w1).sum() + (mem * w2).sum() + (ov_t * w3).sum()
        outs.append((sim3, se3, mem, ov_t))
    return loss, outs


@pytest.mark.parametrize("precision,tol", [(1, 2e-3), (0, 1e-1)])
def test_alignment_head_gradients_vs_oracle_autograd(precision, tol):
    """Whole head, two chained chunks (back-propagation through the carried overlap tokens and memory): every parameter gradient vs
    the oracle's autograd.  Loss = fixed random linear functional of all four outputs of both chunks."""
    from aligned_vggt.heads.alignment_head import AlignmentHead
    from lsvs_b200 import train
    head = AlignmentHead()
    sd = load_synth_weights(head, seed=7, ls_gamma=0.2)
    head = head.cuda().train()
    S, gh, gw, ov = 4, 2, 3, 2
    P, hw = 5 + gh * gw, (gh * 14, gw * 14)
    toks = [rnd(30 + i, 1, S, P, 2048) for i in range(2)]
    shapes = [(1, 1, 8), (1, S - 1, 7), (1, 8, 512), (1, 1 + ov, P + 1, 1024)]
    weights = [[rnd(40 + 10 * c + j, *sh) for j, sh in enumerate(shapes)] for c in range(2)]
    # oracle: parameters as leaves
    sd_r = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_r, outs_r = _oracle_loss(sd_r, toks, hw, ov, weights)
    loss_r.backward()
    # drop-in training path
    ov_t = mem = None
    loss = 0.0
    for t, ws, (r0, r1, r2, r3) in zip(toks, weights, outs_r):
        sim3, se3, mem, ov_t = train.alignment_head_forward_train(head, t.cuda(), hw, ov, ov_t, mem, precision=precision)
        assert rel_l2(sim3.detach(), r0.detach()) < tol and rel_l2(ov_t.detach(), r3.detach()) < max(tol, 1e-2) * 0.5
        loss = loss + sum((o * w.cuda()).sum() for o, w in zip((sim3, se3, mem, ov_t), ws))
    loss.backward()
    worst, missing = 0.0, []
    for name, p in head.named_parameters():
        ref = sd_r[name].grad
        if ref is None or float(ref.norm()) == 0.0:
            continue
        if p.grad is None:
            missing.append(name)
            continue
        worst = max(worst, rel_l2(p.grad, ref))
    assert not missing, f"no gradient reached {missing[:5]}"
    assert worst < tol, worst


def test_model_training_step_reaches_only_the_head():
    """FeatureAlignedVGGT in train() mode with a frozen Aggregator / camera head (the reference's freeze list, *aggregator* ...): a
    loss on pose_enc + chunk Sim(3) + depth back-propagates into the alignment head through the differentiable pose chain; an
    optimizer step changes the next forward.  eval() / no_grad keep using the fused engine."""
    from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT
    torch.manual_seed(0)
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False, depth=1, patch_embed_depth=1,
                               intermediate_layer_indices=(0, 0, 0, 0)).cuda()
    for n, p in model.named_parameters():
        p.requires_grad_(n.startswith("alignment_head."))
    model.train()
    S, ov, H, W = 3, 1, 28, 42
    imgs = [torch.rand(1, S, 3, H, W, device="cuda") for _ in range(2)]
    dep = torch.rand(1, S, H, W, 1, device="cuda") + 0.5
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-2)

    def loss_of(preds):
        return sum(p.square().mean() for p in preds["pose_enc"]) + preds["chunk_sim3_alignment_enc"].square().mean() \
            + preds["frame_se3_alignment_enc"].square().mean() + sum(d.mean() for d in preds["depth"])
    p1 = model(imgs[0], ov, None, raw_depth=dep)
    p2 = model(imgs[1], ov, p1, raw_depth=dep)
    assert p2["pose_enc"][-1].requires_grad and p2["chunk_sim3_alignment_enc"].requires_grad
    loss = loss_of(p2)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert grads and all(n.startswith("alignment_head.") for n in grads)
    assert any(n.startswith("alignment_head.frame_blocks.0.attn.qkv") for n in grads) and "alignment_head.memory_token" in grads
    assert all(torch.isfinite(g).all() for g in grads.values())
    opt.step()
    with torch.no_grad():
        model.eval()
        q1 = model(imgs[0], ov, None, raw_depth=dep)
        assert not q1["pose_enc"][-1].requires_grad
        # the inference engine sees the updated parameters (version counters) and agrees with the training-path forward
        model.train()
    with torch.enable_grad():
        r1 = model(imgs[0], ov, None, raw_depth=dep)
    assert rel_l2(q1["chunk_sim3_alignment_enc"], r1["chunk_sim3_alignment_enc"].detach()) < 5e-2
    assert rel_l2(q1["pose_enc"][-1], r1["pose_enc"][-1].detach()) < 5e-2
    assert float((r1["chunk_sim3_alignment_enc"].detach() - p1["chunk_sim3_alignment_enc"].detach()).abs().max()) > 0

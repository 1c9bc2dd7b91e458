"""CPU: chunk index generation and the multi-rank chunk pipeline (gloo, world_size 2 and 3) with stand-in stage
functions: the pipelined / sharded execution must reproduce the sequential chunk loop exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lsvs_b200 import scheduler as sch
from oracle import aligned as OA


def test_generate_chunks_matches_oracle():
    for args in [(1000, "chunk_overlap", 32, 8), (14, "chunk_overlap", 5, 1), (3, "chunk_overlap", 5, 1), (33, "chunk_overlap", 32, 8),
                 (17, "chunk_gt", 5, 0), (9, "all", 4, 1), (75, "chunk_overlap", 75, 30), (1, "chunk_overlap", 5, 1)]:
        assert sch.generate_chunks(*args) == OA.generate_chunks(*args)
    ch = sch.generate_chunks(1000, "chunk_overlap", 32, 8)
    assert len(ch) == 42 and ch[-1] == list(range(984, 1000))
    with pytest.raises(ValueError):
        sch.generate_chunks(10, "two_chunks_typo", 4, 1)


def test_round_owners():
    assert sch.round_owners(0, 1, 0.1) == [0]
    # world 8, head 12 % of an aggregator pass: rank 0 only aligns
    assert all(sch.round_owners(j, 8, 0.125) == list(range(1, 8)) for j in range(10))
    # world 2, head cost 0.1: rank 0 encodes in 80 % of the rounds
    took = sum(0 in sch.round_owners(j, 2, 0.1) for j in range(100))
    assert took == 80
    assert all(set(sch.round_owners(j, 4, 0.1)) >= {1, 2, 3} for j in range(20))


# ---- stand-in stages: cheap deterministic arithmetic with a genuinely sequential context ------------
def _encode(inputs):
    x = inputs[0]
    return (x * 2.0 + 1.0), x.sum().reshape(1, 1)


def _align(tokens, cam, ctx):
    prev = torch.zeros(1, device=tokens.device) if ctx is None else ctx["state"]
    state = 0.5 * prev + tokens.mean().reshape(1) + cam.reshape(1)   # depends on every earlier chunk, in order
    packet = torch.cat([state, tokens.flatten()[:3]])
    return packet, {"state": state}


def _apply(packet, inputs):
    return packet[0] * inputs[1] + packet[1:4].sum()


def _sequential(n_rounds, world, head_cost):
    ctx, outs, k = None, [], 0
    for j in range(n_rounds):
        for o in sch.round_owners(j, world, head_cost):
            inp = _inputs(k)
            t, c = _encode(inp)
            packet, ctx = _align(t, c, ctx)
            outs.append((o, _apply(packet, inp)))
            k += 1
    return outs


def _inputs(k):
    g = torch.Generator().manual_seed(k)
    return (torch.randn(4, 5, generator=g), torch.randn(3, generator=g))


def _worker(rank, world, port, head_cost, n_rounds, q, lag=2, defer=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fwd, bwd = dist.new_group(), dist.new_group()
    pipe = sch.ChunkPipeline(_encode, _align, _apply, rank, world, head_cost=head_cost, packet_numel=4,
                             tokens_like=lambda: torch.empty(4, 5), cam_like=lambda: torch.empty(1, 1), fwd_group=fwd, bwd_group=bwd,
                             lag=lag, defer_chain=defer)
    k = 0
    for j in range(n_rounds):
        owners = pipe.owners()
        mine = None
        for o in owners:
            if o == rank:
                mine = _inputs(k)
            k += 1
        pipe.step(mine)
    res = pipe.flush()
    q.put((rank, [r.numpy().copy() for r in res]))     # numpy: nothing that needs this process alive when the parent unpickles
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,head_cost,lag,defer", [(2, 0.1, 2, True), (3, 0.4, 2, True), (2, 0.6, 1, False), (3, 0.2, 3, True),
                                                       (2, 0.3, 1, True), (3, 0.1, 2, False)])
def test_pipeline_equals_sequential_gloo(world, head_cost, lag, defer):
    """Every (lag, deferred-chain) setting only moves work in time: same results as the sequential loop, bit for bit."""
    n_rounds = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, head_cost, n_rounds, q, lag, defer)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _sequential(n_rounds, world, head_cost)
    for r in range(world):
        mine = [v for o, v in ref if o == r]
        assert len(mine) == len(got[r])
        for a, b in zip(mine, got[r]):
            assert torch.equal(a, torch.from_numpy(b))
    assert sum(len(v) for v in got.values()) == len(ref) == sum(len(sch.round_owners(j, world, head_cost)) for j in range(n_rounds))


def test_pipeline_single_rank_and_argument_errors():
    """world 1 needs no transport: the deferred chain hands results back one round late, flush() returns the rest."""
    pipe = sch.ChunkPipeline(_encode, _align, _apply, 0, 1)
    outs = []
    for k in range(4):
        pipe.step(_inputs(k))
        outs += pipe.results
        pipe.results = []
        assert len(outs) == k          # chunk k is chained at the start of round k+1
    outs += pipe.flush()
    ref = [v for _, v in _sequential(4, 1, 0.1)]
    assert len(outs) == 4 and all(torch.equal(a, b) for a, b in zip(outs, ref))
    with pytest.raises(ValueError):
        sch.ChunkPipeline(_encode, _align, _apply, 0, 1, lag=0)
    with pytest.raises(ValueError):
        sch.ChunkPipeline(_encode, _align, _apply, 0, 1).step(None)

    class FewSlots:
        slots = 2
    with pytest.raises(ValueError):
        sch.ChunkPipeline(_encode, _align, _apply, 1, 2, transport=FewSlots(), lag=2)


# ---- GPU: the CUDA-IPC mailbox transport (csrc/peer.cu) between two processes ------------------------------------------
def _peer_device(rank, mock):
    """CUDA device of this rank and the backend of the mailbox calls (None = liblsvs_b200.so); mock: host memory emulation."""
    if mock:
        import mock_peer
        return torch.device("cpu"), mock_peer
    dev = torch.device("cuda", rank)      # one GPU per rank: kernels that wait on a peer must never share a GPU with it
    torch.cuda.set_device(dev)
    return dev, None


def _need_gpus(world):
    """Ranks whose kernels wait on one another's flags must run on different GPUs (nothing guarantees that two processes'
    kernels are resident on one GPU at the same time; B200_PROFILING.md reports Xid 109 for exactly that), so these tests need
    `world` GPUs -- `gpurun --gpus N`; on a smaller box they skip and the protocol is covered on emulated peer memory instead."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (one per rank)")


def _peer_worker(rank, world, port, n_rounds, q, lag, mock=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)      # handshake only; payloads go through the mailboxes
    try:
        dev, backend = _peer_device(rank, mock)
        tx = sch.PeerTransport(rank, world, (4, 5), torch.float32, (1, 1), 4, dev, slots=lag + 1, timeout_s=60.0, backend=backend)
        pipe = sch.ChunkPipeline(_encode, _align, _apply, rank, world, head_cost=0.2, transport=tx, lag=lag)
        k = 0
        for j in range(n_rounds):
            mine = None
            for o in pipe.owners():
                if o == rank:
                    mine = tuple(t.to(dev) for t in _inputs(k))
                k += 1
            pipe.step(mine)
        res = [r.cpu().numpy() for r in pipe.flush()]
        tx.close()
        q.put((rank, res, None))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()))
        return
    dist.barrier()
    dist.destroy_process_group()


def _check_peer_mailboxes(world, lag, mock):
    n_rounds = 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, n_rounds, q, lag, mock)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    try:
        for _ in range(world):
            rank, res, err = q.get(timeout=120)
            assert err is None, f"rank {rank}: {err}"
            got[rank] = res
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:          # a failed rank must not leave its peer spinning on a mailbox
            if p.is_alive():
                p.terminate()
    dev = torch.device("cpu") if mock else torch.device("cuda", 0)
    ctx_, ref, k = None, {r: [] for r in range(world)}, 0
    for j in range(n_rounds):
        for o in sch.round_owners(j, world, 0.2):
            inp = tuple(t.to(dev) for t in _inputs(k))
            t, c = _encode(inp)
            packet, ctx_ = _align(t, c, ctx_)
            ref[o].append(_apply(packet, inp).cpu())
            k += 1
    for r in range(world):
        assert len(got[r]) == len(ref[r]) > 0
        for a, b in zip(got[r], ref[r]):
            assert torch.equal(torch.from_numpy(a), b)


@pytest.mark.gpu
@pytest.mark.parametrize("world,lag", [(2, 1), (2, 2), (3, 2), (4, 2)])
def test_peer_mailboxes_between_processes(world, lag):
    """`world` processes, one GPU each, exchange chunks and packets through the IPC mailboxes for more rounds than there are
    slots; results equal the sequential loop on the same device type.  world > 2: several owner ranks, i.e. one inbox per owner
    on rank 0."""
    _need_gpus(world)
    _check_peer_mailboxes(world, lag, mock=False)


@pytest.mark.parametrize("world,lag", [(2, 1), (3, 2)])
def test_peer_mailbox_protocol_on_host_memory(world, lag):
    """The same exchange with the lsvs_peer_* calls emulated over shared host memory (tests/mock_peer.py): layout, slot reuse and
    flag arithmetic of PeerTransport in the CPU suite."""
    _check_peer_mailboxes(world, lag, mock=True)


def _lost_peer_worker(rank, world, port, q):
    """Rank 1 owns chunks but never publishes them (a hung / dead producer).  Rank 0 must raise within a round of the timeout
    instead of chaining on a stale mailbox, and what it sends back must be poison so that rank 1 fails fast as well."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mock_peer
    tx = sch.PeerTransport(rank, world, (4, 5), torch.float32, (1, 1), 4, torch.device("cpu"), slots=3, timeout_s=0.5 if rank == 0 else 20.0,
                           backend=mock_peer)   # rank 1 would wait 20 s: only the poison can wake it earlier
    outcome = "no error"
    try:
        if rank == 0:
            pipe = sch.ChunkPipeline(_encode, _align, _apply, rank, world, head_cost=0.6, transport=tx, lag=2)   # rank 0 only aligns
            for j in range(3):
                pipe.step(None)
        else:
            tx.recv_packet(0)       # the answer to a chunk this rank never sent
            tx.poll()
    except mock_peer.NativeError as e:
        outcome = str(e)
    status_after = int(tx.status[0])
    q.put((rank, outcome, status_after))
    dist.barrier()
    tx.close()
    dist.destroy_process_group()


def test_lost_producer_raises_within_a_round_and_poisons_its_peers():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lost_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = {}
        for _ in range(2):
            rank, outcome, status_after = q.get(timeout=60)
            got[rank] = (outcome, status_after)
        for p in procs:
            p.join(timeout=30)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    assert "did not publish its message" in got[0][0], got      # rank 0: its wait for the chunk timed out -> raised from step()
    assert "reported a failure" in got[1][0], got               # rank 1: got poison instead of a packet built on stale tokens
    assert got[0][1] == 0 and got[1][1] == 0                    # the status word is cleared once the error has been raised


# ---- finite sequence with a short tail chunk (generate_chunks' last chunk, data.py:196-203): run_sequence ------------------
FRAMES = [4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2]     # 11 chunks; with 3 ranks the last round is cut to the chunks that remain


def _shapes_of(frames):
    return (frames, 5), (1, 1), 1 + frames


def _inputs_fr(k, dev="cpu"):
    g = torch.Generator().manual_seed(1000 + k)
    return (torch.randn(FRAMES[k], 5, generator=g).to(dev), torch.randn(3, generator=g).to(dev))


def _align_var(tokens, cam, ctx):
    prev = torch.zeros(1, device=tokens.device) if ctx is None else ctx["state"]
    state = 0.5 * prev + tokens.mean().reshape(1) + cam.reshape(1)
    return torch.cat([state, tokens[:, 0]]), {"state": state}     # packet length follows the chunk length


def _apply_var(packet, inputs):
    assert packet.numel() == 1 + inputs[0].shape[0]
    return packet[0] * inputs[1] + packet[1:].sum()


def _sequence_reference(dev="cpu"):
    ctx, out = None, []
    for k in range(len(FRAMES)):
        inp = _inputs_fr(k, dev)
        t, c = _encode(inp)
        packet, ctx = _align_var(t, c, ctx)
        out.append(_apply_var(packet, inp).cpu())
    return out


def _sequence_worker(rank, world, port, q, use_peer, mock=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kw = dict(head_cost=0.2, lag=2, chunk_frames=FRAMES, shapes_of=_shapes_of)
        if use_peer:
            dev, backend = _peer_device(rank, mock)
            tx = sch.PeerTransport(rank, world, (4, 5), torch.float32, (1, 1), 5, dev, slots=3, timeout_s=60.0, backend=backend)
            pipe = sch.ChunkPipeline(_encode, _align_var, _apply_var, rank, world, transport=tx, **kw)
        else:
            dev = "cpu"
            pipe = sch.ChunkPipeline(_encode, _align_var, _apply_var, rank, world, packet_numel=5, tokens_like=lambda: torch.empty(4, 5),
                                     cam_like=lambda: torch.empty(1, 1), **kw)
        res = sch.run_sequence(pipe, lambda k: _inputs_fr(k, dev))
        if use_peer:
            assert sch.mailbox_self_check(pipe) is True          # one more patterned message per owner and direction
            assert sch.mailbox_self_check(pipe) is True          # (repeatable: sequence numbers and slots move on)
            tx.close()
        else:
            assert sch.mailbox_self_check(pipe) is None
        q.put((rank, [(k, v.cpu().numpy()) for k, v in res], None))     # numpy: no fd hand-over that could outlive this process
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()))
        return
    dist.barrier()
    dist.destroy_process_group()


def _check_sequence(world, use_peer, mock=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sequence_worker, args=(r, world, port, q, use_peer, mock)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    try:
        for _ in range(world):
            rank, res, err = q.get(timeout=120)
            assert err is None, f"rank {rank}: {err}"
            got[rank] = res
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    ref = _sequence_reference("cuda" if use_peer and not mock else "cpu")
    seen = sorted(k for res in got.values() for k, _ in res)
    assert seen == list(range(len(FRAMES)))                       # every chunk exactly once, tail included
    for res in got.values():
        assert [k for k, _ in res] == sorted(k for k, _ in res)
        for k, v in res:
            assert torch.equal(torch.from_numpy(v), ref[k]), k


@pytest.mark.parametrize("world", [1, 2, 3])
def test_run_sequence_with_short_tail_gloo(world):
    _check_sequence(world, use_peer=False)


@pytest.mark.gpu
def test_run_sequence_with_short_tail_peer_mailboxes():
    _need_gpus(3)
    _check_sequence(3, use_peer=True)


def test_run_sequence_with_short_tail_mailbox_protocol_on_host_memory():
    _check_sequence(3, use_peer=True, mock=True)


def test_finite_sequence_bookkeeping():
    pipe = sch.ChunkPipeline(_encode, _align_var, _apply_var, 1, 3, transport=object(), head_cost=0.2, chunk_frames=FRAMES, shapes_of=_shapes_of)
    dealt = [pipe.owners(r) for r in range(6)]
    assert sum(len(o) for o in dealt) == len(FRAMES) and dealt[-1] == [] and len(dealt[-2]) <= 3
    assert pipe.chunk_start(0) == 0 and pipe.chunks_in_rounds(6) == len(FRAMES)
    assert pipe._shapes(10) == ((2, 5), (1, 1), 3)
    with pytest.raises(ValueError):
        sch.run_sequence(sch.ChunkPipeline(_encode, _align, _apply, 0, 1), lambda k: None)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_config4_dealing(world):
    """BASELINE config 4 (1000 frames, 32-frame chunks, 8 overlap): 42 chunks, every output frame exactly once, each chunk dealt to
    exactly one rank; with 8 ranks the alignment rank encodes in about a third of the rounds (head_cost 0.085)."""
    chunks = sch.generate_chunks(1000, "chunk_overlap", 32, 8)
    frames = [len(c) for c in chunks]
    assert len(chunks) == 42 and frames[-1] == 16 and sum(frames) == 1328
    new = [c if i == 0 else c[8:] for i, c in enumerate(chunks)]
    assert [f for c in new for f in c] == list(range(1000))
    pipe = sch.ChunkPipeline(_encode, _align, _apply, 0, world, transport=object() if world > 1 else None, head_cost=0.085,
                             chunk_frames=frames)
    dealt, r = [], 0
    while pipe.chunk_start(r) < len(frames):
        dealt.append(pipe.owners(r))
        r += 1
    assert sum(len(o) for o in dealt) == 42 and pipe.owners(r) == []
    per_rank = [sum(o.count(k) for o in dealt) for k in range(world)]
    assert sum(per_rank) == 42 and all(n > 0 for n in per_rank)
    if world == 8:
        assert per_rank[0] in (1, 2) and max(per_rank[1:]) - min(per_rank[1:]) <= 1     # rank 0 mostly aligns
    assert pipe.shapes_of is None and pipe._shapes(41) == (None, None, None)


def test_model_stages_run_dpt_heads_on_the_owner(monkeypatch):
    """ModelStages with a stand-in model: without raw maps the owner runs the model's DPT heads at encode time and applies the
    chunk's Sim(3) to their outputs when the packet arrives; with raw maps given the heads are not called."""
    from aligned_vggt.utils import alignment as al
    monkeypatch.setattr(al, "apply_sim3_alignment_on_point_maps", lambda p, T, s: p * s.view(-1, 1, 1, 1, 1) + T[:, :3, 3].view(-1, 1, 1, 1, 3))
    monkeypatch.setattr(al, "scale_depth", lambda d, s: d * s.view(-1, 1, 1, 1, 1))
    S, H, W = 3, 28, 42
    calls = []

    class Head:
        def __init__(self, c):
            self.c = c

        def __call__(self, taps, images, patch_start_idx):
            calls.append(self.c)
            return torch.full((1, S, H, W, self.c), float(self.c)), torch.ones(1, S, H, W)

    class Model:
        intermediate_layer_indices = [0, 0, 0, 1]
        depth_head, point_head = Head(1), Head(3)

        def aggregator(self, images):
            return [torch.zeros(1, S, 11, 2048), torch.ones(1, S, 11, 2048)], 5

        def camera_head(self, toks):
            return [torch.zeros(1, S, 9)]

    # with the head prefix on the owner what travels is the head's token stream (alignment token + P tokens, 1024 wide, fp32)
    assert sch.ModelStages(Model(), 1, S, H, W, "cpu").shapes_of(2) == ((1, 2, 12, 1024), (1, 2, 9), 1 + 16 + 18 + 8 + 7)
    assert sch.ModelStages(Model(), 1, S, H, W, "cpu").tokens_like().shape == (1, S, 12, 1024)
    st = sch.ModelStages(Model(), 1, S, H, W, "cpu", head_prefix_on_owner=False)
    assert st.shapes_of(2) == ((1, 2, 11, 2048), (1, 2, 9), 1 + 16 + 18 + 8 + 7)
    packet = torch.zeros(st.packet_numel)
    packet[0] = 2.0                                      # scale
    packet[1:17] = torch.eye(4).flatten()
    packet[4] = 0.5                                      # T[0, 3]: x translation
    inputs = (torch.zeros(1, S, 3, H, W), None, None)
    tok, cam = st.encode(inputs)
    assert tok.dtype == torch.bfloat16 and float(tok.float().mean()) == 1.0 and sorted(calls) == [1, 3] and len(st._maps) == 1
    out = st.apply(packet, inputs)
    assert st._maps == {} and set(out) >= {"world_points", "depth", "world_points_conf", "depth_conf", "pose_enc"}
    assert torch.equal(out["depth"], torch.full((1, S, H, W, 1), 2.0)) and float(out["world_points"][..., 0].mean()) == 6.5
    calls.clear()
    given = (torch.zeros(1, S, 3, H, W), torch.ones(1, S, H, W, 3), torch.ones(1, S, H, W, 1))
    st.encode(given)
    out = st.apply(packet, given)
    assert calls == [] and "world_points_conf" not in out and float(out["depth"].mean()) == 2.0


def _self_check_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev, backend = _peer_device(rank, mock=True)
        tx = sch.PeerTransport(rank, world, (4, 5), torch.float32, (1, 1), 4, dev, slots=3, timeout_s=30.0, backend=backend)
        pipe = sch.ChunkPipeline(_encode, _align, _apply, rank, world, head_cost=0.2, transport=tx)
        clean = sch.mailbox_self_check(pipe)
        if rank == 2:                                   # rank 2 damages one element of its next chunk message
            send = tx.send_chunk

            def damaged(seq, tokens, cam, packet_numel=None):
                tokens = tokens.clone()
                tokens.view(-1)[7] += 1.0
                return send(seq, tokens, cam, packet_numel)
            tx.send_chunk = damaged
        q.put((rank, (clean, sch.mailbox_self_check(pipe)), None))
        tx.close()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()))
        return
    dist.barrier()
    dist.destroy_process_group()


def test_mailbox_self_check_detects_a_damaged_message():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_self_check_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    try:
        for _ in range(world):
            rank, res, err = q.get(timeout=120)
            assert err is None, f"rank {rank}: {err}"
            got[rank] = res
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    assert got[0] == (True, False)                      # rank 0 sees the damaged chunk of rank 2
    assert got[1] == (True, True) and got[2] == (True, True)

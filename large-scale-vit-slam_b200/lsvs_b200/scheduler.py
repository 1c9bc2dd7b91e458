"""Chunk scheduler: the reference's sequential chunk loop (training/run_model.py:326-338,
training/training_metrics.py:616-659, aligned_vggt/utils/data.py:155-225) re-cut for one process per GPU.

What shards and what does not: the Aggregator (+ camera head) of a chunk depends only on that chunk's
images, so chunks are dealt to ranks; the alignment head of chunk k needs the processed overlap tokens / memory /
aligned poses of chunk k-1, so it runs as one sequential chain on the alignment rank (rank 0).  The only data-path
exchange is point-to-point: owner -> rank 0 carries the chunk's last-layer tokens (bf16, which is what the head's
first GEMM consumes) and the 9-d camera encodings; rank 0 -> owner carries the decoded Sim(3) packet (< 3 KB).  The
owner then applies the Sim(3) to its own depth / point maps `lag` chunks late, so that the chain on rank 0 overlaps
the following rounds' Aggregator work everywhere else.

Because rank 0 also pays for every chunk's head, it is given proportionally fewer Aggregator chunks
(`head_cost` = head time / aggregator time): with world*head_cost >= 1 it only aligns.

Transport: on GPUs, CUDA-IPC mailboxes in the receiver's HBM filled by the copy engines over NVLink and ordered by
sequence flags (PeerTransport; include/lsvs_b200.h lsvs_peer_*), because a pending NCCL send/recv kernel busy-waits on
SMs that the persistent encoder kernels need; torch.distributed isend/irecv (DistTransport) remains for gloo on CPU
tensors (tests/test_scheduler.py drives this file with stand-in stage functions) and as the multi-node option.
"""
from collections import deque
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------ chunk indices
def generate_chunks(num_frames: int, mode: str, seq_width: int, overlap: int) -> List[List[int]]:
    """Drop-in for aligned_vggt/utils/data.py:155-207.  "two_chunks" (:192-204) draws from the `random` module like the
    reference: a random non-empty proper subset of the frames, in sampling order, and the remaining frames in index order."""
    out: List[List[int]] = []
    if mode == "chunk_gt":
        for i in range(0, num_frames - seq_width + 1, seq_width):
            out.append(list(range(i, i + seq_width)))
        if len(out) * seq_width < num_frames:
            out.append(list(range(len(out) * seq_width, num_frames)))
    elif mode == "chunk_overlap":
        if num_frames < seq_width:
            out.append(list(range(num_frames)))
        else:
            step = seq_width - overlap
            for i in range(0, num_frames - seq_width + 1, step):
                out.append(list(range(i, i + seq_width)))
            if len(out) * step < num_frames - overlap:
                out.append(list(range(len(out) * step, num_frames)))
    elif mode == "all":
        out = [list(range(num_frames))]
    elif mode == "two_chunks":
        if num_frames < 2:
            raise ValueError("Number of frames must be at least 2 for two_chunks mode.")
        if num_frames == 2:
            return [[0, 1]]
        import random
        first = random.sample(range(num_frames), random.randint(1, num_frames - 1))
        taken = set(first)
        out = [first, [i for i in range(num_frames) if i not in taken]]
    else:
        raise ValueError(f"Unknown sequence generation mode: {mode}")
    return out


def round_owners(round_idx: int, world: int, head_cost: float) -> List[int]:
    """Ranks that encode a chunk in round `round_idx`, in chunk order.  Ranks 1..world-1 always do; rank 0 (which
    also runs every head) takes a share a = max(0, 1 - world*head_cost) of the rounds, spread evenly."""
    if world == 1:
        return [0]
    share = max(0.0, 1.0 - world * head_cost)
    takes = int((round_idx + 1) * share) > int(round_idx * share)
    return ([0] if takes else []) + list(range(1, world))


# ------------------------------------------------------------------------------------------------ transports
class DistTransport:
    """Owner <-> alignment-rank exchange over torch.distributed isend/irecv: gloo on CPU tensors (unit tests) or NCCL.
    On GPUs prefer PeerTransport: an unmatched NCCL send/recv kernel keeps SMs busy-waiting next to the persistent
    encoder kernels (see csrc/peer.cu)."""

    name = "torch.distributed p2p"

    def __init__(self, rank, world, tokens_like, cam_like, packet_numel, fwd_group=None, bwd_group=None, device=None):
        self.rank, self.world = rank, world
        self.tokens_like, self.cam_like, self.packet_numel = tokens_like, cam_like, packet_numel
        self.fwd, self.bwd, self.device = fwd_group, bwd_group, device
        self._packets = {}          # owner side: seq -> (work, buffer)
        self._inflight = deque()    # (work handles, tensors) of sends that must stay alive

    def _retire(self, keep):
        while len(self._inflight) > keep:
            works, _tensors = self._inflight.popleft()
            for w in works:
                w.wait()

    # owner side (shapes: None = the fixed full-size chunk; otherwise the shapes of a shorter chunk, e.g. a sequence's tail)
    def send_chunk(self, seq, tokens, cam, packet_numel=None):
        w1 = dist.isend(tokens, dst=0, group=self.fwd)
        w2 = dist.isend(cam, dst=0, group=self.fwd)
        self._inflight.append(([w1, w2], (tokens, cam)))
        packet = torch.empty(packet_numel or self.packet_numel, dtype=torch.float32, device=cam.device)
        self._packets[seq] = (dist.irecv(packet, src=0, group=self.bwd), packet)
        self._retire(keep=4)

    def recv_packet(self, seq, packet_numel=None):
        work, packet = self._packets.pop(seq)
        work.wait()
        return packet

    # alignment-rank side
    def recv_chunk(self, owner, seq, tokens_shape=None, cam_shape=None):
        t, c = self.tokens_like(), self.cam_like()
        if tokens_shape is not None:
            t, c = t.new_empty(tokens_shape), c.new_empty(cam_shape)
        w1 = dist.irecv(t, src=owner, group=self.fwd)
        w2 = dist.irecv(c, src=owner, group=self.fwd)
        w1.wait()
        w2.wait()
        return t, c

    def send_packet(self, owner, seq, packet):
        self._inflight.append(([dist.isend(packet, dst=owner, group=self.bwd)], (packet,)))
        self._retire(keep=2 * self.world)

    def finish(self):
        self._retire(keep=0)

    def close(self):
        pass


def _align256(n: int) -> int:
    return (n + 255) // 256 * 256


class _RawCudaBytes:
    """Exposes raw device memory to torch (zero copy) through the CUDA array interface."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerTransport:
    """Owner <-> alignment-rank exchange through CUDA-IPC mailboxes in the *receiver's* HBM (include/lsvs_b200.h, lsvs_peer_*):
    payloads are pushed over NVLink by the copy engines and ordered by sequence numbers, so no SM is ever held waiting
    for a peer and nobody rendezvous.

    Layout.  Rank 0 owns one inbox per owner rank: [flag | slots x (cam, tokens)]; every owner rank owns one mailbox:
    [flag | slots x packet]; slots are sized for the largest chunk, a shorter chunk (a sequence's tail) fills a prefix.
    Flags hold `number of messages published so far`.  A slot is reused every `slots`
    messages; that is safe when slots >= lag + 1 (lag = how many of its own chunks an owner keeps in flight before it
    waits for a packet), because receiving packet k proves that rank 0 has consumed chunk k, and an owner waits for
    packet seq - lag - 1 (stream order) before it overwrites anything belonging to message seq - slots."""

    name = "CUDA-IPC peer mailboxes (copy engines + sequence flags)"

    def __init__(self, rank, world, tokens_shape, tokens_dtype, cam_shape, packet_numel, device, group=None, slots=3, timeout_s=120.0,
                 backend=None):
        """backend: object with lib() / stream_ptr() / check() / NativeError like lsvs_b200.native (the default); the CPU unit
        tests pass an emulation of the lsvs_peer_* calls over shared host memory to exercise this class without a GPU."""
        import ctypes
        from . import native
        native = backend or native
        self._on_cuda = torch.device(device).type == "cuda"
        self.rank, self.world, self.slots, self.timeout_s, self.device = rank, world, slots, float(timeout_s), device
        self.tokens_shape, self.tokens_dtype, self.cam_shape, self.packet_numel = tuple(tokens_shape), tokens_dtype, tuple(cam_shape), packet_numel
        self._ct, self._native, self._lib = ctypes, native, native.lib()
        if backend is None:  # ctypes prototypes of the entry points called with plain Python integers
            self._lib.lsvs_peer_wait.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
            self._lib.lsvs_peer_signal.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
            self._lib.lsvs_peer_put.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
            self._lib.lsvs_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        self.tok_bytes = int(torch.tensor([], dtype=tokens_dtype).element_size())
        for d in self.tokens_shape:
            self.tok_bytes *= d
        self.cam_bytes = 4
        for d in self.cam_shape:
            self.cam_bytes *= d
        self.cam_stride = _align256(self.cam_bytes)
        self.chunk_stride = self.cam_stride + _align256(self.tok_bytes)
        self.packet_stride = _align256(4 * packet_numel)
        # health word (include/lsvs_b200.h): pinned host memory on a GPU, so that step() can read it without a device sync
        self.status = torch.zeros(1, dtype=torch.int32)
        if self._on_cuda:
            self.status = self.status.pin_memory()
        self._local, self._mapped = {}, {}     # peer rank -> base pointer (local allocation / mapped remote allocation)
        # two collective phases, each followed by an exchange of outcomes, so that a failure on one rank (IPC not
        # permitted, out of memory) raises on every rank instead of leaving the others waiting
        handles, err = {}, None
        try:
            if rank == 0:
                for o in range(1, world):
                    self._local[o] = self._alloc(256 + slots * self.chunk_stride)
                    handles[o] = self._export(self._local[o])
            else:
                self._local[0] = self._alloc(256 + slots * self.packet_stride)
                handles[0] = self._export(self._local[0])
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        table = [None] * world
        dist.all_gather_object(table, (handles, err), group=group)
        self._raise_any([t[1] for t in table])
        try:
            if rank == 0:
                for o in range(1, world):
                    self._mapped[o] = self._open(table[o][0][0])
            else:
                self._mapped[0] = self._open(table[0][0][rank])
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        outcome = [None] * world
        dist.all_gather_object(outcome, err, group=group)
        self._raise_any(outcome)
        self._views, self._zero_copy = {}, self._on_cuda

    def _raise_any(self, errors):
        errors = [e for e in errors if e]
        if errors:
            for p in self._local.values():
                self._lib.lsvs_peer_free(self._ct.c_void_p(p))
            self._local = {}
            raise self._native.NativeError("peer mailbox setup failed: " + "; ".join(errors))

    # -- raw memory helpers -----------------------------------------------------------------------
    def _alloc(self, nbytes):
        p = self._ct.c_void_p()
        self._native.check(self._lib.lsvs_peer_alloc(nbytes, self._ct.byref(p)), "lsvs_peer_alloc")
        return int(p.value)

    def _export(self, ptr):
        buf = (self._ct.c_ubyte * 64)()
        self._native.check(self._lib.lsvs_peer_export(self._ct.c_void_p(ptr), buf), "lsvs_peer_export")
        return bytes(buf)

    def _open(self, handle):
        p = self._ct.c_void_p()
        buf = (self._ct.c_ubyte * 64).from_buffer_copy(handle)
        self._native.check(self._lib.lsvs_peer_open(buf, self._ct.byref(p)), "lsvs_peer_open")
        return int(p.value)

    def _put(self, dst, src_tensor, nbytes):
        self._native.check(self._lib.lsvs_peer_put(dst, src_tensor.data_ptr(), nbytes, self._native.stream_ptr()), "lsvs_peer_put")

    def _signal(self, flag, value):
        self._native.check(self._lib.lsvs_peer_signal(flag, value & 0xFFFFFFFF, self.status.data_ptr(), self._native.stream_ptr()),
                           "lsvs_peer_signal")

    def _wait(self, flag, value):
        self._native.check(self._lib.lsvs_peer_wait(flag, value & 0xFFFFFFFF, self.status.data_ptr(), self.timeout_s,
                                                    self._native.stream_ptr()), "lsvs_peer_wait")

    def _read(self, ptr, nbytes, dtype, shape, private=False):
        """Torch tensor over local mailbox memory: a zero-copy alias (a private copy if `private`); if this torch build
        cannot alias raw device memory, a copy-engine copy into a fresh tensor, in stream order."""
        if self._zero_copy:
            try:
                key = (ptr, dtype, tuple(shape))
                if key not in self._views:
                    raw = torch.as_tensor(_RawCudaBytes(ptr, nbytes), device=self.device)
                    if raw.data_ptr() != ptr:
                        raise RuntimeError("torch copied the mailbox instead of aliasing it")
                    self._views[key] = raw.view(dtype).view(shape)
                return self._views[key].clone() if private else self._views[key]
            except Exception as e:  # noqa: BLE001
                import sys
                print(f"[lsvs_b200] mailbox views fall back to staged copies: {e}", file=sys.stderr, flush=True)
                self._zero_copy = False
        out = torch.empty(shape, dtype=dtype, device=self.device)
        self._native.check(self._lib.lsvs_peer_put(out.data_ptr(), ptr, nbytes, self._native.stream_ptr()), "lsvs_peer_put")
        return out

    @staticmethod
    def _nbytes(shape, itemsize):
        n = itemsize
        for d in shape:
            n *= d
        return n

    # -- owner side ------------------------------------------------------------------------------
    def send_chunk(self, seq, tokens, cam, packet_numel=None):
        assert tokens.is_contiguous() and cam.is_contiguous() and tokens.dtype == self.tokens_dtype and cam.dtype == torch.float32
        tok_bytes, cam_bytes = tokens.numel() * tokens.element_size(), cam.numel() * 4
        assert 0 < tok_bytes <= self.tok_bytes and 0 < cam_bytes <= self.cam_bytes, "chunk larger than the mailbox slot"
        base = self._mapped[0] + 256 + (seq % self.slots) * self.chunk_stride
        self._put(base, cam, cam_bytes)
        self._put(base + self.cam_stride, tokens, tok_bytes)
        self._signal(self._mapped[0], seq + 1)

    def recv_packet(self, seq, packet_numel=None):
        n = packet_numel or self.packet_numel
        assert n <= self.packet_numel
        self._wait(self._local[0], seq + 1)
        ptr = self._local[0] + 256 + (seq % self.slots) * self.packet_stride
        return self._read(ptr, 4 * n, torch.float32, (n,), private=True)

    # -- alignment-rank side ---------------------------------------------------------------------
    def recv_chunk(self, owner, seq, tokens_shape=None, cam_shape=None):
        tokens_shape, cam_shape = tuple(tokens_shape or self.tokens_shape), tuple(cam_shape or self.cam_shape)
        tok_bytes = self._nbytes(tokens_shape, self.tok_bytes // max(1, self._nbytes(self.tokens_shape, 1)))
        cam_bytes = self._nbytes(cam_shape, 4)
        assert tok_bytes <= self.tok_bytes and cam_bytes <= self.cam_bytes
        self._wait(self._local[owner], seq + 1)
        base = self._local[owner] + 256 + (seq % self.slots) * self.chunk_stride
        cam = self._read(base, cam_bytes, torch.float32, cam_shape)
        tokens = self._read(base + self.cam_stride, tok_bytes, self.tokens_dtype, tokens_shape)
        return tokens, cam

    def send_packet(self, owner, seq, packet):
        assert packet.is_contiguous() and packet.dtype == torch.float32 and packet.numel() <= self.packet_numel
        self._put(self._mapped[owner] + 256 + (seq % self.slots) * self.packet_stride, packet, 4 * packet.numel())
        self._signal(self._mapped[owner], seq + 1)

    def poll(self):
        """Non-blocking health check (a plain read of the pinned status word): raises once a wait of this rank has given up — from
        then on this rank publishes poison instead of results (lsvs_peer_signal), so its peers raise within one message too.  The
        word is cleared when the error is raised; the transport itself must be rebuilt (its sequence numbers are out of step)."""
        code = int(self.status[0])
        if code != 0:
            self.status[0] = 0
            why = {1: f"a peer did not publish its message within {self.timeout_s} s", 2: "a peer reported a failure (poisoned mailbox)"}
            raise self._native.NativeError(f"rank {self.rank}: {why.get(code, f'transport status {code}')}")

    def finish(self):
        """Host-synchronising health check: waits for the stream's pending waits, then raises like poll()."""
        if self._on_cuda:
            torch.cuda.current_stream().synchronize()
        self.poll()

    def close(self, group=None):
        """Collective: unmap the peers' buffers, then free the local ones."""
        if self._on_cuda:
            torch.cuda.synchronize()
        dist.barrier(group=group)
        for p in self._mapped.values():
            self._lib.lsvs_peer_close(self._ct.c_void_p(p))
        self._mapped, self._views = {}, {}
        dist.barrier(group=group)
        for p in self._local.values():
            self._lib.lsvs_peer_free(self._ct.c_void_p(p))
        self._local = {}


# ------------------------------------------------------------------------------------------------ pipeline
class ChunkPipeline:
    """SPMD driver: every rank calls step() once per round with the inputs of the chunk it owns in that round
    (or None).  Stage functions:
        encode_fn(inputs)                -> (tokens, cam)      context-free part (Aggregator + camera head)
        align_fn(tokens, cam, ctx)       -> (packet, new_ctx)  sequential part, rank 0 only; packet is a flat fp32 tensor
        apply_fn(packet, inputs)         -> result             owner side (Sim(3) application)

    Timing structure.  `defer_chain`: rank 0 runs the alignment chain of round r-1 at the start of round r, when those
    tokens have long arrived, and only then encodes its own chunk of round r, so its stream never idles waiting for the
    other ranks' encoders.  `lag`: an owner applies the Sim(3) packet of its chunk k only while stepping chunk k+lag, so a
    round in which rank 0 both encodes and aligns (longer than everybody else's) is absorbed instead of stalling the
    owners.  Results are independent of both (tests/test_scheduler.py compares with the sequential loop bit for bit)."""

    def __init__(self, encode_fn: Callable, align_fn: Callable, apply_fn: Callable, rank: int, world: int, *,
                 head_cost: float = 0.1, packet_numel: int = 0, tokens_like: Callable = None, cam_like: Callable = None,
                 device=None, fwd_group=None, bwd_group=None, transport=None, lag: int = 2, defer_chain: bool = True,
                 chunk_frames: Optional[Sequence[int]] = None, shapes_of: Optional[Callable] = None):
        """chunk_frames: frames of every chunk of a finite sequence, in chunk order (generate_chunks); the last round is then cut
        to the chunks that remain, and chunks may differ in length (the short tail chunk, data.py:196-203) if
        shapes_of(frames) -> (tokens shape, cam shape, packet numel) is given."""
        self.encode_fn, self.align_fn, self.apply_fn = encode_fn, align_fn, apply_fn
        self.rank, self.world, self.head_cost = rank, world, head_cost
        self.device = device
        if transport is None and world > 1:
            transport = DistTransport(rank, world, tokens_like, cam_like, packet_numel, fwd_group, bwd_group, device)
        self.tx = transport
        if lag < 1:
            raise ValueError("lag must be >= 1")
        if getattr(transport, "slots", lag + 1) < lag + 1:
            raise ValueError(f"transport has {transport.slots} slots; lag {lag} needs {lag + 1}")
        self.lag, self.defer = lag, defer_chain
        self.chunk_frames = None if chunk_frames is None else [int(f) for f in chunk_frames]
        self.shapes_of = shapes_of
        self._starts = [0]                # _starts[r] = global index of the first chunk of round r
        self.round = 0
        self.ctx = None
        self.seq = 0                      # owner side: chunks this rank has encoded
        self.pending = deque()            # owner side: (seq, inputs, global chunk index) awaiting their Sim(3) packet
        self._own = {}                    # rank 0: round -> (tokens, cam, inputs) of its own chunk, until chained
        self._oseq = [0] * world          # rank 0: chunks received per owner
        self._chained = 0                 # rank 0: rounds [0, _chained) have been through the chain
        self.results = []
        self.result_chunks = []           # global chunk index of every entry of `results`

    # -- bookkeeping -----------------------------------------------------------------------------
    def chunk_start(self, round_idx: int) -> int:
        """Global index of the first chunk of round `round_idx` (chunks are dealt round by round, owners in rank order)."""
        while len(self._starts) <= round_idx:
            r = len(self._starts) - 1
            self._starts.append(self._starts[r] + len(self.owners(r)))
        return self._starts[round_idx]

    def owners(self, round_idx: Optional[int] = None) -> List[int]:
        r = self.round if round_idx is None else round_idx
        ow = round_owners(r, self.world, self.head_cost)
        if self.chunk_frames is not None:
            ow = ow[:max(0, len(self.chunk_frames) - self.chunk_start(r))]
        return ow

    def owns(self, round_idx: Optional[int] = None) -> bool:
        return self.rank in self.owners(round_idx)

    def done(self) -> bool:
        """Finite sequences: every chunk has been dealt."""
        return self.chunk_frames is not None and self.chunk_start(self.round) >= len(self.chunk_frames)

    def chunks_in_rounds(self, n_rounds: int, start: int = 0) -> int:
        return self.chunk_start(start + n_rounds) - self.chunk_start(start)

    def _shapes(self, k: int):
        if self.chunk_frames is None or self.shapes_of is None:
            return None, None, None
        return self.shapes_of(self.chunk_frames[k])

    def _apply_ready(self, keep: int):
        while len(self.pending) > keep:
            seq, inputs, k = self.pending.popleft()
            self.results.append(self.apply_fn(self.tx.recv_packet(seq, self._shapes(k)[2]), inputs))
            self.result_chunks.append(k)

    def _chain_through(self, last_round: int):
        """Rank 0: run the alignment chain for every not yet chained round <= last_round, chunks in order."""
        while self._chained <= last_round:
            r = self._chained
            for idx, o in enumerate(self.owners(r)):
                k = self.chunk_start(r) + idx
                if o == 0:
                    t, c, inputs = self._own.pop(r)
                else:
                    tshape, cshape, _ = self._shapes(k)
                    t, c = self.tx.recv_chunk(o, self._oseq[o], tshape, cshape)
                packet, self.ctx = self.align_fn(t, c, self.ctx)
                if o == 0:
                    self.results.append(self.apply_fn(packet, inputs))
                    self.result_chunks.append(k)
                else:
                    self.tx.send_packet(o, self._oseq[o], packet)
                    self._oseq[o] += 1
            self._chained += 1

    # -- one round -------------------------------------------------------------------------------
    def step(self, inputs):
        owners = self.owners()
        mine = self.rank in owners
        if mine and inputs is None:
            raise ValueError(f"rank {self.rank} owns a chunk in round {self.round} but got no inputs")
        if self.rank == 0 and self.defer:
            self._chain_through(self.round - 1)
        if mine:
            k = self.chunk_start(self.round) + owners.index(self.rank)
            tokens, cam = self.encode_fn(inputs)
            if self.rank == 0:
                self._own[self.round] = (tokens, cam, inputs)
            else:
                self.tx.send_chunk(self.seq, tokens, cam, self._shapes(k)[2])
                self.pending.append((self.seq, inputs, k))
                self.seq += 1
        if self.rank == 0 and not self.defer:
            self._chain_through(self.round)
        # owners: apply the packet of the chunk encoded `lag` chunks ago (this round's encode is already queued ahead of it)
        self._apply_ready(keep=self.lag)
        self.round += 1
        if self.tx is not None and hasattr(self.tx, "poll"):
            self.tx.poll()  # a timed-out / poisoned mailbox surfaces within a round, not at the end of the sequence

    def flush(self, with_chunk_ids: bool = False):
        if self.rank == 0:
            self._chain_through(self.round - 1)
        self._apply_ready(keep=0)
        if self.tx is not None:
            self.tx.finish()
        out, ids = self.results, self.result_chunks
        self.results, self.result_chunks = [], []
        return list(zip(ids, out)) if with_chunk_ids else out


def mailbox_self_check(pipe: ChunkPipeline) -> Optional[bool]:
    """Integrity check of the mailbox transport, to be called by every rank right after pipe.flush(): each owner pushes one more
    full-size chunk message filled with a known bit pattern, rank 0 compares every element and answers with a patterned packet
    that the owner compares in turn.  True / False = everything this rank received was / was not bit-exact (host-synchronising);
    None when the pipeline does not run on PeerTransport."""
    tx = pipe.tx
    if not isinstance(tx, PeerTransport):
        return None

    def pattern(shape, dtype, salt):  # integers below 251: exact in bf16 as well
        n = 1
        for d in shape:
            n *= d
        return ((torch.arange(n, device=tx.device) + salt) % 251).to(dtype).view(tuple(shape))

    ok = True
    if pipe.rank == 0:
        for o in range(1, pipe.world):
            seq = pipe._oseq[o]
            t, c = tx.recv_chunk(o, seq)
            ok = ok and torch.equal(t, pattern(tx.tokens_shape, tx.tokens_dtype, 7 * o + seq % 5))
            ok = ok and torch.equal(c, pattern(tx.cam_shape, torch.float32, 3 * o + 1))
            tx.send_packet(o, seq, pattern((tx.packet_numel,), torch.float32, 11 * o))
            pipe._oseq[o] += 1
    else:
        seq, o = pipe.seq, pipe.rank
        tx.send_chunk(seq, pattern(tx.tokens_shape, tx.tokens_dtype, 7 * o + seq % 5), pattern(tx.cam_shape, torch.float32, 3 * o + 1))
        ok = torch.equal(tx.recv_packet(seq), pattern((tx.packet_numel,), torch.float32, 11 * o))
        pipe.seq += 1
    tx.finish()
    return bool(ok)


def run_sequence(pipe: ChunkPipeline, load_chunk: Callable[[int], tuple]):
    """The reference's chunk loop (training/run_model.py:326-338, training_metrics.py:636-657) over a finite sequence on
    every rank of the pipeline: `pipe` was built with chunk_frames = [len(c) for c in generate_chunks(...)], load_chunk(k)
    returns the stage inputs of chunk k and is only called on the rank that owns it.  Returns [(chunk index, result)] of the
    chunks this rank owns, in chunk order."""
    if pipe.chunk_frames is None:
        raise ValueError("run_sequence needs a pipeline built with chunk_frames")
    out = []
    while not pipe.done():
        owners = pipe.owners()
        mine = pipe.chunk_start(pipe.round) + owners.index(pipe.rank) if pipe.rank in owners else None
        pipe.step(load_chunk(mine) if mine is not None else None)
        out += list(zip(pipe.result_chunks, pipe.results))
        pipe.results, pipe.result_chunks = [], []
    out += pipe.flush(with_chunk_ids=True)
    return sorted(out, key=lambda kv: kv[0])


# ------------------------------------------------------------------------------------------------ model binding
class ModelStages:
    """Stage functions of ChunkPipeline for a FeatureAlignedVGGT drop-in (batch size 1 per chunk).
    inputs = (images (1,S,3,H,W), raw_points (1,S,H,W,3) | None, raw_depth (1,S,H,W,1) | None); with None and a model that has its
    DPT heads, the owner computes the maps itself (results then also carry world_points_conf / depth_conf)."""

    def __init__(self, model, num_overlap: int, S: int, H: int, W: int, device, head_prefix_on_owner: bool = True):
        """head_prefix_on_owner: the owner of a chunk also runs the context-free prefix of the alignment head (project_in,
        token_norm, first frame block: 0.45 of the 4.2 ms the head costs the alignment rank per chunk) and ships the fp32 token
        stream (1,S,P+1,1024) — the same 54 MB as the bf16 tapped tokens it replaces; rank 0 resumes from it."""
        self.model, self.ov, self.S, self.H, self.W, self.device = model, num_overlap, S, H, W, device
        self.prefix = bool(head_prefix_on_owner)
        self.P = 5 + (H // 14) * (W // 14)
        self.packet_numel = 1 + 16 + S * 9 + 8 + (S - 1) * 7
        # the last-layer tokens travel as bf16 — exactly what the head's first GEMM consumes — unless the model runs a precision
        # mode whose head takes fp32-class operands
        self.tokens_dtype = torch.bfloat16 if not getattr(model, "precision", None) else torch.float32
        if self.prefix:
            self.tokens_dtype = torch.float32
        self._maps = {}   # id(inputs) -> DPT head outputs of a chunk whose packet has not arrived yet

    def shapes_of(self, frames: int):
        """(tokens shape, camera-encoding shape, packet numel) of a chunk of `frames` frames."""
        return self.tokens_shape(frames), (1, frames, 9), 1 + 16 + frames * 9 + 8 + (frames - 1) * 7

    def tokens_shape(self, frames: int):
        """what travels owner -> alignment rank: the head's prefix stream, or the last tapped Aggregator layer"""
        return (1, frames, self.P + 1, 1024) if self.prefix else (1, frames, self.P, 2048)

    def tokens_like(self):
        return torch.empty(self.tokens_shape(self.S), dtype=self.tokens_dtype, device=self.device)

    def cam_like(self):
        return torch.empty(1, self.S, 9, dtype=torch.float32, device=self.device)

    def encode(self, inputs):
        images = inputs[0]
        m = self.model
        tokens_list, patch_start_idx = m.aggregator(images)
        taps = [tokens_list[i] for i in m.intermediate_layer_indices]
        last = taps[-1]
        cam = m.camera_head([last])[-1]
        # DPT heads of the model (featureAligned_vggt.py:166-168, :183-185) run on the chunk's owner when no stand-in maps are given;
        # their outputs wait here for the chunk's Sim(3) packet
        maps = {}
        if inputs[2] is None and getattr(m, "depth_head", None) is not None:
            maps["depth"], maps["depth_conf"] = m.depth_head(taps, images=images, patch_start_idx=patch_start_idx)
        if inputs[1] is None and getattr(m, "point_head", None) is not None:
            maps["points"], maps["points_conf"] = m.point_head(taps, images=images, patch_start_idx=patch_start_idx)
        if maps:
            self._maps[id(inputs)] = maps
        if self.prefix:
            return m.alignment_head.forward_prefix(last, (self.H, self.W)), cam
        return last.to(self.tokens_dtype), cam

    def align(self, tokens, cam, ctx):
        from .engine import pose_chain
        m = self.model
        S = tokens.shape[1]
        overlap = self.ov if S > self.ov else S - 1
        ov_in = mem_in = prev = None
        if ctx is not None:
            ov_in, mem_in, prev = ctx["overlap_tokens"], ctx["memory_tokens"], ctx["pose_enc"]
        head = m.alignment_head.forward_from_prefix if self.prefix else m.alignment_head
        sim3, se3, mem, ov_out = head(tokens, (self.H, self.W), overlap, overlap_tokens=ov_in, memory_tokens=mem_in)
        pose, point_T, scale = pose_chain(sim3, se3, cam, prev, overlap, (self.H, self.W))
        packet = torch.cat([scale.reshape(-1), point_T.reshape(-1), pose.reshape(-1), sim3.reshape(-1), se3.reshape(-1)])
        return packet, {"overlap_tokens": ov_out, "memory_tokens": mem, "pose_enc": pose}

    def apply(self, packet, inputs):
        from aligned_vggt.utils import alignment as al
        S = inputs[0].shape[1]
        scale, T = packet[0:1], packet[1:17].view(1, 4, 4)
        out = {"pose_enc": packet[17:17 + S * 9].view(1, S, 9), "chunk_sim3_alignment_enc": packet[17 + S * 9:25 + S * 9].view(1, 1, 8),
               "frame_se3_alignment_enc": packet[25 + S * 9:].view(1, S - 1, 7)}
        maps = self._maps.pop(id(inputs), {})
        pts = inputs[1] if inputs[1] is not None else maps.get("points")
        dep = inputs[2] if inputs[2] is not None else maps.get("depth")
        if pts is not None:
            out["world_points"] = al.apply_sim3_alignment_on_point_maps(pts, T, scale)
        if dep is not None:
            out["depth"] = al.scale_depth(dep, scale)
        for src, dst in (("points_conf", "world_points_conf"), ("depth_conf", "depth_conf")):
            if src in maps:
                out[dst] = maps[src]
        return out


def model_pipeline(model, num_overlap: int, S: int, H: int, W: int, rank: int, world: int, device, head_cost: float = 0.075,
                   fwd_group=None, bwd_group=None, transport: str = "auto", lag: int = 2, defer_chain: bool = True,
                   handshake_group=None, chunk_frames: Optional[Sequence[int]] = None, head_prefix_on_owner: bool = True) -> ChunkPipeline:
    """transport: "peer" (CUDA-IPC mailboxes, one box), "dist" (torch.distributed isend/irecv) or "auto" (peer, and
    torch.distributed only if every rank agrees that the mailboxes could not be set up).  S = frames of the largest chunk;
    chunk_frames = per-chunk frame counts of a finite sequence (see run_sequence), None = endless rounds of S-frame chunks."""
    st = ModelStages(model, num_overlap, S, H, W, device, head_prefix_on_owner=head_prefix_on_owner)
    tx = None
    if world > 1 and transport in ("auto", "peer"):
        try:
            tx = PeerTransport(rank, world, st.tokens_shape(S), st.tokens_dtype, (1, S, 9), st.packet_numel, device,
                               group=handshake_group, slots=lag + 1)
        except Exception as e:  # noqa: BLE001  (raised on every rank together, see PeerTransport.__init__)
            if transport == "peer":
                raise
            import sys
            print(f"[lsvs_b200] peer mailboxes unavailable, using torch.distributed p2p: {e}", file=sys.stderr, flush=True)
    if world > 1 and tx is None:
        tx = DistTransport(rank, world, st.tokens_like, st.cam_like, st.packet_numel, fwd_group, bwd_group, device)
        if torch.device(device).type == "cuda":
            # NCCL serialises eager point-to-point calls per communicator; only the original schedule (chain in the same
            # round, packet needed one chunk later) has been measured to run on it
            lag, defer_chain = 1, False
    return ChunkPipeline(st.encode, st.align, st.apply, rank, world, head_cost=head_cost, packet_numel=st.packet_numel,
                         device=device, transport=tx, lag=lag, defer_chain=defer_chain, chunk_frames=chunk_frames,
                         shapes_of=st.shapes_of if chunk_frames is not None else None)

"""Chunk scheduler: the reference's sequential chunk loop (training/run_model.py:326-338,
training/training_metrics.py:616-659, aligned_vggt/utils/data.py:155-225) re-cut for one process per GPU.

What shards and what does not: the Aggregator (+ camera head) of a chunk depends only on that chunk's
images, so chunks are dealt to ranks; the alignment head of chunk k needs the processed overlap tokens / memory /
aligned poses of chunk k-1, so it runs as one sequential chain on the alignment rank (rank 0).  The only data-path
exchange is point-to-point: owner -> rank 0 carries the chunk's last-layer tokens (bf16, which is what the head's
first GEMM consumes) and the 9-d camera encodings; rank 0 -> owner carries the decoded Sim(3) packet (< 3 KB).  The
owner then applies the Sim(3) to its own depth / point maps, one round late, so that the chain on rank 0 overlaps
the next round's Aggregator work everywhere else.

Because rank 0 also pays for every chunk's head, it is given proportionally fewer Aggregator chunks
(`head_cost` = head time / aggregator time): with world*head_cost >= 1 it only aligns.

The transport is torch.distributed isend/irecv: NCCL over NVLink on GPUs, gloo on CPU tensors in the unit tests
(tests/test_scheduler.py drives this file with stand-in stage functions).
"""
from collections import deque
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------ chunk indices
def generate_chunks(num_frames: int, mode: str, seq_width: int, overlap: int) -> List[List[int]]:
    """Drop-in for aligned_vggt/utils/data.py:155-207 (deterministic modes)."""
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


class ChunkPipeline:
    """SPMD driver: every rank calls step() once per round with the inputs of the chunk it owns in that round
    (or None).  Stage functions:
        encode_fn(inputs)                -> (tokens, cam)      context-free part (Aggregator + camera head)
        align_fn(tokens, cam, ctx)       -> (packet, new_ctx)  sequential part, rank 0 only; packet is a flat fp32 tensor
        apply_fn(packet, inputs)         -> result             owner side (Sim(3) application)
    `tokens_like()` / `cam_like()` (fresh receive buffers) and `packet_numel` describe the transfers."""

    def __init__(self, encode_fn: Callable, align_fn: Callable, apply_fn: Callable, rank: int, world: int, *,
                 head_cost: float = 0.1, packet_numel: int = 0, tokens_like: Callable = None, cam_like: Callable = None,
                 device=None, fwd_group=None, bwd_group=None):
        self.encode_fn, self.align_fn, self.apply_fn = encode_fn, align_fn, apply_fn
        self.rank, self.world, self.head_cost = rank, world, head_cost
        self.packet_numel, self.tokens_like, self.cam_like = packet_numel, tokens_like, cam_like
        self.device = device
        self.fwd, self.bwd = fwd_group, bwd_group
        self.round = 0
        self.ctx = None
        self.pending = deque()   # (work, packet buffer, inputs) awaiting their Sim(3) packet
        self.inflight = deque()  # (work handles, tensors) of sends that must stay alive
        self.results = []

    # -- bookkeeping -----------------------------------------------------------------------------
    def owners(self, round_idx: Optional[int] = None) -> List[int]:
        return round_owners(self.round if round_idx is None else round_idx, self.world, self.head_cost)

    def owns(self, round_idx: Optional[int] = None) -> bool:
        return self.rank in self.owners(round_idx)

    def chunks_in_rounds(self, n_rounds: int, start: int = 0) -> int:
        return sum(len(round_owners(j, self.world, self.head_cost)) for j in range(start, start + n_rounds))

    def _retire_sends(self, keep: int):
        while len(self.inflight) > keep:
            works, _tensors = self.inflight.popleft()
            for w in works:
                w.wait()

    def _apply_ready(self, keep: int):
        while len(self.pending) > keep:
            work, packet, inputs = self.pending.popleft()
            if work is not None:
                work.wait()
            self.results.append(self.apply_fn(packet, inputs))

    # -- one round -------------------------------------------------------------------------------
    def step(self, inputs):
        owners = self.owners()
        mine = self.rank in owners
        if mine and inputs is None:
            raise ValueError(f"rank {self.rank} owns a chunk in round {self.round} but got no inputs")
        tokens = cam = None
        if mine:
            tokens, cam = self.encode_fn(inputs)
        if self.rank == 0:
            for o in owners:
                if o == 0:
                    t, c = tokens, cam
                else:
                    t, c = self.tokens_like(), self.cam_like()
                    w1 = dist.irecv(t, src=o, group=self.fwd)
                    w2 = dist.irecv(c, src=o, group=self.fwd)
                    w1.wait()
                    w2.wait()
                packet, self.ctx = self.align_fn(t, c, self.ctx)
                if o == 0:
                    self.pending.append((None, packet, inputs))
                else:
                    w = dist.isend(packet, dst=o, group=self.bwd)
                    self.inflight.append(([w], (packet,)))
        elif mine:
            w1 = dist.isend(tokens, dst=0, group=self.fwd)
            w2 = dist.isend(cam, dst=0, group=self.fwd)
            self.inflight.append(([w1, w2], (tokens, cam)))
            packet = torch.empty(self.packet_numel, dtype=torch.float32, device=cam.device)
            wr = dist.irecv(packet, src=0, group=self.bwd)
            self.pending.append((wr, packet, inputs))
        # apply the previous round's packet now (this round's encode is already queued ahead of the wait)
        self._apply_ready(keep=1)
        self._retire_sends(keep=2 * self.world)
        self.round += 1

    def flush(self):
        self._apply_ready(keep=0)
        self._retire_sends(keep=0)
        out, self.results = self.results, []
        return out


# ------------------------------------------------------------------------------------------------ model binding
class ModelStages:
    """Stage functions of ChunkPipeline for a FeatureAlignedVGGT drop-in (batch size 1 per chunk).
    inputs = (images (1,S,3,H,W), raw_points (1,S,H,W,3) | None, raw_depth (1,S,H,W,1) | None)."""

    def __init__(self, model, num_overlap: int, S: int, H: int, W: int, device):
        self.model, self.ov, self.S, self.H, self.W, self.device = model, num_overlap, S, H, W, device
        self.P = 5 + (H // 14) * (W // 14)
        self.packet_numel = 1 + 16 + S * 9 + 8 + (S - 1) * 7

    def tokens_like(self):
        return torch.empty(1, self.S, self.P, 2048, dtype=torch.bfloat16, device=self.device)

    def cam_like(self):
        return torch.empty(1, self.S, 9, dtype=torch.float32, device=self.device)

    def encode(self, inputs):
        images = inputs[0]
        tokens_list, _ = self.model.aggregator(images)
        last = tokens_list[self.model.intermediate_layer_indices[-1]]
        cam = self.model.camera_head([last])[-1]
        return last.to(torch.bfloat16), cam

    def align(self, tokens, cam, ctx):
        from .engine import pose_chain
        m = self.model
        S = tokens.shape[1]
        overlap = self.ov if S > self.ov else S - 1
        ov_in = mem_in = prev = None
        if ctx is not None:
            ov_in, mem_in, prev = ctx["overlap_tokens"], ctx["memory_tokens"], ctx["pose_enc"]
        sim3, se3, mem, ov_out = m.alignment_head(tokens.float(), (self.H, self.W), overlap, overlap_tokens=ov_in, memory_tokens=mem_in)
        pose, point_T, scale = pose_chain(sim3, se3, cam, prev, overlap, (self.H, self.W))
        packet = torch.cat([scale.reshape(-1), point_T.reshape(-1), pose.reshape(-1), sim3.reshape(-1), se3.reshape(-1)])
        return packet, {"overlap_tokens": ov_out, "memory_tokens": mem, "pose_enc": pose}

    def apply(self, packet, inputs):
        from aligned_vggt.utils import alignment as al
        S = self.S
        scale, T = packet[0:1], packet[1:17].view(1, 4, 4)
        out = {"pose_enc": packet[17:17 + S * 9].view(1, S, 9), "chunk_sim3_alignment_enc": packet[17 + S * 9:25 + S * 9].view(1, 1, 8),
               "frame_se3_alignment_enc": packet[25 + S * 9:].view(1, S - 1, 7)}
        if inputs[1] is not None:
            out["world_points"] = al.apply_sim3_alignment_on_point_maps(inputs[1], T, scale)
        if inputs[2] is not None:
            out["depth"] = al.scale_depth(inputs[2], scale)
        return out


def model_pipeline(model, num_overlap: int, S: int, H: int, W: int, rank: int, world: int, device, head_cost: float = 0.1,
                   fwd_group=None, bwd_group=None) -> ChunkPipeline:
    st = ModelStages(model, num_overlap, S, H, W, device)
    return ChunkPipeline(st.encode, st.align, st.apply, rank, world, head_cost=head_cost, packet_numel=st.packet_numel,
                         tokens_like=st.tokens_like, cam_like=st.cam_like, device=device, fwd_group=fwd_group, bwd_group=bwd_group)

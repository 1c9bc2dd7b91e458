"""BASELINE config 4: a synthetic driving sequence (default 1000 frames, 32-frame chunks, 8 overlap -> 42 chunks, the last one
16 frames long) through the chunk pipeline on 1 / 2 / 4 / 8 GPUs of one box:

    python tools/run_sequence.py                                                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_sequence.py

Chunks are dealt to the ranks by lsvs_b200.scheduler (run_sequence); the sequence is run twice and the second pass is timed on
the device (max over ranks).  Prints one JSON line: output frames/s (1000 / time) and frame-forwards/s (1328 / time).
NOT YET RUN ON A GPU: written after the round-1 GPU budget was spent; the host logic is what tests/test_scheduler.py covers."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "large-scale-vit-slam_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from aligned_vggt.models.featureAligned_vggt import FeatureAlignedVGGT  # noqa: E402
from lsvs_b200.scheduler import generate_chunks, model_pipeline, run_sequence  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--chunk", type=int, default=32)
ap.add_argument("--overlap", type=int, default=8)
ap.add_argument("--hw", type=int, nargs=2, default=[154, 518])
ap.add_argument("--head-cost", type=float, default=0.085)
ap.add_argument("--check", action="store_true", help="rank 0 re-runs the sequence alone and compares every chunk's aligned poses")
args = ap.parse_args()

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.set_grad_enabled(False)
torch.manual_seed(0)
with torch.device(dev):
    model = FeatureAlignedVGGT(enable_point=False, enable_depth=False, enable_track=False).eval()
H, W = args.hw
chunks = generate_chunks(args.frames, "chunk_overlap", args.chunk, args.overlap)
frames_of = [len(c) for c in chunks]
S_max = max(frames_of)
pts = torch.randn(1, S_max, H, W, 3, device=dev) * 10
dep = torch.rand(1, S_max, H, W, 1, device=dev) + 0.5


def load_chunk(k):
    """Synthetic frames of chunk k (seeded by the chunk index) + stand-ins for its DPT point / depth maps."""
    g = torch.Generator(device=dev).manual_seed(k)
    S = frames_of[k]
    return torch.rand(1, S, 3, H, W, device=dev, generator=g), pts[:, :S].contiguous(), dep[:, :S].contiguous()


def one_pass(solo=False):
    r, w = (0, 1) if solo else (rank, world)
    pipe = model_pipeline(model, args.overlap, S_max, H, W, r, w, dev, head_cost=args.head_cost, chunk_frames=frames_of)
    if solo:
        return 0.0, run_sequence(pipe, load_chunk)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = run_sequence(pipe, load_chunk)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        pipe.tx.close()
    return float(ms.item()), res


one_pass()                       # warm-up pass (allocations, kernel module loads, mailbox mapping)
ms, res = one_pass()
owned = [k for k, _ in res]
if world > 1:
    table = [None] * world
    dist.all_gather_object(table, owned)
    owned = sorted(k for t in table for k in t)
assert owned == list(range(len(chunks))), "every chunk must come back exactly once"
max_dev = None
if args.check and world > 1:
    # the sharded run must reproduce the sequential chunk loop: same poses for every chunk (split-K reduce-adds make the encoder
    # reproducible only to rounding, hence a tolerance instead of bit equality)
    mine = [(k, r["pose_enc"].cpu()) for k, r in res]
    table = [None] * world
    dist.all_gather_object(table, mine)
    if rank == 0:
        got = dict(kv for t in table for kv in t)
        _, solo = one_pass(solo=True)
        max_dev = max(float((got[k] - r["pose_enc"].cpu()).abs().max()) for k, r in solo)
        assert max_dev < 1e-3, f"sharded run deviates from the sequential loop: {max_dev}"
    dist.barrier()
if rank == 0:
    print(json.dumps({"max_pose_enc_deviation_vs_sequential": max_dev, "config": f"{args.frames} frames, {args.chunk}-frame chunks, {args.overlap} overlap, {H}x{W}", "n_gpus": world,
                      "chunks": len(chunks), "tail_chunk_frames": frames_of[-1], "seconds": ms / 1e3,
                      "output_frames_per_s": args.frames / (ms / 1e3), "frame_forwards_per_s": sum(frames_of) / (ms / 1e3)}), flush=True)
if world > 1:
    dist.destroy_process_group()

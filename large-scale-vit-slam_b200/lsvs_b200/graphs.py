"""One CUDA-graph replay per chunk.

Short chunks (the reference ships chunk 5 / overlap 1 for the feature-aligned model, training/config/
test_featureAlignedVGGT_vkitti.yaml:13,15) are launch-bound: ~800 kernel launches of a few microseconds each per chunk.
`GraphedChunk` captures one steady-state `model.forward(images, num_overlap, context)` — a chunk WITH context, fixed shapes —
into a CUDA graph whose inputs (images, stand-in maps) and inter-chunk state (processed overlap tokens, memory tokens, aligned
poses of the previous chunk) live in static buffers; a call copies the new frames in, replays the graph and the graph's last
nodes move the new state into the static context, so consecutive calls chain exactly like the reference's chunk loop
(training/run_model.py:326-338).

What makes the forward capturable: the native engine allocates its workspace and uploads its position-id tables only when a
shape is seen for the first time (two eager warm-up passes precede the capture), tensor maps are encoded on the host, and every
kernel — including the programmatic-dependent-launch ones — goes to torch's current stream.
"""
from typing import Dict, Optional

import torch


class GraphedChunk:
    def __init__(self, model, num_overlap: int, images: torch.Tensor, context: Dict, raw_points: Optional[torch.Tensor] = None,
                 raw_depth: Optional[torch.Tensor] = None, warmup: int = 2):
        """images (B,S,3,H,W): shape (and device) of every chunk to come; context: the reference-style context of the PREVIOUS chunk
        (predictions of a forward, or what a GraphedChunk's context() returns) — at least overlap_tokens, memory_tokens, pose_enc."""
        if not images.is_cuda:
            raise ValueError("GraphedChunk needs CUDA tensors")
        self.model, self.ov = model, int(num_overlap)
        self.S = images.shape[1]
        self.img = images.clone()
        self.pts = None if raw_points is None else raw_points.clone()
        self.dep = None if raw_depth is None else raw_depth.clone()
        self.ctx_ov = context["overlap_tokens"].detach().clone()
        self.ctx_pose = context["pose_enc"][-1].detach().clone()
        mem = context.get("memory_tokens")
        self.ctx_mem = None if not mem else mem[-1].detach().clone()
        B = images.shape[0]
        self._dummy_sim3 = torch.zeros(B, 1, 8, device=images.device)
        self._dummy_se3 = torch.zeros(B, max(self.S - 1, 1), 7, device=images.device)
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):   # every allocation / table upload / kernel attribute of this shape happens here
                self._forward(advance=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._forward(advance=True)

    def context(self) -> Dict:
        """Fresh reference-style context dict over the static state (what model.forward expects as `context`)."""
        ctx = {"overlap_tokens": self.ctx_ov, "pose_enc": [self.ctx_pose], "chunk_sim3_alignment_enc": self._dummy_sim3,
               "frame_se3_alignment_enc": self._dummy_se3}
        if self.ctx_mem is not None:
            ctx["memory_tokens"] = [self.ctx_mem]
        return ctx

    def _forward(self, advance: bool) -> Dict:
        pred = self.model(self.img, self.ov, self.context(), raw_depth=self.dep, raw_points=self.pts)
        out = {"pose_enc": pred["pose_enc"][-1], "chunk_sim3_alignment_enc": pred["chunk_sim3_alignment_enc"][:, -1:],
               "frame_se3_alignment_enc": pred["frame_se3_alignment_enc"][:, -(self.S - 1):] if self.S > 1 else pred["frame_se3_alignment_enc"][:, :0],
               "overlap_tokens": pred["overlap_tokens"]}
        if self.ctx_mem is not None:
            out["memory_tokens"] = pred["memory_tokens"][-1]
        for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
            if k in pred:
                out[k] = pred[k][-1]
        if advance:  # last nodes of the graph: this chunk's state becomes the next chunk's context
            self.ctx_ov.copy_(out["overlap_tokens"])
            self.ctx_pose.copy_(out["pose_enc"])
            if self.ctx_mem is not None:
                self.ctx_mem.copy_(out["memory_tokens"])
        return out

    def __call__(self, images: torch.Tensor, raw_points: Optional[torch.Tensor] = None, raw_depth: Optional[torch.Tensor] = None) -> Dict:
        """Run the next chunk.  Returns this chunk's outputs (static tensors: valid until the next call)."""
        if images.shape != self.img.shape:
            raise ValueError(f"GraphedChunk was captured for chunks of shape {tuple(self.img.shape)}, got {tuple(images.shape)}")
        self.img.copy_(images)
        if raw_points is not None:
            self.pts.copy_(raw_points)
        if raw_depth is not None:
            self.dep.copy_(raw_depth)
        self.graph.replay()
        return self.out

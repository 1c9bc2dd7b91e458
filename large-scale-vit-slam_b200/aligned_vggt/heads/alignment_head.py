"""Drop-in for the reference's aligned_vggt/heads/alignment_head.py: same class name, constructor arguments,
forward signature and state_dict keys (:52-221, :224-345).  Inference (eval(), torch.no_grad() or a frozen head) runs the fused
native engine (csrc/engine.cu: lsvs_alignment_head_forward); in train() mode with gradients enabled the forward builds an autograd
graph whose Linears and attention are this library's kernels (lsvs_b200/train.py) — the reference trains exactly this module
(alignment_head.py:361,385,498,527) with everything else frozen."""
from typing import Tuple

import torch

from lsvs_b200 import specs
from lsvs_b200.engine import Engine
from lsvs_b200.modules import _EngineBound


class AlignmentHead(_EngineBound):
    _prefix = "alignment_head."

    def __init__(self, patch_size=14, in_dim=2048, embed_dim=1024, dec_dim=512, depth_aa=4, depth_decoder=2, num_heads=8,
                 mlp_ratio=4.0, num_register_tokens=4, qkv_bias=True, proj_bias=True, ffn_bias=True,
                 aa_order=["frame", "temporal"], aa_block_size=1, qk_norm=True, rope_freq=100, init_values=0.01,
                 num_memory_tokens=8, temporal_attention=True):
        if depth_aa % aa_block_size != 0:
            raise ValueError(f"depth ({depth_aa}) must be divisible by aa_block_size ({aa_block_size})")
        if not temporal_attention:
            # the reference's own temporal_attention=False branch cannot run (alignment_head.py:80,146: self.aa_order
            # is stored before the local aa_order is rebound, forward then needs the never-built temporal_blocks)
            raise AttributeError("'AlignmentHead' object has no attribute 'temporal_blocks' (temporal_attention=False "
                                 "is not runnable in the reference either)")
        if not (patch_size == 14 and in_dim == 2048 and embed_dim == 1024 and dec_dim == 512 and depth_decoder == 2
                and num_heads == 8 and mlp_ratio == 4.0 and num_register_tokens == 4 and aa_block_size == 1 and qk_norm
                and rope_freq > 0 and num_memory_tokens in (0, 8) and list(aa_order) == ["frame", "temporal"]):
            raise ValueError("only the reference's AlignmentHead geometry is built on this path (dim 1024, 8 heads, decoder dim 512, "
                             "num_memory_tokens 8 or 0)")
        super().__init__(specs.alignment_head_spec(in_dim, embed_dim, dec_dim, depth_aa, depth_decoder, num_heads, num_memory_tokens))
        self.num_memory_tokens, self.temporal_attention = num_memory_tokens, temporal_attention
        self.depth_aa, self.patch_size, self.rope_freq = depth_aa, patch_size, float(rope_freq)
        self.patch_start_idx = 1 + 1 + num_register_tokens
        specs.init_default_(self)

    def _make_engine(self):
        return Engine(0, 0, self.depth_aa, self.num_memory_tokens, True, False, self.rope_freq)

    def forward(self, tokens: torch.Tensor, image_size: Tuple[int, int], next_num_overlap: int,
                overlap_tokens: torch.Tensor = None, memory_tokens: torch.Tensor = None):
        """tokens (B,S,P,2048) -> chunk_sim3 (B,1,8), frame_se3 (B,S-1,7), memory (B,8,512),
        overlap tokens (B,1+next_num_overlap,P+1,1024) contiguous."""
        from lsvs_b200 import train as _train
        if _train.wants_training_path(self):
            owner = getattr(self, "_owner", None)
            precision = getattr(owner, "precision", None) or 0
            return _train.alignment_head_forward_train(self, tokens, image_size, next_num_overlap, overlap_tokens, memory_tokens,
                                                       precision=1 if precision else 0)
        return self._engine().alignment_head_forward(tokens, image_size, next_num_overlap, overlap_tokens, memory_tokens)

    def forward_prefix(self, tokens: torch.Tensor, image_size: Tuple[int, int]) -> torch.Tensor:
        """The part of forward() that needs no context of the previous chunk (project_in, token_norm, alignment token, first frame
        block): (B,S,P,2048) -> fp32 token stream (B,S,P+1,1024).  The chunk scheduler runs it on the rank that encoded
        the chunk and ships the stream instead of the tapped tokens."""
        return self._engine().alignment_head_prefix(tokens, image_size)

    def forward_from_prefix(self, prefix: torch.Tensor, image_size: Tuple[int, int], next_num_overlap: int,
                            overlap_tokens: torch.Tensor = None, memory_tokens: torch.Tensor = None):
        """forward() continued from forward_prefix()'s stream: identical outputs, bit for bit."""
        return self._engine().alignment_head_resume(prefix, image_size, next_num_overlap, overlap_tokens, memory_tokens)

    def _decode_alignments(self, frame_alignment_tokens: torch.Tensor, num_overlap: int, is_first_chunk: bool,
                           memory_tokens: torch.Tensor = None):
        """reference :427-540 (eval path): (B,S,1024) -> chunk_sim3 (B,1,8), frame_se3 (B,S-1,7), memory (B,8,512).
        As in the reference, a first chunk is recognised by memory_tokens being None."""
        return self._engine().alignment_decode_forward(frame_alignment_tokens, memory_tokens)

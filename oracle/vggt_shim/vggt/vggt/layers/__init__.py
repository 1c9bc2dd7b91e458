"""Clean-room stand-in for vggt.vggt.layers (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from .mlp import Mlp
from .layer_scale import LayerScale
from .attention import Attention, MemEffAttention
from .block import Block, NestedTensorBlock
from .patch_embed import PatchEmbed

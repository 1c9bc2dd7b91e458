import torch.nn as nn
from oracle import functional as OF


class PositionGetter:
    def __init__(self):
        self.position_cache = {}

    def __call__(self, batch_size, height, width, device):
        return OF.grid_positions(batch_size, height, width, device)


class RotaryPositionEmbedding2D(nn.Module):
    def __init__(self, frequency=100.0, scaling_factor=1.0):
        super().__init__()
        self.base_frequency = frequency
        self.scaling_factor = scaling_factor

    def forward(self, tokens, positions):
        assert tokens.size(-1) % 2 == 0
        assert positions.ndim == 3 and positions.shape[-1] == 2
        return OF.rope_apply_2d(tokens, positions, self.base_frequency)

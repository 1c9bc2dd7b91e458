"""TrackHead is out of scope: import-time stand-in only."""
import torch.nn as nn


class TrackHead(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("TrackHead is outside the oracle's scope; construct the model with enable_track=False")

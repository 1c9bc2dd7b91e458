"""DPTHead is out of scope: import-time stand-in only."""
import torch.nn as nn


class DPTHead(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("DPTHead is outside the oracle's scope; construct the model with enable_depth/point=False")

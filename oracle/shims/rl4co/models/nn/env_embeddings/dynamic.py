import torch.nn as nn


class StaticEmbedding(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, td):
        return 0, 0, 0

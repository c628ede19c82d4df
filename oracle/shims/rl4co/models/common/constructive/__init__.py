import torch.nn as nn


class AutoregressiveEncoder(nn.Module):
    pass


class AutoregressiveDecoder(nn.Module):
    pass

import torch
from torch.distributions import Uniform


def batch_to_scalar(param):
    if len(param.shape) > 0:
        return param.flatten()[0].item() if hasattr(param, "flatten") else param[0]
    return param.item() if isinstance(param, torch.Tensor) else param


def get_sampler(val_name, distribution, low=0, high=1.0, **kwargs):
    if distribution in (Uniform, "uniform"):
        return Uniform(low=low, high=high)
    raise NotImplementedError(f"shim get_sampler: {distribution}")


class Generator:
    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def __call__(self, batch_size):
        batch_size = [batch_size] if isinstance(batch_size, int) else batch_size
        return self._generate(batch_size)

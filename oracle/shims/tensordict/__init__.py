"""Stand-in for tensordict.TensorDict: the subset of behaviour the reference hot path touches."""
from tensordict.tensordict import TensorDict  # noqa: F401

import numpy as np
import torch
from tensordict import TensorDict


def load_npz_to_tensordict(filename):
    x = np.load(filename)
    x_dict = {k: torch.from_numpy(v) for k, v in dict(x).items()}
    batch_size = x_dict[list(x_dict.keys())[0]].shape[0]
    return TensorDict(x_dict, batch_size=batch_size)


def save_tensordict_to_npz(td, path, compress=False):
    np.savez(path, **{k: v.numpy() for k, v in td.items()})

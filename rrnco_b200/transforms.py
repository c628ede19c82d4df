"""x8 state augmentation (rrnco/models/utils/transforms.py:15-154): `batchify(td, num_augment)` plus a transform of the
coordinate features (`locs`) only -- the distance / duration matrices are identical across the augmented copies; only
the encoder sees `locs`.

`share_instance_data=True` is the un-materialised form for the fused rollout path: the transformed features get the
`num_augment * B` rows upstream produces, every other entry stays at its `B` rows (no 8-fold copy of the `[N, N]`
matrices and demand rows: 334 MB per 1024-instance batch at n = 100), and the td's batch size becomes `num_augment * B`.
The rrnco_b200 envs and `RRNetPolicy.forward` read instance data as row `r % data_rows`, and the augmented layout is
`a * B + b` (repeat-major batchify), so copy `a` of instance `b` reads row `b`.  Upstream's own per-step loop needs the
materialised form (the default).
"""
from __future__ import annotations

import math
from typing import Callable, Union

import torch

from .tdlite import batchify, make_td_like


def dihedral_8_augmentation(xy: torch.Tensor) -> torch.Tensor:
    """transforms.py:15-37: the 8 symmetries of the unit square, concatenated along the batch ([8 B, n, 2])."""
    x, y = xy.split(1, dim=2)
    z = [torch.cat(p, dim=2) for p in ((x, y), (1 - x, y), (x, 1 - y), (1 - x, 1 - y), (y, x), (1 - y, x), (y, 1 - x),
                                       (1 - y, 1 - x))]
    return torch.cat(z, dim=0)


def dihedral_8_augmentation_wrapper(xy: torch.Tensor, reduce: bool = True, *args, **kw) -> torch.Tensor:
    """transforms.py:40-47: `xy` is the batchified feature; the first eighth is the original data."""
    xy = xy[: xy.shape[0] // 8, ...] if reduce else xy
    return dihedral_8_augmentation(xy)


def symmetric_transform(x, y, phi, offset: float = 0.5):
    """transforms.py:50-70."""
    x, y = x - offset, y - offset
    x_prime = torch.cos(phi) * x - torch.sin(phi) * y
    y_prime = torch.sin(phi) * x + torch.cos(phi) * y
    mask = phi > 2 * math.pi
    xy = torch.cat((x_prime, y_prime), dim=-1)
    xy = torch.where(mask, xy.flip(-1), xy)
    return xy + offset


def symmetric_augmentation(xy: torch.Tensor, num_augment: int = 8, first_augment: bool = False):
    """transforms.py:73-88: random rotation / reflection per augmented copy (the first copy untouched)."""
    phi = torch.rand(xy.shape[0], device=xy.device) * 4 * math.pi
    if not first_augment:
        phi[: xy.shape[0] // num_augment] = 0.0
    x, y = xy[..., [0]], xy[..., [1]]
    return symmetric_transform(x, y, phi[:, None, None])


def min_max_normalize(x):
    return (x - x.min()) / (x.max() - x.min())


def get_augment_function(augment_fn: Union[str, Callable]):
    if callable(augment_fn):
        return augment_fn
    if augment_fn == "dihedral8":
        return dihedral_8_augmentation_wrapper
    if augment_fn == "symmetric":
        return symmetric_augmentation
    raise ValueError(f"Unknown augment_fn: {augment_fn}. Available options: 'symmetric', 'dihedral8' or a custom callable")


class StateAugmentation(object):
    """transforms.py:106-154, same constructor; plus `share_instance_data` (see the module docstring)."""

    def __init__(self, num_augment: int = 8, augment_fn: Union[str, Callable] = "symmetric", first_aug_identity: bool = True,
                 normalize: bool = False, feats: list = None, no_aug_coords: bool = True, share_instance_data: bool = False):
        self.augmentation = get_augment_function(augment_fn)
        assert not (self.augmentation == dihedral_8_augmentation_wrapper and num_augment != 8), \
            "When using the `dihedral8` augmentation function, then num_augment must be 8"
        if no_aug_coords:
            self.feats = []
        elif feats is None:
            self.feats = ["locs"]
        else:
            self.feats = feats
        self.num_augment = num_augment
        self.normalize = normalize
        self.first_aug_identity = first_aug_identity
        self.share_instance_data = share_instance_data

    def __call__(self, td):
        B = td.batch_size[0]
        if self.share_instance_data:
            # coordinate features (depot, locs: what reset concatenates and the encoder reads) get num_augment * B rows
            coords = set(self.feats) | ({"depot"} if "depot" in td.keys() else set()) | ({"locs"} if "locs" in td.keys() else set())
            data = {k: (batchify(td[k], self.num_augment) if k in coords else td[k]) for k in td.keys()}
            td_aug = make_td_like(td, data, [B * self.num_augment, *td.batch_size[1:]])
        else:
            td_aug = batchify(td, self.num_augment)
        for feat in self.feats:
            if not self.first_aug_identity:
                init_aug_feat = td_aug[feat][:B].clone()
            aug_feat = self.augmentation(td_aug[feat], self.num_augment)
            if self.normalize:
                aug_feat = min_max_normalize(aug_feat)
            if not self.first_aug_identity:
                aug_feat[:B] = init_aug_feat
            td_aug[feat] = aug_feat
        return td_aug

"""Dict-of-tensors stand-in for tensordict.TensorDict + the rl4co.utils.ops helpers.

Semantics restated from rl4co 0.6.0; each is corroborated by
how the reference uses it:
  batchify     -> rrnco/models/decoding.py:189, rrnco/models/utils/transforms.py:143
  unbatchify   -> rrnco/models/decoder.py:173, rrnco/models/rl.py:112, test.py:211
  gather_by_index -> rrnco/models/decoder.py:187-193, rrnco/envs/rcvrp/env.py:95-97
"""
from __future__ import annotations

import torch


class TD(dict):
    """Minimal TensorDict: a dict of tensors sharing leading batch dims."""

    def __init__(self, data=None, batch_size=None, device=None):
        super().__init__(data or {})
        if batch_size is None:
            batch_size = []
        if isinstance(batch_size, int):
            batch_size = [batch_size]
        self.batch_size = torch.Size(batch_size)
        self._device = device

    # -- tensordict-like surface ------------------------------------------
    @property
    def device(self):
        if self._device is not None:
            return self._device
        for v in self.values():
            if isinstance(v, torch.Tensor):
                return v.device
        return torch.device("cpu")

    @property
    def shape(self):
        return self.batch_size

    def dim(self):
        return len(self.batch_size)

    def size(self, d=None):
        return self.batch_size if d is None else self.batch_size[d]

    def set(self, key, value):
        self[key] = value
        return self

    def clone(self):
        return TD({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self.items()},
                  batch_size=self.batch_size, device=self._device)

    def to(self, device):
        return TD({k: v.to(device) for k, v in self.items()}, batch_size=self.batch_size, device=device)

    def apply(self, fn, batch_size=None):
        return TD({k: fn(v) for k, v in self.items()},
                  batch_size=self.batch_size if batch_size is None else batch_size, device=self._device)

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        # batch indexing
        out = {k: v[key] for k, v in self.items()}
        probe = torch.empty(self.batch_size, device="meta")[key]
        return TD(out, batch_size=probe.shape, device=self._device)


# ---------------------------------------------------------------------------
# rl4co.utils.ops
# ---------------------------------------------------------------------------

def _batchify_single(x, repeats: int):
    if isinstance(x, TD):
        bs = [x.batch_size[0] * repeats, *x.batch_size[1:]]
        return x.apply(lambda t: _batchify_single(t, repeats), batch_size=bs)
    s = x.shape
    return x.expand(repeats, *s).contiguous().view(s[0] * repeats, *s[1:])


def batchify(x, shape):
    """Repeat-major replicate: flat index = rep * B + b (decoding.py:189)."""
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _batchify_single(x, s) if s > 0 else x
    return x


def _unbatchify_single(x, repeats: int):
    if isinstance(x, TD):
        bs = [x.batch_size[0] // repeats, repeats, *x.batch_size[1:]]
        return x.apply(lambda t: _unbatchify_single(t, repeats), batch_size=bs)
    s = x.shape
    return x.view(repeats, s[0] // repeats, *s[1:]).permute(1, 0, *range(2, len(s) + 1))


def unbatchify(x, shape):
    """Inverse of batchify: [rep*B, ...] -> [B, rep, ...] (a view; decoder.py:173)."""
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _unbatchify_single(x, s) if s > 0 else x
    return x


def gather_by_index(src, idx, dim=1, squeeze=True):
    """idx right-padded with singleton dims, expanded to src with -1 at `dim`, gather,
    squeeze `dim` iff it has size 1 (shapes at decoder.py:187-193, rcvrp/env.py:95-97)."""
    expanded_shape = list(src.shape)
    expanded_shape[dim] = -1
    idx = idx.view(idx.shape + (1,) * (src.dim() - idx.dim())).expand(expanded_shape)
    squeeze = idx.size(dim) == 1 and squeeze
    return src.gather(dim, idx).squeeze(dim) if squeeze else src.gather(dim, idx)

"""Minimal TensorDict stand-in + the rl4co ops the hot path needs.

`tensordict` / `rl4co` are not installed in this image.  When they are (the reference's own environment),
the env / policy classes accept real TensorDicts unchanged: everything here is duck-typed on
`td[key]`, `td.set`, `td.update`, `td.get`, `td.batch_size`, `key in td`.
"""
from __future__ import annotations

import torch


class TensorDictLite(dict):
    def __init__(self, source=None, batch_size=None, device=None):
        super().__init__(source or {})
        if batch_size is None:
            batch_size = []
        if isinstance(batch_size, int):
            batch_size = [batch_size]
        self.batch_size = torch.Size(batch_size)
        self._device = torch.device(device) if device is not None else None

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

    def _map(self, fn, batch_size=None, device=None):
        return TensorDictLite({k: fn(v) for k, v in self.items()},
                              batch_size=self.batch_size if batch_size is None else batch_size,
                              device=self._device if device is None else device)

    def clone(self):
        return self._map(lambda v: v.clone())

    def to(self, device, non_blocking=False):
        return self._map(lambda v: v.to(device, non_blocking=non_blocking), device=device)

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        probe = torch.empty(self.batch_size, device="meta")[key]
        return self._map(lambda v: v[key], batch_size=probe.shape)


def make_td_like(td, data, batch_size):
    """New container of the same kind as `td` (real TensorDict when the caller uses one)."""
    cls = type(td)
    if cls is dict or cls is TensorDictLite or not hasattr(td, "batch_size"):
        return TensorDictLite(data, batch_size=batch_size)
    return cls(data, batch_size=batch_size)


def _batchify_single(x, repeats):
    if isinstance(x, TensorDictLite):
        return x._map(lambda t: _batchify_single(t, repeats), [x.batch_size[0] * repeats, *x.batch_size[1:]])
    s = x.shape
    return x.expand(repeats, *s).contiguous().view(s[0] * repeats, *s[1:])


def batchify(x, shape):
    """Repeat-major replication, flat index rep * B + b (rl4co.utils.ops.batchify)."""
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _batchify_single(x, s) if s > 0 else x
    return x


def _unbatchify_single(x, repeats):
    if isinstance(x, TensorDictLite):
        return x._map(lambda t: _unbatchify_single(t, repeats),
                      [x.batch_size[0] // repeats, repeats, *x.batch_size[1:]])
    s = x.shape
    return x.view(repeats, s[0] // repeats, *s[1:]).permute(1, 0, *range(2, len(s) + 1))


def unbatchify(x, shape):
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _unbatchify_single(x, s) if s > 0 else x
    return x

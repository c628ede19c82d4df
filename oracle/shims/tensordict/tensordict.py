import torch


class TensorDict:
    def __init__(self, source=None, batch_size=None, device=None):
        self._d = dict(source or {})
        if batch_size is None:
            batch_size = []
        if isinstance(batch_size, int):
            batch_size = [batch_size]
        self.batch_size = torch.Size(batch_size)
        self._device = torch.device(device) if device is not None else None

    # mapping surface -------------------------------------------------------
    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def values(self):
        return self._d.values()

    def __contains__(self, k):
        return k in self._d

    def __len__(self):
        return self.batch_size[0] if len(self.batch_size) else 0

    def get(self, k, default=None):
        return self._d.get(k, default)

    def set(self, k, v):
        self._d[k] = v
        return self

    def update(self, other, **kw):
        self._d.update(dict(other.items()) if isinstance(other, TensorDict) else other)
        self._d.update(kw)
        return self

    def __setitem__(self, k, v):
        self._d[k] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return self._d[k]
        nb = len(self.batch_size)
        probe = torch.empty(self.batch_size, device="meta")[k]
        return TensorDict({n: v[k] for n, v in self._d.items()}, batch_size=probe.shape, device=self._device)

    # tensor-like surface over the batch dims -------------------------------
    @property
    def device(self):
        if self._device is not None:
            return self._device
        for v in self._d.values():
            return v.device
        return torch.device("cpu")

    @property
    def shape(self):
        return self.batch_size

    def size(self, d=None):
        return self.batch_size if d is None else self.batch_size[d]

    def dim(self):
        return len(self.batch_size)

    def _map(self, fn, batch_size):
        return TensorDict({k: fn(v) for k, v in self._d.items()}, batch_size=batch_size, device=self._device)

    def to(self, device):
        out = self._map(lambda v: v.to(device), self.batch_size)
        out._device = torch.device(device) if not isinstance(device, torch.device) else device
        return out

    def clone(self):
        return self._map(lambda v: v.clone(), self.batch_size)

    def contiguous(self):
        return self._map(lambda v: v.contiguous(), self.batch_size)

    def expand(self, *shape):
        nb = len(self.batch_size)
        return self._map(lambda v: v.expand(*shape, *v.shape[nb:]), shape)

    def view(self, *shape):
        nb = len(self.batch_size)
        return self._map(lambda v: v.view(*shape, *v.shape[nb:]), shape)

    def permute(self, *dims):
        nb = len(self.batch_size)
        new_bs = [self.batch_size[d] for d in dims]
        return self._map(lambda v: v.permute(*dims, *range(nb, v.dim())), new_bs)

"""Pinned-host -> HBM staging of instance batches on a dedicated copy stream.

The reference moves every batch with `batch.to(device)` on the compute stream (`rrnco/models/rl.py:99-106`,
`test.py:189-200`), so the copy and the rollout serialise.  Here the batch of step k+1 is copied by the copy engine
into a recycled device buffer set while the rollout of step k runs: per step the host->device bytes are the same,
only their place on the timeline changes.  Two events per slot order the streams: `ready` (copy done -> consumer may
read) and `free` (consumer's last kernel enqueued -> the copy stream may overwrite).
"""
from __future__ import annotations

import torch


class HostPrefetcher:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPrefetcher stages into HBM: a CUDA device is required")
        self.depth = depth
        self.stream = torch.cuda.Stream(self.device)
        self._bufs = [None] * depth
        self._ready = [None] * depth
        self._free = [None] * depth
        self._n = 0

    def submit(self, host: dict) -> int:
        """Enqueue the H2D copies of one batch ({name: pinned host tensor}); returns a ticket for `acquire`."""
        slot = self._n % self.depth
        self._n += 1
        bufs = self._bufs[slot]
        if bufs is None or any(k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype
                               for k, v in host.items()) or len(bufs) != len(host):
            bufs = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            self._bufs[slot] = bufs
            # fresh memory may still be in use by kernels enqueued on the consumer's stream
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self._free[slot] is not None:
            self.stream.wait_event(self._free[slot])
        with torch.cuda.stream(self.stream):
            for k, v in host.items():
                if not v.is_pinned():
                    raise RuntimeError(f"HostPrefetcher: '{k}' is not in pinned host memory")
                bufs[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._ready[slot] = ev
        return slot

    def acquire(self, ticket: int) -> dict:
        """Device tensors of a submitted batch; the current stream waits for its copies."""
        torch.cuda.current_stream(self.device).wait_event(self._ready[ticket])
        return self._bufs[ticket]

    def release(self, ticket: int) -> None:
        """Call after the last kernel reading the batch has been enqueued on the current stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[ticket] = ev

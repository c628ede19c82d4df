"""On-disk formats either side of the rollout path, read straight into pinned host memory.

  * city files  `<city>_data.npz` = {distance [L,L] f64 (km), duration [L,L] f64 (min), points [L,2] f64}
    written by data_generation/utilities/create_dataset.py:169-174, read by rcvrp/generator_lazy.py:134-140;
  * test sets   `*.npz` written by scripts/generate_data.py:201-224,354-372,457 and read by test.py:152 through rl4co's
    `load_npz_to_tensordict` [rl4co-recalled: every array becomes a tensor, batch size = leading dimension]:
      atsp    {locs, distance_matrix}
      rcvrp   {depot, locs, demand, capacity, distance_matrix}            (demand in units, capacity per instance)
      rcvrptw {depot/locs, demand_linehaul, time_windows, service_time, vehicle_capacity, speed, distance_matrix,
               duration_matrix}
    `prepare_test_td` applies test.py:160-177's pre-processing (demand / capacity, capacity := 1).

Host-side only: nothing here touches the GPU except the optional `pin=True` page-locking, which needs a CUDA runtime.
Batches then go through `HostPrefetcher` (hostio.py) into HBM.
"""
from __future__ import annotations

import numpy as np
import torch

from .tdlite import TensorDictLite

CITY_KEYS = ("distance", "duration", "points")
TEST_KEYS = {
    "atsp": ("locs", "distance_matrix"),
    "rcvrp": ("depot", "locs", "demand", "capacity", "distance_matrix"),
    "rcvrptw": ("locs", "demand_linehaul", "time_windows", "service_time", "distance_matrix", "duration_matrix"),
}


def load_city_npz(path: str) -> dict:
    """City arrays as float64 NumPy (the dict `Real_World_Sampler.sample` / `CityOnDevice` take)."""
    with np.load(path, allow_pickle=True) as z:
        missing = [k for k in ("distance", "points") if k not in z.files]
        if missing:
            raise KeyError(f"{path}: city file lacks {missing} (has {z.files})")
        out = {k: np.ascontiguousarray(z[k], dtype=np.float64) for k in CITY_KEYS if k in z.files}
    L = out["points"].shape[0]
    for k in ("distance", "duration"):
        if k in out and out[k].shape != (L, L):
            raise ValueError(f"{path}: {k} has shape {out[k].shape}, expected {(L, L)}")
    return out


def load_npz_to_tensordict(path: str, pin: bool = False) -> TensorDictLite:
    """rl4co.data.utils.load_npz_to_tensordict (test.py:152): host TensorDictLite, batch size = leading dimension."""
    with np.load(path) as z:
        if not z.files:
            raise ValueError(f"{path}: empty archive")
        arrays = {k: np.ascontiguousarray(z[k]) for k in z.files}
    batch = next(iter(arrays.values())).shape[0]
    tensors = {}
    for k, v in arrays.items():
        if v.ndim == 0 or v.shape[0] != batch:
            raise ValueError(f"{path}: '{k}' has leading dimension {v.shape[:1]}, expected {batch}")
        t = torch.from_numpy(v)
        tensors[k] = t.pin_memory() if pin else t
    return TensorDictLite(tensors, batch_size=[batch])


def prepare_test_td(td: TensorDictLite, problem: str) -> TensorDictLite:
    """test.py:154-177: per-problem pre-processing of a loaded test set (demand normalised by capacity)."""
    if problem not in TEST_KEYS:
        raise ValueError(f"Problem {problem} not supported")
    missing = [k for k in TEST_KEYS[problem] if k not in td.keys()]
    if missing:
        raise KeyError(f"{problem} test set lacks {missing}")
    if problem == "rcvrp":
        td.set("demand", td["demand"] / td["capacity"].unsqueeze(-1))
        td.set("capacity", torch.ones_like(td["capacity"]))
    return td


def iter_batches(td: TensorDictLite, batch_size: int):
    """test.py:61-71 `get_dataloader` (shuffle=False): consecutive slices; the last one may be short."""
    n = td.batch_size[0]
    for lo in range(0, n, batch_size):
        hi = min(n, lo + batch_size)
        yield TensorDictLite({k: td[k][lo:hi] for k in td.keys()}, batch_size=[hi - lo])

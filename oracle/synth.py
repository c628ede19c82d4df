"""Synthetic city-like inputs (TEST INFRASTRUCTURE; the spec is SURVEY.md section 8(d)).

The real per-city `.npz` files are HuggingFace downloads and absent offline
(rrnco/envs/rcvrp/generator.py:154-157 raises), so the benchmarks and parity tests use
synthetic 1000-node asymmetric city matrices with the on-disk schema of
data_generation/utilities/create_dataset.py:169-174 (float64 `distance` km, `duration` min,
`points`), sub-sampled exactly like the reference generators do.
"""
from __future__ import annotations

import numpy as np
import torch

from . import sampler
from .td import TD


def make_city(city_id: int = 0, length: int = 1000) -> dict:
    rng = np.random.default_rng(1000 + city_id)
    pts = rng.uniform(0.0, 3.0, size=(length, 2))  # 3 km x 3 km
    eu = np.linalg.norm(pts[:, None, :] - pts[None, :, :], axis=-1)
    dist = eu * (1.2 + 0.4 * rng.uniform(size=(length, length)))  # independent (i,j)/(j,i): asymmetric
    np.fill_diagonal(dist, 0.0)
    speed = rng.uniform(20.0, 50.0, size=(length, length))  # km/h
    dur = dist / speed * 60.0  # minutes
    return {"points": pts, "distance": dist, "duration": dur}


def _minmax(x, eps):
    lo = x.amin(dim=1, keepdim=True)
    hi = x.amax(dim=1, keepdim=True)
    return (x - lo) / (hi - lo + eps)


def make_instances(problem: str, batch: int, num_loc: int = 100, seed: int = 1234, city=None,
                   integer_demand: bool = True, indices=None) -> TD:
    """Instance batch in the format the reference generators hand to `env.reset`.

    rcvrp  : rrnco/envs/rcvrp/generator_lazy.py:275-304 (depot = first sampled point, demand / capacity 50)
    rcvrptw: rrnco/envs/rmtvrp/generator_lazy.py:350-370 + generator.py:515-562 (duration min-max
             normalised, Liu-style time windows from the duration matrix)
    atsp   : rrnco/envs/atsp/generator_lazy.py:239-260
    """
    city = make_city(0) if city is None else city
    rng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    n = num_loc if problem == "atsp" else num_loc + 1
    s = sampler.sample(city, batch, n, with_duration=(problem == "rcvrptw"), indices=indices, rng=rng)
    dm = torch.from_numpy(s["distance_matrix"].astype(np.float32))
    pts = torch.from_numpy(s["points"].astype(np.float32))
    capacity = 50.0  # CAPACITIES[100], rcvrp/generator.py:31
    if integer_demand:
        demand = torch.randint(1, 10, (batch, num_loc), generator=g).float() / capacity
    else:  # lazy generator law: continuous U(1,10) (rcvrp/generator_lazy.py:98-100)
        demand = (1 + 9 * torch.rand(batch, num_loc, generator=g)) / capacity
    if problem == "atsp":
        return TD({"locs": _minmax(pts, 1e-6), "distance_matrix": dm}, batch_size=[batch])
    if problem == "rcvrp":
        return TD({"locs": pts[:, 1:], "depot": pts[:, :1], "demand": demand, "distance_matrix": dm,
                   "capacity": torch.full((batch, 1), capacity)}, batch_size=[batch])
    if problem == "rcvrptw":
        dur = torch.from_numpy(s["duration_matrix"].astype(np.float32))
        lo = dur.amin(dim=(1, 2), keepdim=True)
        hi = dur.amax(dim=(1, 2), keepdim=True)
        rngd = torch.where(hi - lo == 0, torch.ones_like(hi), hi - lo)
        dur = (dur - lo) / rngd
        max_time = 4.6
        service = 0.15 + 0.03 * torch.rand(batch, num_loc, generator=g)
        tw_len = 0.18 + 0.02 * torch.rand(batch, num_loc, generator=g)
        d_0i, d_i0 = dur[:, 0, 1:], dur[:, 1:, 0]
        d_max = torch.max(d_0i, d_i0)
        h_max = (max_time - service - tw_len) / (d_max + 1e-6) - 1
        tw_start = d_0i + (h_max - 1) * d_max * torch.rand(batch, num_loc, generator=g)
        tw_end = tw_start + tw_len
        tw = torch.stack((torch.cat((torch.zeros(batch, 1), tw_start), -1),
                          torch.cat((torch.full((batch, 1), max_time), tw_end), -1)), dim=-1)
        service = torch.cat((torch.zeros(batch, 1), service), dim=-1)
        return TD({"locs": _minmax(pts, 1e-8), "demand_linehaul": demand, "distance_matrix": dm,
                   "duration_matrix": dur, "time_windows": tw, "service_time": service,
                   "vehicle_capacity": torch.ones(batch, 1), "capacity_original": torch.full((batch, 1), capacity),
                   "speed": torch.ones(batch, 1)}, batch_size=[batch])
    raise ValueError(problem)


def random_embeddings(batch: int, n: int, embed_dim: int = 128, seed: int = 0):
    """Stand-in encoder output with the statistics of an instance-normalised encoder (unit variance)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, n, embed_dim, generator=g), torch.randn(batch, n, embed_dim, generator=g)

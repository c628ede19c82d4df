"""CPU oracle of the instance sub-sampler (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows rrnco/envs/rcvrp/sampler.py:8-104 (ATSP twin identical; the RMTVRP twin
rrnco/envs/rmtvrp/sampler.py:80 additionally gathers `duration`).  Only the
`uniform` location law is restated (cluster laws: out of scope).
"""
from __future__ import annotations

import numpy as np


def uniform_sample(batch: int, data_length: int, num_sample: int, rng=np.random) -> np.ndarray:
    # sampler.py:97-104 -- one np.random.choice(replace=False) per instance, in order
    return np.array([rng.choice(data_length, num_sample, replace=False) for _ in range(batch)])


def gather_submatrix(mat: np.ndarray, indices: np.ndarray) -> np.ndarray:
    # sampler.py:84-90 -- out[b,i,j] = M[idx[b,i], idx[b,j]], dtype of M (float64 city data)
    return mat[indices[:, :, None], indices[:, None, :]]


def sample(data: dict, batch: int, num_sample: int, with_duration: bool = False, indices=None,
           rng=np.random) -> dict:
    """`data` = {"points":[L,2], "distance":[L,L], "duration":[L,L]} (float64, create_dataset.py:169-174)."""
    if batch <= 0 or num_sample <= 0:
        raise ValueError("batch and num_sample must be positive integers.")
    length = len(data["points"])
    if num_sample > length:
        raise ValueError(f"num_sample ({num_sample}) exceeds the available data size ({length}).")
    if data["distance"].max() > 1e5:
        raise NotImplementedError("outlier-row removal (sampler.py:41-60) is host-side cleaning, not restated")
    if indices is None:
        indices = uniform_sample(batch, length, num_sample, rng)
    out = {"points": data["points"][indices], "distance_matrix": gather_submatrix(data["distance"], indices)}
    if with_duration:
        out["duration_matrix"] = gather_submatrix(data["duration"], indices)
    out["indices"] = indices
    return out

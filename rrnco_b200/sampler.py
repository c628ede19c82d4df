"""On-the-fly instance sub-sampling on the GPU: the reference's `Real_World_Sampler.sample`
(rrnco/envs/rcvrp/sampler.py:8-95; rmtvrp twin also gathers `duration`, rmtvrp/sampler.py:80) with the
NumPy fancy-index gather replaced by `rrnco_gather_submatrix` (fp64 city matrix -> fp32 instances, optional
fused reset normalisation).  Index sampling stays on the host with the reference's own RNG law
(`np.random.choice(L, n, replace=False)` per instance, sampler.py:97-104) so that seeds reproduce.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr


class CityOnDevice:
    """A city's float64 `distance` / `duration` / `points` arrays resident in HBM (8 MB per matrix)."""

    def __init__(self, data: dict, device="cuda"):
        self.length = len(data["points"])
        self.points = torch.as_tensor(np.asarray(data["points"]), dtype=torch.float64).to(device)
        self.distance = torch.as_tensor(np.asarray(data["distance"]), dtype=torch.float64).to(device).contiguous()
        self.duration = None
        if "duration" in data:
            self.duration = torch.as_tensor(np.asarray(data["duration"]), dtype=torch.float64).to(device).contiguous()
        self.max_distance = float(np.asarray(data["distance"]).max())
        # fp32 copies for the gather (made once per city; the fp32 cast commutes with the gather: same bits)
        self.distance_f32 = city_matrix_to_f32(self.distance)
        self.duration_f32 = city_matrix_to_f32(self.duration) if self.duration is not None else None


def city_matrix_to_f32(matrix: torch.Tensor) -> torch.Tensor:
    out = torch.empty(matrix.shape, dtype=torch.float32, device=matrix.device)
    call("rrnco_city_matrix_to_f32", ptr(matrix), matrix.numel(), ptr(out), stream_ptr(matrix.device))
    return out


def gather_submatrix(matrix: torch.Tensor, idx: torch.Tensor, normalize: bool = False):
    """out[b,i,j] = float32(matrix[idx[b,i], idx[b,j]]); with normalize also returns (min, max) per instance.
    `matrix` is the fp64 city matrix or its fp32 copy (CityOnDevice.distance_f32): identical results."""
    if matrix.dtype not in (torch.float64, torch.float32) or not matrix.is_contiguous():
        raise TypeError("gather_submatrix: contiguous float64 / float32 city matrix expected")
    idx = idx.to(device=matrix.device, dtype=torch.int32).contiguous()
    B, n = idx.shape
    out = torch.empty((B, n, n), dtype=torch.float32, device=matrix.device)
    mn = mx = None
    if normalize:
        mn = torch.empty(B, dtype=torch.float32, device=matrix.device)
        mx = torch.empty_like(mn)
    call("rrnco_gather_submatrix" if matrix.dtype == torch.float64 else "rrnco_gather_submatrix_f32", ptr(matrix), matrix.shape[0], ptr(idx), B, n, ptr(out), int(normalize), ptr(mn),
         ptr(mx), stream_ptr(matrix.device))
    return (out, mn, mx) if normalize else out


class Real_World_Sampler:
    def __init__(self, with_duration: bool = False, device="cuda"):
        self.with_duration = with_duration
        self.device = device

    def uniform_sample(self, batch, data_length, num_sample):
        return np.array([np.random.choice(data_length, num_sample, replace=False) for _ in range(batch)])

    def sample(self, data, batch: int, num_sample: int, loc_dist: str = "uniform", num_cluster: int = 5):
        """`data`: dict of numpy arrays (as upstream) or a CityOnDevice.  Returns device fp32 tensors
        {"points", "distance_matrix"[, "duration_matrix"]} (upstream returns float64 NumPy and the generator
        casts to fp32, rcvrp/generator_lazy.py:277,300)."""
        if batch <= 0 or num_sample <= 0:
            raise ValueError("batch and num_sample must be positive integers.")
        city = data if isinstance(data, CityOnDevice) else CityOnDevice(data, self.device)
        if num_sample > city.length:
            raise ValueError(f"num_sample ({num_sample}) exceeds the available data size ({city.length}).")
        if city.max_distance > 1e5:
            raise NotImplementedError("outlier-row removal (sampler.py:41-60) is host-side data cleaning: clean the "
                                      "city arrays before uploading them")
        if loc_dist != "uniform":
            raise NotImplementedError(f"loc_dist='{loc_dist}': only 'uniform' is on the training configs")
        indices = self.uniform_sample(batch, city.length, num_sample)
        idx = torch.from_numpy(indices.astype(np.int32)).to(city.distance.device, non_blocking=True)
        out = {"points": city.points[idx.long()].float(), "distance_matrix": gather_submatrix(city.distance_f32, idx)}
        if self.with_duration:
            out["duration_matrix"] = gather_submatrix(city.duration_f32, idx)
        return out

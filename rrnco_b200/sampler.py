"""On-the-fly instance sub-sampling on the GPU: the reference's `Real_World_Sampler.sample`
(rrnco/envs/rcvrp/sampler.py:8-95; rmtvrp twin also gathers `duration`, rmtvrp/sampler.py:80) with the
NumPy fancy-index gather replaced by `rrnco_gather_submatrix` (fp64 city matrix -> fp32 instances, optional
fused reset normalisation).  Index sampling: `uniform_sample` is the reference's own host law (`np.random.choice(L, n, replace=False)` per
instance, sampler.py:97-104: seeds reproduce); `uniform_sample_device` draws the same distribution (n distinct indices in
uniformly random order) for the whole batch on the device in one top-k of uniform keys.  Cities whose matrix holds
outliers (> 1e5) are cleaned once at upload with the reference's row / column rule (sampler.py:41-60).
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr


def remove_outlier_points(data: dict) -> dict:
    """Real_World_Sampler.sample's data cleaning (rrnco/envs/rcvrp/sampler.py:41-60), applied once per city instead of
    once per call: when the distance matrix holds entries > 1e5 (unreachable pairs in the OSRM tables), drop the
    points named by the FIRST row that has some but fewer than L/2 outliers and by the FIRST column that has fewer than
    L/2 (upstream's loops break at the first hit; a column without any outlier yields an empty set)."""
    distance = np.asarray(data["distance"])
    if not distance.max() > 1e5:
        return data
    L = len(data["points"])
    rows, cols = np.where(distance > 1e5)
    problem_r = problem_c = None
    for i in range(L):
        k = int((rows == i).sum())
        if 0 < k < L // 2:
            problem_r = cols[rows == i]
            break
    for i in range(L):
        if int((cols == i).sum()) < L // 2:
            problem_c = rows[cols == i]
            break
    if problem_r is None or problem_c is None:  # upstream raises UnboundLocalError here
        raise ValueError("outlier removal: no row / column with fewer than L/2 outliers (sampler.py:41-60 would fail)")
    keep = np.delete(np.arange(L), np.concatenate([problem_r, problem_c]))
    out = {"points": np.asarray(data["points"])[keep], "distance": distance[keep][:, keep]}
    if "duration" in data:
        out["duration"] = np.asarray(data["duration"])[keep][:, keep]
    return out


class CityOnDevice:
    """A city's float64 `distance` / `duration` / `points` arrays resident in HBM (8 MB per matrix)."""

    def __init__(self, data: dict, device="cuda"):
        data = remove_outlier_points(data)
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
    normalize = int(normalize)
    from . import torch_ops, _lib
    if torch_ops.enabled():
        _lib.count_call("rrnco_gather_submatrix")
        out, mn, mx = torch_ops.ops().gather_submatrix(matrix, idx, normalize)
        return (out, mn, mx) if normalize else out
    idx = idx.to(device=matrix.device, dtype=torch.int32).contiguous()
    B, n = idx.shape
    out = torch.empty((B, n, n), dtype=torch.float32, device=matrix.device)  # 0 none | 1 env.reset's distance law (eps 1e-6) | 2 generators' duration law (zero-range guard)
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

    @staticmethod
    def uniform_sample_device(batch, data_length, num_sample, device, generator=None):
        """Same law as `uniform_sample` -- per instance `num_sample` distinct indices of `range(data_length)`, every ordered
        selection equally likely -- for the whole batch at once on the device: the positions of the `num_sample` largest
        of `data_length` i.i.d. uniform keys, in decreasing key order.  (The host loop of np.random.choice costs 60-80 ms
        per 4096-instance batch, comparable to the rollout it feeds.)"""
        keys = torch.rand(batch, data_length, device=device, generator=generator)
        return keys.topk(num_sample, dim=1).indices.to(torch.int32)

    def sample(self, data, batch: int, num_sample: int, loc_dist: str = "uniform", num_cluster: int = 5, indices=None,
               normalize_duration: bool = False):
        """`data`: dict of numpy arrays (as upstream) or a CityOnDevice.  Returns device fp32 tensors
        {"points", "distance_matrix"[, "duration_matrix"]} (upstream returns float64 NumPy and the generator
        casts to fp32, rcvrp/generator_lazy.py:277,300)."""
        if batch <= 0 or num_sample <= 0:
            raise ValueError("batch and num_sample must be positive integers.")
        city = data if isinstance(data, CityOnDevice) else CityOnDevice(data, self.device)
        if num_sample > city.length:
            raise ValueError(f"num_sample ({num_sample}) exceeds the available data size ({city.length}).")
        if loc_dist != "uniform":
            raise NotImplementedError(f"loc_dist='{loc_dist}': only 'uniform' is on the training configs")
        if indices is None:
            indices = self.uniform_sample(batch, city.length, num_sample)
        if isinstance(indices, np.ndarray):
            indices = torch.from_numpy(indices.astype(np.int32))
        idx = indices.to(device=city.distance.device, dtype=torch.int32, non_blocking=True)
        out = {"points": city.points[idx.long()].float(), "distance_matrix": gather_submatrix(city.distance_f32, idx)}
        if self.with_duration:
            if normalize_duration:  # the generators' post-gather law fused into the gather (normalize mode 2)
                out["duration_matrix"] = gather_submatrix(city.duration_f32, idx, normalize=2)[0]
            else:
                out["duration_matrix"] = gather_submatrix(city.duration_f32, idx)
        return out

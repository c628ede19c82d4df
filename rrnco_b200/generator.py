"""On-device instance generation: the real-world branch of the reference's lazy generators.

    LazyRCVRPGenerator    rrnco/envs/rcvrp/generator_lazy.py:164-304
    LazyATSPGenerator     rrnco/envs/atsp/generator_lazy.py:239-260
    LazyRMTVRPGenerator   rrnco/envs/rmtvrp/generator_lazy.py:218-419 + rrnco/envs/rmtvrp/generator.py:352-432,445-469,515-607

Upstream builds every batch on the host: 10 random cities per chunk of `chunk_size` instances, `Real_World_Sampler.sample`
(NumPy fancy indexing of the float64 city matrices), `_process_real_world_data` (fp32 casts, min-max normalisation of
`locs` / `duration`, demand / time-window / distance-limit laws), `subsample_problems` (variant presets).  Here the city
matrices live in HBM (`sampler.CityOnDevice`), the sub-matrices come from `rrnco_gather_submatrix` (the duration law
fused into the gather), and the per-node laws run as element-wise device ops on `[B, n]` tensors.  What is NOT here: the
city list / .npz file handling and pickling hooks (host plumbing, out of scope: pass `cities=` -- dicts with the on-disk
schema or `CityOnDevice`s), the synthetic (no city data) branches, cluster location samplers.

Every law takes its uniform random draws from `draws` when given (a dict of [B, n] tensors in [0, 1)), so the arithmetic
is testable bit-for-bit against the reference driven by the same draws (tests/golden/generator_*.npz).
"""
from __future__ import annotations

import random
from typing import Optional

import torch

from .sampler import CityOnDevice, Real_World_Sampler
from .tdlite import TensorDictLite

CAPACITIES = {10: 20.0, 15: 25.0, 20: 30.0, 30: 33.0, 40: 37.0, 50: 40.0, 60: 43.0, 75: 45.0, 100: 50.0, 125: 55.0,
              150: 60.0, 200: 70.0, 500: 100.0, 1000: 150.0}  # rcvrp/generator_lazy.py:21-36

VARIANT_GENERATION_PRESETS = {  # rmtvrp/generator.py:37-57 (the deterministic ones and "all")
    "all": {"O": 0.5, "TW": 0.5, "L": 0.5, "B": 0.5},
    "cvrp": {"O": 0.0, "TW": 0.0, "L": 0.0, "B": 0.0}, "ovrp": {"O": 1.0, "TW": 0.0, "L": 0.0, "B": 0.0},
    "vrpb": {"O": 0.0, "TW": 0.0, "L": 0.0, "B": 1.0}, "vrpl": {"O": 0.0, "TW": 0.0, "L": 1.0, "B": 0.0},
    "vrptw": {"O": 0.0, "TW": 1.0, "L": 0.0, "B": 0.0}, "ovrptw": {"O": 1.0, "TW": 1.0, "L": 0.0, "B": 0.0},
    "ovrpb": {"O": 1.0, "TW": 0.0, "L": 0.0, "B": 1.0}, "ovrpl": {"O": 1.0, "TW": 0.0, "L": 1.0, "B": 0.0},
    "vrpbl": {"O": 0.0, "TW": 0.0, "L": 1.0, "B": 1.0}, "vrpbtw": {"O": 0.0, "TW": 1.0, "L": 0.0, "B": 1.0},
    "vrpltw": {"O": 0.0, "TW": 1.0, "L": 1.0, "B": 0.0}, "ovrpbl": {"O": 1.0, "TW": 0.0, "L": 1.0, "B": 1.0},
    "ovrpbtw": {"O": 1.0, "TW": 1.0, "L": 0.0, "B": 1.0}, "ovrpltw": {"O": 1.0, "TW": 1.0, "L": 1.0, "B": 0.0},
    "vrpbltw": {"O": 0.0, "TW": 1.0, "L": 1.0, "B": 1.0}, "ovrpbltw": {"O": 1.0, "TW": 1.0, "L": 1.0, "B": 1.0},
}


def get_vehicle_capacity(num_loc: int) -> int:
    """rmtvrp/generator.py:22-34."""
    if num_loc > 1000:
        return 200
    if num_loc > 20 and num_loc <= 1000:
        return 30 + num_loc // 5
    return 30


def _minmax_locs(points: torch.Tensor, eps: float) -> torch.Tensor:
    """(p - min) / (max - min + eps) per instance and axis, fp32 (generator_lazy.py: np.min / np.max over axis 1)."""
    lo = points.amin(dim=1, keepdim=True)
    hi = points.amax(dim=1, keepdim=True)
    return (points - lo) / (hi - lo + eps)


class _Draws:
    """Uniform [0, 1) draws: taken from a dict (golden tests) or generated on the device."""

    def __init__(self, draws, device, generator):
        self.draws, self.device, self.generator = draws, device, generator

    def __call__(self, name, *shape):
        if self.draws is not None:
            u = self.draws[name].to(self.device)
            assert tuple(u.shape) == tuple(shape), (name, tuple(u.shape), shape)
            return u
        return torch.rand(*shape, device=self.device, generator=self.generator)


class _LazyGeneratorBase:
    with_duration = False

    def __init__(self, num_loc: int = 20, cities=None, device="cuda", chunk_size: int = 1000, seed: Optional[int] = None,
                 index_sampling: str = "device", **kwargs):
        self.num_loc = num_loc
        self.device = torch.device(device)
        self.chunk_size = chunk_size
        self.loc_sampler = Real_World_Sampler(with_duration=self.with_duration, device=self.device)
        self.index_sampling = index_sampling  # "device": one top-k per city; "numpy": upstream's host loop (seeds reproduce)
        self.cities = []
        for c in (cities or []):
            self.cities.append(c if isinstance(c, CityOnDevice) else CityOnDevice(c, self.device))
        self.generator = None
        self._rng = random.Random(seed)
        if seed is not None and self.device.type == "cuda":
            self.generator = torch.Generator(device=self.device).manual_seed(seed)
        elif seed is not None:
            self.generator = torch.Generator().manual_seed(seed)

    # number of nodes sampled per instance (the VRPs add the depot)
    def _num_sample(self):
        return self.num_loc + 1

    def __call__(self, batch_size):
        batch_size = [batch_size] if isinstance(batch_size, int) else list(batch_size)
        return self._generate(batch_size)

    def _generate(self, batch_size) -> TensorDictLite:
        """generator_lazy.py:164-187: chunks of `chunk_size`, concatenated."""
        if not self.cities:
            raise ValueError("no city data: pass cities=[{points, distance[, duration]} | CityOnDevice, ...]")
        total = batch_size[0]
        tds = []
        for start in range(0, total, self.chunk_size):
            tds.append(self._generate_real_world_chunk([min(self.chunk_size, total - start)]))
        if len(tds) == 1:
            return self._finish(tds[0])
        keys = list(tds[0].keys())
        merged = {k: torch.cat([td[k] for td in tds], 0) for k in keys}
        return self._finish(TensorDictLite(merged, batch_size=[sum(td.batch_size[0] for td in tds)]))

    def _finish(self, td):
        return td

    def _generate_real_world_chunk(self, batch_size) -> TensorDictLite:
        """generator_lazy.py:196-246: min(10, #cities) random cities, `target // #cities` instances each (upstream's
        integer division: a chunk that does not divide evenly comes back slightly smaller than asked)."""
        target = batch_size[0]
        n_cities = min(10, len(self.cities))
        cities = self._rng.sample(self.cities, n_cities)
        sub = max(1, target // n_cities)
        parts, total = [], 0
        for city in cities:
            if total >= target:
                break
            cur = min(sub, target - total)
            if self.index_sampling == "device":
                idx = Real_World_Sampler.uniform_sample_device(cur, city.length, self._num_sample(), self.device, self.generator)
            else:
                idx = None
            parts.append(self.loc_sampler.sample(city, cur, self._num_sample(), indices=idx,
                                                 normalize_duration=self.with_duration))
            total += cur
        chunk = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
        return self._process_real_world_data(chunk, batch_size, duration_normalized=self.with_duration)


class LazyRCVRPGenerator(_LazyGeneratorBase):
    """rcvrp/generator_lazy.py: depot = first sampled point, raw (un-normalised) coordinates, continuous demands
    U(min_demand, max_demand) / capacity."""

    def __init__(self, num_loc: int = 20, min_demand: int = 1, max_demand: int = 10, vehicle_capacity: float = 1.0,
                 capacity: Optional[float] = None, **kwargs):
        super().__init__(num_loc=num_loc, **kwargs)
        self.min_demand, self.max_demand = min_demand, max_demand
        self.vehicle_capacity = vehicle_capacity
        self.capacity = CAPACITIES.get(num_loc, 50.0) if capacity is None else capacity

    def _process_real_world_data(self, chunk_data: dict, batch_size, draws=None, duration_normalized=False) -> TensorDictLite:
        """generator_lazy.py:275-304."""
        points = chunk_data["points"].float()
        B = points.shape[0]
        u = _Draws(draws, points.device, self.generator)
        # torch.distributions.Uniform(lo, hi).sample(): lo + rand * (hi - lo)
        demand = self.min_demand + u("demand", B, self.num_loc) * (self.max_demand - self.min_demand)
        td = {"locs": points[:, 1:, :].contiguous(), "depot": points[:, :1, :].contiguous(), "demand": demand / self.capacity,
              "capacity": torch.full((B, 1), self.capacity, dtype=torch.float32, device=points.device)}
        if "distance_matrix" in chunk_data:
            td["distance_matrix"] = chunk_data["distance_matrix"].float()
        return TensorDictLite(td, batch_size=[B])


class LazyATSPGenerator(_LazyGeneratorBase):
    """atsp/generator_lazy.py:239-260: locs min-max normalised per instance (eps 1e-6), distance as sampled."""

    def _num_sample(self):
        return self.num_loc

    def _process_real_world_data(self, chunk_data: dict, batch_size, draws=None, duration_normalized=False) -> TensorDictLite:
        points = chunk_data["points"].float()
        return TensorDictLite({"locs": _minmax_locs(points, 1e-6), "distance_matrix": chunk_data["distance_matrix"].float()},
                              batch_size=[points.shape[0]])


class LazyRMTVRPGenerator(_LazyGeneratorBase):
    """rmtvrp/generator_lazy.py + rmtvrp/generator.py: locs (eps 1e-8) and duration min-max normalised, integer linehaul /
    backhaul demands, Liu-et-al time windows from the duration matrix, distance limits, variant sub-sampling."""

    with_duration = True

    def __init__(self, num_loc: int = 20, capacity: Optional[float] = None, min_demand: int = 1, max_demand: int = 10,
                 min_backhaul: int = 1, max_backhaul: int = 10, scale_demand: bool = True, max_time: float = 4.6,
                 backhaul_ratio: float = 0.2, backhaul_class: int = 1, sample_backhaul_class: bool = False,
                 max_distance_limit: float = 2.8, speed: float = 1.0, variant_preset="vrptw", use_combinations: bool = False,
                 subsample: bool = True, **kwargs):
        super().__init__(num_loc=num_loc, **kwargs)
        self.capacity = get_vehicle_capacity(num_loc) if capacity is None else capacity
        self.min_demand, self.max_demand = min_demand, max_demand
        self.min_backhaul, self.max_backhaul = min_backhaul, max_backhaul
        self.scale_demand, self.max_time, self.backhaul_ratio = scale_demand, max_time, backhaul_ratio
        assert backhaul_class in (1, 2), "Backhaul class must be in [1, 2]"
        self.backhaul_class, self.sample_backhaul_class = backhaul_class, sample_backhaul_class
        self.max_distance_limit, self.speed = max_distance_limit, speed
        if variant_preset not in VARIANT_GENERATION_PRESETS:
            raise NotImplementedError(f"variant preset {variant_preset!r}: the single_feat presets / free probabilities draw "
                                      "one variant per instance from a categorical law and are not on the RRNCO configs")
        if sample_backhaul_class:
            raise NotImplementedError("sample_backhaul_class=True")
        self.variant_probs = VARIANT_GENERATION_PRESETS[variant_preset]
        self.variant_preset = variant_preset
        self.use_combinations = use_combinations and variant_preset == "all"  # generator.py:177-179
        if variant_preset == "all" and not self.use_combinations:
            raise NotImplementedError("variant_preset='all' without use_combinations draws from a categorical law")
        self.subsample = subsample

    def _finish(self, td):
        return self.subsample_problems(td) if self.subsample else td  # generator_lazy.py:255-257

    def generate_demands(self, B, u):
        """generator.py:445-469: (uniform_(lo - 1, hi - 1).int() + 1).float(), backhaul where rand <= backhaul_ratio."""
        n = self.num_loc
        lo, hi = float(self.min_demand - 1), float(self.max_demand - 1)
        linehaul = ((lo + (hi - lo) * u("linehaul", B, n)).int() + 1).float()
        lo, hi = float(self.min_backhaul - 1), float(self.max_backhaul - 1)
        backhaul = ((lo + (hi - lo) * u("backhaul", B, n)).int() + 1).float()
        is_linehaul = u("is_linehaul", B, n) > self.backhaul_ratio
        return linehaul * is_linehaul, backhaul * ~is_linehaul

    def generate_time_windows_with_duration_matrix(self, duration, u):
        """generator.py:515-562 (Liu et al. 2024 law on the asymmetric, normalised duration matrix)."""
        B, n = duration.shape[0], duration.shape[1] - 1
        a, b, c = 0.15, 0.18, 0.2
        service_time = a + (b - a) * u("service_time", B, n)
        tw_length = b + (c - b) * u("tw_length", B, n)
        d_0i = duration[:, 0, 1:]
        d_i0 = duration[:, 1:, 0]
        d_max = torch.max(d_0i, d_i0)
        h_max = (self.max_time - service_time - tw_length) / (d_max + 1e-6) - 1
        tw_start = d_0i + (h_max - 1) * d_max * u("tw_start", B, n)
        tw_end = tw_start + tw_length
        zeros = torch.zeros(B, 1, device=duration.device)
        time_windows = torch.stack((torch.cat((zeros, tw_start), -1),
                                    torch.cat((torch.full((B, 1), self.max_time, device=duration.device), tw_end), -1)), dim=-1)
        return time_windows, torch.cat((zeros, service_time), dim=-1)

    def generate_distance_limit(self, locs, u):
        """generator.py:564-582: Uniform(2 max_i |depot - loc_i| + 1e-6, max(max_distance_limit, that + 1e-6))."""
        d = (locs[:, 1:] - locs[:, 0:1]).square().sum(-1).sqrt()
        lower = 2 * d.amax(dim=1) + 1e-6
        upper = torch.maximum(torch.full_like(lower, self.max_distance_limit), lower + 1e-6)
        return (lower + u("distance_limit", locs.shape[0]) * (upper - lower))[..., None]

    def _process_real_world_data(self, chunk_data: dict, batch_size, draws=None, duration_normalized=False) -> TensorDictLite:
        """generator_lazy.py:350-419."""
        points = chunk_data["points"].float()
        distance = chunk_data["distance_matrix"].float()
        duration = chunk_data["duration_matrix"].float()
        B, dev = points.shape[0], points.device
        u = _Draws(draws, dev, self.generator)
        locs = _minmax_locs(points, 1e-8)
        if not duration_normalized:  # (the sampler fuses this law into the gather: normalize mode 2)
            lo = duration.amin(dim=(1, 2), keepdim=True)
            hi = duration.amax(dim=(1, 2), keepdim=True)
            duration = (duration - lo) / torch.where(hi - lo == 0, torch.ones_like(hi), hi - lo)
        vehicle_capacity = torch.full((B, 1), float(self.capacity), dtype=torch.float32, device=dev)
        capacity_original = vehicle_capacity.clone()
        demand_linehaul, demand_backhaul = self.generate_demands(B, u)
        backhaul_class = torch.full((B, 1), float(self.backhaul_class), dtype=torch.float32, device=dev)
        speed = torch.full((B, 1), float(self.speed), dtype=torch.float32, device=dev)
        time_windows, service_time = self.generate_time_windows_with_duration_matrix(duration, u)
        open_route = torch.ones((B, 1), dtype=torch.bool, device=dev)
        distance_limit = self.generate_distance_limit(locs, u)
        if self.scale_demand:
            demand_backhaul = demand_backhaul / vehicle_capacity
            demand_linehaul = demand_linehaul / vehicle_capacity
            vehicle_capacity = vehicle_capacity / vehicle_capacity
        return TensorDictLite({
            "locs": locs, "demand_backhaul": demand_backhaul, "demand_linehaul": demand_linehaul,
            "backhaul_class": backhaul_class, "distance_limit": distance_limit, "time_windows": time_windows,
            "service_time": service_time, "vehicle_capacity": vehicle_capacity, "capacity_original": capacity_original,
            "open_route": open_route, "speed": speed, "distance_matrix": distance, "duration_matrix": duration},
            batch_size=[B])

    def subsample_problems(self, td, draws=None):
        """generator.py:352-432 for the presets whose keep-mask is fixed, and "all" with use_combinations."""
        B, dev = td.batch_size[0], td["locs"].device
        probs = torch.tensor(list(self.variant_probs.values()), device=dev)
        if self.use_combinations:
            keep = _Draws(draws, dev, self.generator)("variant", B, 4) >= probs  # O, TW, L, B
        elif self.variant_preset == "cvrp":
            keep = torch.zeros(B, 4, dtype=torch.bool, device=dev)  # Categorical([0, 0, 0, 0, 0.5]) always picks "cvrp"
        else:
            keep = (probs > 0)[None].expand(B, 4)
        rm_o, rm_tw, rm_l, rm_b = (~keep[:, i] for i in range(4))
        td["open_route"][rm_o] = False
        default_tw = torch.zeros_like(td["time_windows"])
        default_tw[..., 1] = float("inf")
        td["time_windows"][rm_tw] = default_tw[rm_tw]
        td["service_time"][rm_tw] = 0.0
        td["distance_limit"][rm_l] = float("inf")
        td["demand_linehaul"][rm_b] = td["demand_linehaul"][rm_b] + td["demand_backhaul"][rm_b]
        td["demand_backhaul"][rm_b] = 0
        return td

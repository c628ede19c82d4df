"""Drop-in environments: same class names, methods, td keys / shapes / dtypes as the reference's
rl4co-style envs, with `_reset` normalisation, `_step`, `get_action_mask` and `_get_reward` running as
CUDA kernels of librrnco_b200 (include/rrnco_b200.h).

  ATSPEnv    <- rrnco/envs/atsp/env.py:79-220
  RCVRPEnv   <- rrnco/envs/rcvrp/env.py:90-249
  RMTVRPEnv  <- rrnco/envs/rmtvrp/env.py:155-455,566-570 (name "rcvrptw", all O/B/L/MB mask branches)

td containers: real `tensordict.TensorDict` when the caller uses one, else `TensorDictLite`.
There is no CPU path: td tensors are moved to `device` (CUDA) at reset.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import torch

from . import _lib
from ._lib import InstanceData, RmtvrpState, call, ptr, stream_ptr
from .tdlite import make_td_like


def _f32(t):
    return t.to(torch.float32).contiguous()


def _u8(t):
    """bool / uint8 tensor as a contiguous uint8 view (no copy for bool)."""
    t = t.contiguous()
    return t.view(torch.uint8) if t.dtype == torch.bool else t.to(torch.uint8)


def _lib_count(name):
    from . import _lib
    _lib.count_call(name)


def minmax_normalize(distance):
    """(d - min) / (max - min + 1e-6) per instance -> (normalised fp32, min [B], max [B])."""
    d = _f32(distance)
    from . import torch_ops
    if torch_ops.enabled():
        _lib_count("rrnco_minmax_normalize")
        return torch_ops.ops().minmax_normalize(d)
    B, N = d.shape[0], d.shape[-1]
    out = torch.empty_like(d)
    mn = torch.empty(B, dtype=torch.float32, device=d.device)
    mx = torch.empty_like(mn)
    call("rrnco_minmax_normalize", B, N, ptr(d), ptr(out), ptr(mn), ptr(mx), stream_ptr(d.device))
    return out, mn, mx


def tour_reward(actions, distance, prepend_depot, open_route=None, min_d=None, max_d=None):
    from . import torch_ops
    if torch_ops.enabled():
        _lib_count("rrnco_tour_reward")
        real, norm = torch_ops.ops().tour_reward(actions, distance, bool(prepend_depot), open_route, min_d, max_d)
        return (real if min_d is not None else None), norm
    actions = actions.contiguous()
    R, T = actions.shape
    dm = _f32(distance)
    N = dm.shape[-1]
    norm = torch.empty(R, dtype=torch.float32, device=dm.device)
    real = torch.empty_like(norm) if min_d is not None else None
    orp = None if open_route is None else _u8(open_route.reshape(-1))
    call("rrnco_tour_reward", R, T, N, dm.shape[0], ptr(actions), ptr(dm), int(prepend_depot), ptr(orp),
         ptr(None if min_d is None else _f32(min_d)), ptr(None if max_d is None else _f32(max_d)), ptr(norm),
         ptr(real), stream_ptr(dm.device))
    return real, norm


class _EnvBase:
    """The RL4COEnvBase surface the hot path uses."""

    name = "base"
    has_depot = True

    def __init__(self, generator=None, generator_params=None, normalize: bool = True, check_solution: bool = True,
                 device="cuda", **kwargs):
        gp = dict(generator_params or {})
        gp.pop("_target_", None)
        self.generator = generator if generator is not None else SimpleNamespace(
            num_loc=gp.get("num_loc", 20), vehicle_capacity=gp.get("vehicle_capacity", 1.0))
        self.normalize = normalize
        self.check_solution = check_solution
        self.device = torch.device(device)

    # -- rl4co wrappers ----------------------------------------------------------------------------
    def reset(self, td=None, batch_size=None):
        if td is None:
            raise ValueError("rrnco_b200 envs do not generate data: pass the instance td to reset()")
        if td.device.type != "cuda":
            td = td.to(self.device)  # H2D of the instance batch (part of the e2e timing)
        bs = list(td.batch_size) if batch_size is None else list(batch_size)
        out = self._reset(td, bs)
        out.set("done", torch.zeros(*bs, 1, dtype=torch.bool, device=td.device))
        return out

    def step(self, td):
        return {"next": self._step(td)}  # in place on td, like rl4co

    def get_reward(self, td, actions):
        if self.check_solution:
            self.check_solution_validity(td, actions)
        return self._get_reward(td, actions)

    def get_num_starts(self, td):
        return td["action_mask"].shape[-1]  # "rcvrp"/"atsp" are not in rl4co's minus-one list

    def select_start_nodes(self, td, num_starts):
        sel = torch.arange(num_starts, device=td.device).repeat_interleave(td.batch_size[0]) % self.generator.num_loc
        return sel + 1 if self.has_depot else sel

    def to(self, device):
        self.device = torch.device(device)
        return self


class ATSPEnv(_EnvBase):
    name = "atsp"
    has_depot = False

    def _reset(self, td, batch_size):
        dev = td.device
        distance = td["distance_matrix"]
        data = {}
        if self.normalize:
            distance, mn, mx = minmax_normalize(distance)
            data.update(min_distance=mn, max_distance=mx)
        cur = torch.zeros((*batch_size, 1), dtype=torch.int64, device=dev)
        data.update(distance_matrix=distance, first_node=cur, current_node=cur,
                    i=torch.zeros((*batch_size, 1), dtype=torch.int64, device=dev),
                    action_mask=torch.ones((*batch_size, distance.shape[-1]), dtype=torch.bool, device=dev))
        if "locs" in td.keys():
            data["locs"] = td["locs"]
        return make_td_like(td, data, batch_size)

    @staticmethod
    def _step(td):
        from . import torch_ops
        if torch_ops.enabled():
            _lib_count("rrnco_atsp_step")
            mask_out, first_out, cur_out, done = torch_ops.ops().atsp_step(td["action"], td["i"], td["action_mask"], td["first_node"])
            td.update({"first_node": first_out, "current_node": cur_out, "i": td["i"] + 1, "action_mask": mask_out,
                       "reward": torch.zeros_like(done), "done": done})
            return td
        action = td["action"].contiguous()
        mask_in = td["action_mask"].contiguous()
        R, N = mask_in.shape
        first_in = td["first_node"].reshape(-1).contiguous()
        mask_out = torch.empty_like(mask_in)
        first_out = torch.empty(R, dtype=torch.int64, device=action.device)
        cur_out = torch.empty_like(first_out)
        done = torch.empty(R, dtype=torch.bool, device=action.device)
        call("rrnco_atsp_step", R, N, ptr(action), ptr(td["i"]), ptr(mask_in), ptr(first_in), ptr(mask_out),
             ptr(first_out), ptr(cur_out), ptr(done), stream_ptr(action.device))
        td.update({"first_node": first_out, "current_node": cur_out, "i": td["i"] + 1, "action_mask": mask_out,
                   "reward": torch.zeros_like(done), "done": done})
        return td

    def _get_reward(self, td, actions):
        if self.normalize:
            return tour_reward(actions, td["distance_matrix"], False, None, td["min_distance"], td["max_distance"])
        return tour_reward(actions, td["distance_matrix"], False)[1]

    @staticmethod
    def check_solution_validity(td, actions):
        ref = torch.arange(actions.size(1), device=actions.device).view(1, -1).expand_as(actions)
        assert (ref == actions.sort(1)[0]).all(), "Invalid tour"


class RCVRPEnv(_EnvBase):
    name = "rcvrp"

    def _reset(self, td, batch_size):
        dev = td.device
        distance = td["distance_matrix"]
        data = {}
        if self.normalize:
            distance, mn, mx = minmax_normalize(distance)
            data.update(min_distance=mn, max_distance=mx)
        depot = td["depot"].unsqueeze(1) if td["depot"].ndim == 2 else td["depot"]
        n = td["locs"].shape[-2] + 1
        data.update(
            locs=torch.cat((depot, td["locs"]), dim=-2), distance_matrix=distance, demand=td["demand"],
            current_node=torch.zeros(*batch_size, 1, dtype=torch.long, device=dev),
            used_capacity=torch.zeros((*batch_size, 1), device=dev),
            vehicle_capacity=torch.full((*batch_size, 1), float(self.generator.vehicle_capacity), device=dev),
            visited=torch.zeros((*batch_size, n), dtype=torch.uint8, device=dev))
        out = make_td_like(td, data, batch_size)
        out.set("action_mask", self.get_action_mask(out))
        return out

    @staticmethod
    def _kernel(td, action):
        from . import torch_ops
        if torch_ops.enabled():
            _lib_count("rrnco_rcvrp_step")
            cur, used, visited, done, mask = torch_ops.ops().rcvrp_step(
                action, td["demand"], td["vehicle_capacity"], td["used_capacity"], td["visited"],
                td["current_node"] if action is None else None)
            return mask if action is None else (cur, used, visited, done, mask)
        visited_in = td["visited"].contiguous()
        R, N = visited_in.shape
        dev = visited_in.device
        demand, cap = _f32(td["demand"]), _f32(td["vehicle_capacity"]).reshape(-1)
        used_in = _f32(td["used_capacity"]).reshape(-1)
        mask = torch.empty((R, N), dtype=torch.bool, device=dev)
        if action is None:
            cur_in = td["current_node"].reshape(-1).contiguous()
            call("rrnco_rcvrp_step", R, N, demand.shape[0], None, ptr(demand), ptr(cap), cap.shape[0], ptr(used_in),
                 ptr(visited_in), ptr(cur_in), None, None, None, None, ptr(mask), stream_ptr(dev))
            return mask
        action = action.contiguous()
        used = torch.empty((R, 1), dtype=torch.float32, device=dev)
        visited = torch.empty_like(visited_in)
        cur = torch.empty((R, 1), dtype=torch.int64, device=dev)
        done = torch.empty(R, dtype=torch.bool, device=dev)
        call("rrnco_rcvrp_step", R, N, demand.shape[0], ptr(action), ptr(demand), ptr(cap), cap.shape[0],
             ptr(used_in), ptr(visited_in), None, ptr(used), ptr(visited), ptr(cur), ptr(done), ptr(mask),
             stream_ptr(dev))
        return cur, used, visited, done, mask

    def _step(self, td):
        cur, used, visited, done, mask = self._kernel(td, td["action"])
        td.update({"current_node": cur, "used_capacity": used, "visited": visited,
                   "reward": torch.zeros_like(done), "done": done})
        td.set("action_mask", mask)
        return td

    @staticmethod
    def get_action_mask(td):
        return RCVRPEnv._kernel(td, None)

    def _get_reward(self, td, actions):
        if self.normalize:
            return tour_reward(actions, td["distance_matrix"], True, None, td["min_distance"], td["max_distance"])
        return tour_reward(actions, td["distance_matrix"], True)[1]

    @staticmethod
    def check_solution_validity(td, actions):
        bsz, n = td["demand"].size()
        srt = actions.sort(1)[0]
        want = torch.arange(1, n + 1, device=actions.device).view(1, -1).expand(bsz, n)
        assert (want == srt[:, -n:]).all() and (srt[:, :-n] == 0).all(), "Invalid tour"
        d = torch.cat((-td["vehicle_capacity"], td["demand"]), 1).gather(1, actions)
        used = torch.zeros_like(td["demand"][:, 0])
        for i in range(actions.size(1)):
            used += d[:, i]
            used[used < 0] = 0
            assert (used <= td["vehicle_capacity"] + 1e-5).all(), "Used more than capacity"


class RMTVRPEnv(_EnvBase):
    name = "rcvrptw"

    def __init__(self, generator=None, generator_params=None, select_start_nodes_fn="all", normalize=True,
                 check_solution=False, **kwargs):
        super().__init__(generator, generator_params, normalize, check_solution, **kwargs)
        if select_start_nodes_fn != "all":
            raise NotImplementedError("only the default 'all' (POMO) start-node rule is implemented")

    def get_num_starts(self, td):
        return td["locs"].shape[-2] - 1

    def select_start_nodes(self, td, num_starts):
        n = td["locs"].shape[-2] - 1
        return torch.arange(num_starts, device=td.device).repeat_interleave(td.batch_size[0]) % n + 1

    def _reset(self, td, batch_size):
        dev = td.device
        keys = td.keys()
        zero_col = torch.zeros_like(td["demand_linehaul"][..., :1])
        dl = torch.cat([zero_col, td["demand_linehaul"]], dim=1)
        db = td["demand_backhaul"] if "demand_backhaul" in keys else torch.zeros_like(td["demand_linehaul"])
        db = torch.cat([zero_col, db], dim=1)
        bclass = td["backhaul_class"] if "backhaul_class" in keys else torch.full(
            (*batch_size, 1), 1, dtype=torch.int32, device=dev)
        if "time_windows" in keys:
            tw = td["time_windows"]
        else:
            tw = torch.zeros_like(td["locs"])
            tw[..., 1] = float("inf")
        service = td["service_time"] if "service_time" in keys else torch.zeros_like(dl)
        open_route = td["open_route"] if "open_route" in keys else torch.zeros_like(dl[..., :1], dtype=torch.bool)
        limit = td["distance_limit"] if "distance_limit" in keys else torch.full_like(dl[..., :1], float("inf"))
        dm = td["distance_matrix"] if "distance_matrix" in keys else torch.cdist(td["locs"], td["locs"], p=2)
        data = {}
        if self.normalize:
            dm, mn, mx = minmax_normalize(dm)
            data.update(min_distance=mn, max_distance=mx)
        speed = td["speed"] if "speed" in keys else torch.ones_like(dl[..., :1])
        dur = td["duration_matrix"] if "duration_matrix" in keys else dm / speed[:, None]
        ones = torch.ones_like(dl[..., :1])
        data.update(
            locs=td["locs"], distance_matrix=dm, duration_matrix=dur, demand_backhaul=db, demand_linehaul=dl,
            backhaul_class=bclass, distance_limit=limit, service_time=service, open_route=open_route,
            time_windows=tw, speed=speed,
            vehicle_capacity=td["vehicle_capacity"] if "vehicle_capacity" in keys else ones,
            capacity_original=td["capacity_original"] if "capacity_original" in keys else ones,
            current_node=torch.zeros((*batch_size,), dtype=torch.long, device=dev),
            current_route_length=torch.zeros((*batch_size, 1), dtype=torch.float32, device=dev),
            current_time=torch.zeros((*batch_size, 1), dtype=torch.float32, device=dev),
            used_capacity_backhaul=torch.zeros((*batch_size, 1), device=dev),
            used_capacity_linehaul=torch.zeros((*batch_size, 1), device=dev),
            visited=torch.zeros((*batch_size, td["locs"].shape[-2]), dtype=torch.bool, device=dev))
        out = make_td_like(td, data, batch_size)
        out.set("action_mask", self.get_action_mask(out))
        return out

    @staticmethod
    def instance_data(td, keep):
        """InstanceData struct over the td's tensors (`keep` holds the contiguous fp32 views alive)."""
        def put(t, dtype=torch.float32):
            t = _u8(t) if dtype == torch.uint8 else t.to(dtype).contiguous()
            keep.append(t)
            return ptr(t)

        d = InstanceData()
        d.data_rows = td["distance_matrix"].shape[0]
        d.distance = put(td["distance_matrix"])
        d.duration = put(td["duration_matrix"])
        d.demand = put(td["demand_linehaul"])
        d.demand_backhaul = put(td["demand_backhaul"])
        d.time_windows = put(td["time_windows"])
        d.service_time = put(td["service_time"])
        d.vehicle_capacity = put(td["vehicle_capacity"].reshape(-1))
        d.distance_limit = put(td["distance_limit"].reshape(-1))
        d.open_route = put(td["open_route"].reshape(-1), torch.uint8)
        d.backhaul_class = put(td["backhaul_class"].reshape(-1))
        if "min_distance" in td.keys():
            d.min_distance = put(td["min_distance"])
            d.max_distance = put(td["max_distance"])
        return d

    @staticmethod
    def _state(td, keep):
        def put(t, dtype):
            t = _u8(t) if dtype == torch.uint8 else t.reshape(-1).to(dtype).contiguous()
            keep.append(t)
            return ptr(t)

        s = RmtvrpState()
        s.current_node = put(td["current_node"], torch.int64)
        s.current_time = put(td["current_time"], torch.float32)
        s.current_route_length = put(td["current_route_length"], torch.float32)
        s.used_capacity_linehaul = put(td["used_capacity_linehaul"], torch.float32)
        s.used_capacity_backhaul = put(td["used_capacity_backhaul"], torch.float32)
        s.visited = put(td["visited"], torch.uint8)
        return s

    @staticmethod
    def _kernel(td, action):
        keep = []
        R, N = td["visited"].shape
        dev = td["visited"].device
        data = RMTVRPEnv.instance_data(td, keep)
        s_in = RMTVRPEnv._state(td, keep)
        mask = torch.empty((R, N), dtype=torch.bool, device=dev)
        if action is None:
            call("rrnco_rmtvrp_step", R, N, C.byref(data), None, C.byref(s_in), None, None, ptr(mask),
                 stream_ptr(dev))
            return mask
        action = action.contiguous()
        out = {
            "current_node": torch.empty(R, dtype=torch.int64, device=dev),
            "current_time": torch.empty((R, 1), dtype=torch.float32, device=dev),
            "current_route_length": torch.empty((R, 1), dtype=torch.float32, device=dev),
            "used_capacity_linehaul": torch.empty((R, 1), dtype=torch.float32, device=dev),
            "used_capacity_backhaul": torch.empty((R, 1), dtype=torch.float32, device=dev),
            "visited": torch.empty((R, N), dtype=torch.bool, device=dev),
        }
        s_out = RmtvrpState()
        for k, v in out.items():
            setattr(s_out, k, ptr(v))
        done = torch.empty(R, dtype=torch.bool, device=dev)
        call("rrnco_rmtvrp_step", R, N, C.byref(data), ptr(action), C.byref(s_in), C.byref(s_out), ptr(done),
             ptr(mask), stream_ptr(dev))
        out["done"] = done
        out["reward"] = torch.zeros_like(done).float()
        return out, mask

    def _step(self, td):
        out, mask = self._kernel(td, td["action"])
        td.update(out)
        td.set("action_mask", mask)
        return td

    @staticmethod
    def get_action_mask(td):
        return RMTVRPEnv._kernel(td, None)

    def _get_reward(self, td, actions):
        # upstream zeroes column 0 in place for open routes (rmtvrp/env.py:433) -- on the batchified td it was handed, a
        # copy that lives for this call.  The same visible side effect is kept for a td in that layout (one matrix per
        # rollout).  An un-replicated reset td (data_rows < R: rollout r reads row r % data_rows) is read again by later
        # rollouts / decoder calls, so it is NOT touched: the kernel multiplies the legs into the depot by 0 for
        # open-route rows, which gives the same sum.
        cm = td["distance_matrix"]
        if cm.shape[0] == actions.shape[0]:
            cm[:, :, 0] = cm[:, :, 0] * ~td["open_route"]
        if self.normalize:
            return tour_reward(actions, cm, True, td["open_route"], td["min_distance"], td["max_distance"])
        return tour_reward(actions, cm, True, td["open_route"])[1]

    @staticmethod
    def check_solution_validity(td, actions):
        raise NotImplementedError(
            "This method is not implemented yet. Please modify considering the dist and dur matrices.")


def get_env(name: str, **kw):
    return {"atsp": ATSPEnv, "rcvrp": RCVRPEnv, "rcvrptw": RMTVRPEnv}[name](**kw)

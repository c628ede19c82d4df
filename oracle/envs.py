"""CPU oracle of the three RRNCO environments (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, statement by statement:
  ATSP    rrnco/envs/atsp/env.py:79-220
  RCVRP   rrnco/envs/rcvrp/env.py:90-249
  RCVRPTW rrnco/envs/rmtvrp/env.py:155-455,566-570 + rrnco/envs/rmtvrp/selectstartnodes.py:31-50
and the rl4co RL4COEnvBase reset/step/get_reward/select_start_nodes/get_num_starts
wrappers.  Arithmetic order is kept exactly as written upstream
because masks must be bit-exact.
"""
from __future__ import annotations

import torch

from .td import TD, gather_by_index


def _minmax_normalise(distance):
    # rcvrp/env.py:138-145, atsp/env.py:113-120, rmtvrp/env.py:271-280
    lo = distance.amin(dim=(-2, -1), keepdim=True)
    hi = distance.amax(dim=(-2, -1), keepdim=True)
    d = ((distance - lo) / (hi - lo + 1e-6)).to(torch.float32)
    return d, lo.squeeze(-1).squeeze(-1), hi.squeeze(-1).squeeze(-1)


def _real_reward(neg_len, td):
    # rcvrp/env.py:212-217 (adds min once, not per edge)
    return neg_len * (td["max_distance"] - td["min_distance"] + 1e-6) + td["min_distance"]


class _EnvBase:
    """rl4co RL4COEnvBase surface used by the hot path (App. A)."""

    name = "base"
    has_depot = True

    def __init__(self, num_loc: int, normalize: bool = True, check_solution: bool = True):
        self.num_loc = num_loc
        self.normalize = normalize
        self.check_solution = check_solution

    def reset(self, td: TD) -> TD:
        out = self._reset(td, list(td.batch_size))
        out["done"] = torch.zeros(*td.batch_size, 1, dtype=torch.bool)  # torchrl fills done=False
        return out

    def step(self, td: TD) -> dict:
        return {"next": self._step(td)}  # in place, same object

    def get_reward(self, td: TD, actions):
        if self.check_solution:
            self.check_solution_validity(td, actions)
        return self._get_reward(td, actions)

    def get_num_starts(self, td: TD) -> int:
        # rl4co get_num_starts: action-mask width; "rcvrp"/"atsp" are not in rl4co's minus-one list
        return td["action_mask"].shape[-1]

    def select_start_nodes(self, td: TD, num_starts: int):
        sel = torch.arange(num_starts).repeat_interleave(td.batch_size[0]) % self.num_loc
        return sel + 1 if self.has_depot else sel


class ATSPEnv(_EnvBase):
    name = "atsp"
    has_depot = False

    def _reset(self, td, batch_size):  # atsp/env.py:107-155
        distance = td["distance_matrix"]
        out = {}
        if self.normalize:
            distance, lo, hi = _minmax_normalise(distance)
            out.update(min_distance=lo, max_distance=hi)
        cur = torch.zeros((*batch_size, 1), dtype=torch.int64)
        out.update(
            distance_matrix=distance,
            first_node=cur,
            current_node=cur,
            i=torch.zeros((*batch_size, 1), dtype=torch.int64),
            action_mask=torch.ones((*batch_size, distance.shape[-1]), dtype=torch.bool),
        )
        if "locs" in td:
            out["locs"] = td["locs"]
        return TD(out, batch_size=batch_size)

    @staticmethod
    def _step(td):  # atsp/env.py:79-105
        cur = td["action"]
        first = cur if int(td["i"].flatten()[0]) == 0 else td["first_node"]
        avail = td["action_mask"].scatter(-1, cur.unsqueeze(-1).expand_as(td["action_mask"]), 0)
        done = torch.count_nonzero(avail, dim=-1) <= 0
        td.update(first_node=first, current_node=cur, i=td["i"] + 1, action_mask=avail,
                  reward=torch.zeros_like(done), done=done)
        return td

    def _get_reward(self, td, actions):  # atsp/env.py:192-211
        dm = td["distance_matrix"]
        b = torch.arange(dm.shape[0]).unsqueeze(1)
        neg = -dm[b, actions, torch.roll(actions, -1, dims=1)].sum(-1)
        if self.normalize:
            return _real_reward(neg, td), neg
        return neg

    @staticmethod
    def check_solution_validity(td, actions):  # atsp/env.py:213-220
        ref = torch.arange(actions.size(1)).view(1, -1).expand_as(actions)
        assert (ref == actions.sort(1)[0]).all(), "Invalid tour"


class RCVRPEnv(_EnvBase):
    name = "rcvrp"

    def __init__(self, num_loc, normalize=True, check_solution=True, vehicle_capacity: float = 1.0):
        super().__init__(num_loc, normalize, check_solution)
        self.vehicle_capacity = vehicle_capacity  # generator.vehicle_capacity (rcvrp/generator.py)

    def _reset(self, td, batch_size):  # rcvrp/env.py:124-181
        distance = td["distance_matrix"]
        out = {}
        if self.normalize:
            distance, lo, hi = _minmax_normalise(distance)
            out.update(min_distance=lo, max_distance=hi)
        depot = td["depot"].unsqueeze(1) if td["depot"].ndim == 2 else td["depot"]
        out.update(
            locs=torch.cat((depot, td["locs"]), dim=-2),
            distance_matrix=distance,
            demand=td["demand"],
            current_node=torch.zeros(*batch_size, 1, dtype=torch.long),
            used_capacity=torch.zeros((*batch_size, 1)),
            vehicle_capacity=torch.full((*batch_size, 1), self.vehicle_capacity),
            visited=torch.zeros((*batch_size, td["locs"].shape[-2] + 1), dtype=torch.uint8),
        )
        res = TD(out, batch_size=batch_size)
        res["action_mask"] = self.get_action_mask(res)
        return res

    def _step(self, td):  # rcvrp/env.py:90-122
        cur = td["action"][:, None]
        n_loc = td["demand"].size(-1)
        sel = gather_by_index(td["demand"], torch.clamp(cur - 1, 0, n_loc - 1), squeeze=False)
        used = (td["used_capacity"] + sel) * (cur != 0).float()
        visited = td["visited"].scatter(-1, cur, 1)
        done = visited.sum(-1) == visited.size(-1)
        td.update(current_node=cur, used_capacity=used, visited=visited,
                  reward=torch.zeros_like(done), done=done)
        td["action_mask"] = self.get_action_mask(td)
        return td

    @staticmethod
    def get_action_mask(td):  # rcvrp/env.py:183-195
        exceeds = td["demand"] + td["used_capacity"] > td["vehicle_capacity"]
        mask_loc = td["visited"][..., 1:].to(exceeds.dtype) | exceeds
        mask_depot = (td["current_node"] == 0) & ((mask_loc == 0).int().sum(-1) > 0)[:, None]
        return ~torch.cat((mask_depot, mask_loc), -1)

    def _get_reward(self, td, actions):  # rcvrp/env.py:197-219
        dm = td["distance_matrix"]
        go_from = torch.cat((torch.zeros_like(actions[:, :1]), actions), dim=1)
        go_to = torch.roll(go_from, -1, dims=1)
        legs = gather_by_index(gather_by_index(dm, go_from, dim=1, squeeze=False),
                               go_to, dim=2, squeeze=False).squeeze(-1)
        neg = -legs.sum(-1)
        if self.normalize:
            return _real_reward(neg, td), neg
        return neg

    @staticmethod
    def check_solution_validity(td, actions):  # rcvrp/env.py:221-249
        bsz, n = td["demand"].size()
        srt = actions.sort(1)[0]
        want = torch.arange(1, n + 1).view(1, -1).expand(bsz, n)
        assert (want == srt[:, -n:]).all() and (srt[:, :-n] == 0).all(), "Invalid tour"
        d = torch.cat((-td["vehicle_capacity"], td["demand"]), 1).gather(1, actions)
        used = torch.zeros_like(td["demand"][:, 0])
        for i in range(actions.size(1)):
            used += d[:, i]
            used[used < 0] = 0
            assert (used <= td["vehicle_capacity"] + 1e-5).all(), "Used more than capacity"


class RMTVRPEnv(_EnvBase):
    """name = "rcvrptw" (rmtvrp/env.py:107); all O/B/L/MB branches of the mask are kept."""

    name = "rcvrptw"

    def __init__(self, num_loc, normalize=True, check_solution=False):
        super().__init__(num_loc, normalize, check_solution)

    def get_num_starts(self, td):  # selectstartnodes.py:31-34
        return td["locs"].shape[-2] - 1

    def select_start_nodes(self, td, num_starts):  # selectstartnodes.py:42-50
        n = td["locs"].shape[-2] - 1
        return torch.arange(num_starts).repeat_interleave(td.batch_size[0]) % n + 1

    def _reset(self, td, batch_size):  # rmtvrp/env.py:217-341
        zero_col = torch.zeros_like(td["demand_linehaul"][..., :1])
        dl = torch.cat([zero_col, td["demand_linehaul"]], dim=1)
        db = td.get("demand_backhaul", torch.zeros_like(td["demand_linehaul"]))
        db = torch.cat([zero_col, db], dim=1)
        bclass = td.get("backhaul_class", torch.full((*batch_size, 1), 1, dtype=torch.int32))
        tw = td.get("time_windows", None)
        if tw is None:
            tw = torch.zeros_like(td["locs"])
            tw[..., 1] = float("inf")
        service = td.get("service_time", torch.zeros_like(dl))
        open_route = td.get("open_route", torch.zeros_like(dl[..., :1], dtype=torch.bool))
        limit = td.get("distance_limit", torch.full_like(dl[..., :1], float("inf")))
        dm = td["distance_matrix"] if "distance_matrix" in td else torch.cdist(td["locs"], td["locs"], p=2)
        out = {}
        if self.normalize:
            dm, lo, hi = _minmax_normalise(dm)
            out.update(min_distance=lo, max_distance=hi)
        speed = td.get("speed", torch.ones_like(dl[..., :1]))
        dur = td["duration_matrix"] if "duration_matrix" in td else dm / speed[:, None]
        out.update(
            locs=td["locs"], distance_matrix=dm, duration_matrix=dur,
            demand_backhaul=db, demand_linehaul=dl, backhaul_class=bclass,
            distance_limit=limit, service_time=service, open_route=open_route,
            time_windows=tw, speed=speed,
            vehicle_capacity=td.get("vehicle_capacity", torch.ones_like(dl[..., :1])),
            capacity_original=td.get("capacity_original", torch.ones_like(dl[..., :1])),
            current_node=torch.zeros((*batch_size,), dtype=torch.long),
            current_route_length=torch.zeros((*batch_size, 1)),
            current_time=torch.zeros((*batch_size, 1)),
            used_capacity_backhaul=torch.zeros((*batch_size, 1)),
            used_capacity_linehaul=torch.zeros((*batch_size, 1)),
            visited=torch.zeros((*batch_size, td["locs"].shape[-2]), dtype=torch.bool),
        )
        res = TD(out, batch_size=batch_size)
        res["action_mask"] = self.get_action_mask(res)
        return res

    def _step(self, td):  # rmtvrp/env.py:155-215
        prev, cur = td["current_node"], td["action"]
        b = torch.arange(td.batch_size[0])
        dist = td["distance_matrix"][b, prev, cur]
        dur = td["duration_matrix"][b, prev, cur]
        service = gather_by_index(td["service_time"], cur, dim=1, squeeze=False)
        start = gather_by_index(td["time_windows"], cur, dim=1, squeeze=False)[..., 0]
        away = cur[:, None] != 0
        time = away * (torch.max(td["current_time"] + dur[:, None], start) + service)
        route = away * (td["current_route_length"] + dist[:, None])
        sel_l = gather_by_index(td["demand_linehaul"], cur, dim=1, squeeze=False)
        sel_b = gather_by_index(td["demand_backhaul"], cur, dim=1, squeeze=False)
        used_l = away * (td["used_capacity_linehaul"] + sel_l)
        used_b = away * (td["used_capacity_backhaul"] + sel_b)
        visited = td["visited"].scatter(-1, cur[..., None], True)
        done = visited.sum(-1) == visited.size(-1)
        td.update(current_node=cur, current_route_length=route, current_time=time, done=done,
                  reward=torch.zeros_like(done).float(), used_capacity_linehaul=used_l,
                  used_capacity_backhaul=used_b, visited=visited)
        td["action_mask"] = self.get_action_mask(td)
        return td

    @staticmethod
    def get_action_mask(td):  # rmtvrp/env.py:343-428
        cur = td["current_node"]
        b = torch.arange(td.batch_size[0])
        dist_ij = td["distance_matrix"][b, cur, :]
        dist_j0 = td["distance_matrix"][:, :, 0]
        dur_ij = td["duration_matrix"][b, cur, :]
        dur_j0 = td["duration_matrix"][:, :, 0]
        early, late = td["time_windows"][..., 0], td["time_windows"][..., 1]
        closed = ~td["open_route"]
        arrival = td["current_time"] + dur_ij
        can_reach_customer = arrival < late
        can_reach_depot = (torch.max(arrival, early) + td["service_time"] + dur_j0) * closed < late[..., 0:1]
        exceeds_limit = td["current_route_length"] + dist_ij + (dist_j0 * closed) > td["distance_limit"]
        exc_l = td["demand_linehaul"] + td["used_capacity_linehaul"] > td["vehicle_capacity"]
        exc_b = td["demand_backhaul"] + td["used_capacity_backhaul"] > td["vehicle_capacity"]
        linehauls_missing = ((td["demand_linehaul"] * ~td["visited"]).sum(-1) > 0)[..., None]
        carrying_b = gather_by_index(td["demand_backhaul"], cur, dim=1, squeeze=False) > 0
        ok1 = (linehauls_missing & ~exc_l & ~carrying_b & (td["demand_linehaul"] > 0)) | (
            ~exc_b & (td["demand_backhaul"] > 0))
        cannot_l = td["demand_linehaul"] > td["vehicle_capacity"] - td["used_capacity_backhaul"]
        ok2 = ~exc_l & ~exc_b & ~cannot_l
        ok = ((td["backhaul_class"] == 1) & ok1) | ((td["backhaul_class"] == 2) & ok2)
        can = can_reach_customer & can_reach_depot & ok & ~exceeds_limit & ~td["visited"]
        can[:, 0] = ~((cur == 0) & (can[:, 1:].sum(-1) > 0))
        return can

    def _get_reward(self, td, actions):  # rmtvrp/env.py:430-455 (mutates column 0, :433)
        cm = td["distance_matrix"]
        cm[:, :, 0] = cm[:, :, 0] * ~td["open_route"]
        go_from = torch.cat((torch.zeros_like(actions[:, :1]), actions), dim=1)
        go_to = torch.roll(go_from, -1, dims=1)
        legs = gather_by_index(gather_by_index(cm, go_from, dim=1, squeeze=False),
                               go_to, dim=2, squeeze=False).squeeze(-1)
        neg = -legs.sum(-1)
        if self.normalize:
            return _real_reward(neg, td), neg
        return neg

    @staticmethod
    def check_solution_validity(td, actions):  # rmtvrp/env.py:458-461
        raise NotImplementedError("upstream raises here too")


def make_env(name: str, num_loc: int, **kw):
    return {"atsp": ATSPEnv, "rcvrp": RCVRPEnv, "rcvrptw": RMTVRPEnv}[name](num_loc, **kw)


# rmtvrp/utils.py:63-79 -- forced action-sequence driver used by the parity tests
def rollout_actions(env, td, actions):
    for i in range(actions.size(1)):
        td["action"] = actions[:, i]
        td = env.step(td)["next"]
    return td

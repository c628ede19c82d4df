"""Tiny driver for ncu: two launches each of the RCVRP / ATSP / RCVRPTW env-step kernels (C2 / C2 / C3 rollout counts,
reference layout) and of the gather (+ fused normalisation, fp32 city copy):

    ncu --set full --clock-control none --import-source on -k regex:"step_vec|rmtvrp_step|gather_submatrix" \
        -o gpurun_out/prof_env python tools/profile_env_kernels.py
"""
import ctypes as C
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from rrnco_b200._lib import call, ptr, stream_ptr  # noqa: E402
from rrnco_b200.envs import RMTVRPEnv  # noqa: E402
from rrnco_b200.sampler import CityOnDevice, gather_submatrix  # noqa: E402
from bench import make_city  # noqa: E402
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
# RCVRP
R, N = 8192 * 101, 101
demand = torch.rand(R, N - 1, device=dev, generator=g) * 0.2
cap = torch.ones(R, device=dev); used = torch.rand(R, device=dev, generator=g) * 0.5
visited = (torch.rand(R, N, device=dev, generator=g) < 0.3).to(torch.uint8)
action = torch.randint(1, N, (R,), device=dev, generator=g)
used_o, vis_o = torch.empty_like(used), torch.empty_like(visited)
cur_o = torch.empty(R, dtype=torch.int64, device=dev); done_o = torch.empty(R, dtype=torch.bool, device=dev)
mask_o = torch.empty(R, N, dtype=torch.bool, device=dev)
for _ in range(2):
    call("rrnco_rcvrp_step", R, N, R, ptr(action), ptr(demand), ptr(cap), R, ptr(used), ptr(visited), None, ptr(used_o),
         ptr(vis_o), ptr(cur_o), ptr(done_o), ptr(mask_o), stream_ptr(dev))
# ATSP
Na = 100
mask_in = torch.rand(R, Na, device=dev, generator=g) < 0.7
act_a = torch.randint(0, Na, (R,), device=dev, generator=g)
step_i = torch.full((1,), 5, dtype=torch.int64, device=dev)
first_in = torch.randint(0, Na, (R,), device=dev, generator=g)
mask_a = torch.empty_like(mask_in); first_o = torch.empty_like(first_in); cur_a = torch.empty_like(first_in)
for _ in range(2):
    call("rrnco_atsp_step", R, Na, ptr(act_a), ptr(step_i), ptr(mask_in), ptr(first_in), ptr(mask_a), ptr(first_o),
         ptr(cur_a), ptr(done_o), stream_ptr(dev))
# RCVRPTW (C3: 1024 instances x 100 starts, instance data un-replicated)
Bt, St = 1024, 100
env_tw = RMTVRPEnv(generator_params={"num_loc": N - 1}, check_solution=False, device=dev)
dm = torch.rand(Bt, N, N, device=dev, generator=g)
td0 = env_tw.reset(rb.TensorDictLite({
    "locs": torch.rand(Bt, N, 2, device=dev, generator=g), "distance_matrix": dm,
    "duration_matrix": dm * (0.5 + torch.rand(Bt, N, N, device=dev, generator=g)),
    "demand_linehaul": torch.rand(Bt, N - 1, device=dev, generator=g) * 0.2,
    "time_windows": torch.stack([torch.rand(Bt, N, device=dev, generator=g), 4 + torch.rand(Bt, N, device=dev, generator=g)], -1),
    "service_time": torch.rand(Bt, N, device=dev, generator=g) * 0.05}, batch_size=[Bt]))
td_r = rb.batchify(td0, St)
Rt = Bt * St
keep = []
data = RMTVRPEnv.instance_data(td0, keep)
td_r.update({"visited": torch.rand(Rt, N, device=dev, generator=g) < 0.3,
             "current_node": torch.randint(1, N, (Rt,), device=dev, generator=g),
             "current_time": torch.rand(Rt, 1, device=dev, generator=g)})
s_in = RMTVRPEnv._state(td_r, keep)
out = {k: torch.empty_like(td_r[k].reshape(-1) if k != "visited" else td_r[k]) for k in
       ("current_node", "current_time", "current_route_length", "used_capacity_linehaul", "used_capacity_backhaul", "visited")}
s_out = type(s_in)()
for k, v in out.items():
    setattr(s_out, k, ptr(v))
act_t = torch.randint(1, N, (Rt,), device=dev, generator=g)
done_t = torch.empty(Rt, dtype=torch.bool, device=dev); mask_t = torch.empty(Rt, N, dtype=torch.bool, device=dev)
for _ in range(2):
    call("rrnco_rmtvrp_step", Rt, N, C.byref(data), ptr(act_t), C.byref(s_in), C.byref(s_out), ptr(done_t), ptr(mask_t),
         stream_ptr(dev))
# gather
city = CityOnDevice(make_city(3), dev)
rng = np.random.RandomState(1)
idx = torch.from_numpy(np.array([rng.choice(1000, N, replace=False) for _ in range(4096)])).to(dev)
for _ in range(2):
    gather_submatrix(city.distance_f32, idx, normalize=True)
torch.cuda.synchronize()
print("ok")

"""Tiny driver for ncu: one RCVRP env-step launch at the C2 rollout count and one gather launch."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from rrnco_b200._lib import call, ptr, stream_ptr  # noqa: E402
from rrnco_b200.sampler import CityOnDevice, gather_submatrix  # noqa: E402
from bench import make_city  # noqa: E402
dev = torch.device("cuda", 0)
R, N = 8192 * 101, 101
g = torch.Generator(device=dev).manual_seed(0)
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
city = CityOnDevice(make_city(3), dev)
rng = np.random.RandomState(1)
idx = torch.from_numpy(np.array([rng.choice(1000, N, replace=False) for _ in range(4096)])).to(dev)
for _ in range(2):
    gather_submatrix(city.distance, idx, normalize=True)
torch.cuda.synchronize()
print("ok")

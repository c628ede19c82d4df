"""Standalone timing of the RCVRP env-step kernel at the C2 rollout count (reference layout), for configuration sweeps:

    for c in 2x6 3x4 4x3 2x5; do RRNCO_STEP_CFG=$c python tools/env_step_probe.py; done
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rrnco_b200._lib import call, ptr, stream_ptr  # noqa: E402

dev = torch.device("cuda", 0)
R, N = int(os.environ.get("R", 8192 * 101)), int(os.environ.get("N", 101))
g = torch.Generator(device=dev).manual_seed(0)
demand = torch.rand(R, N - 1, device=dev, generator=g) * 0.2
cap = torch.ones(R, device=dev)
used = torch.rand(R, device=dev, generator=g) * 0.5
visited = (torch.rand(R, N, device=dev, generator=g) < 0.3).to(torch.uint8)
action = torch.randint(1, N, (R,), device=dev, generator=g)
used_o, vis_o = torch.empty_like(used), torch.empty_like(visited)
cur_o = torch.empty(R, dtype=torch.int64, device=dev)
done_o = torch.empty(R, dtype=torch.bool, device=dev)
mask_o = torch.empty(R, N, dtype=torch.bool, device=dev)


def launch():
    call("rrnco_rcvrp_step", R, N, R, ptr(action), ptr(demand), ptr(cap), R, ptr(used), ptr(visited), None, ptr(used_o),
         ptr(vis_o), ptr(cur_o), ptr(done_o), ptr(mask_o), stream_ptr(dev))


for _ in range(3):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    launch()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = R * (7 * N + 30)
print(f"cfg={os.environ.get('RRNCO_STEP_CFG', 'default')} R={R} N={N}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s "
      f"({nbytes / ms / 1e6 / 6539.5:.3f} of 6539.5)")

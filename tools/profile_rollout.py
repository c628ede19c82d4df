"""Tiny driver for ncu: one fused RCVRP n=100 rollout over `B` instances (default one full wave of 148 CTAs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from bench import host_instances, stand_in_embeddings, N_LOC, N_START  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
rb.set_precision(passes)
env = rb.RCVRPEnv(generator_params={"num_loc": N_LOC}, check_solution=False, device=dev)
torch.manual_seed(1234)
dec = rb.RRNetDecoder(env_name="rcvrp").to(dev)
raw = host_instances(B, 7)
td = env.reset(rb.TensorDictLite(raw, batch_size=[B]))
row, col = stand_in_embeddings(B, 8)
cache = dec._precompute_cache((row.to(dev), col.to(dev)))
for _ in range(reps):
    out = rb.fused_rollout(dec, cache, env, td, N_START, True, "greedy", check=False)
torch.cuda.synchronize()
print("T", out["actions"].shape[1], "mean best cost", -out["reward"].view(N_START, B).amax(0).mean().item())

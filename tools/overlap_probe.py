"""How well do two resident tiles per SM overlap?  Times one fused RCVRP n=100 rollout with 148 CTAs (one per SM), 296 (two per
SM, one wave) and 592 (two waves): t(296) / t(148) = 1 is perfect overlap, 2 is none."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from bench import host_instances, stand_in_embeddings, N_LOC, N_START  # noqa: E402

dev = torch.device("cuda", 0)
env = rb.RCVRPEnv(generator_params={"num_loc": N_LOC}, check_solution=False, device=dev)
torch.manual_seed(1234)
dec = rb.RRNetDecoder(env_name="rcvrp").to(dev)
for eng in ([2, 1] if len(sys.argv) < 2 else [int(sys.argv[1])]):
    rb._lib.set_ffn_engine(eng)
    for B in (148, 296, 592, 1184):
        raw = host_instances(B, 7)
        td = env.reset(rb.TensorDictLite(raw, batch_size=[B]))
        row, col = stand_in_embeddings(B, 8)
        cache = dec._precompute_cache((row.to(dev), col.to(dev)))
        for _ in range(2):
            out = rb.fused_rollout(dec, cache, env, td, N_START, True, "greedy", check=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = rb.fused_rollout(dec, cache, env, td, N_START, True, "greedy", check=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        T = out["actions"].shape[1]
        print(f"engine {eng} B={B:5d}: {ms:8.3f} ms per rollout, T={T}, {ms * 1e3 / T:7.2f} us per decode step, {B / ms * 1e3:8.1f} inst/s")

"""Small runs of the staged env-step kernels (RCVRP / ATSP vec pipelines, RCVRPTW shared staging), the gather and the
tiled any-N decoder for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_env_case.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from oracle import synth, model as omodel  # noqa: E402  (input generator + default-initialised weights only)
from rrnco_b200.sampler import CityOnDevice, gather_submatrix  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
# env steps in the reference layout: several blocks of 32 rollouts per warp + a tail, odd row lengths
for name, n, B, S in (("rcvrp", 100, 30, 101), ("rcvrp", 28, 9, 29), ("atsp", 100, 40, 100), ("atsp", 37, 7, 37),
                      ("rcvrptw", 30, 5, 11)):
    raw = synth.make_instances(name, B, n, seed=n, integer_demand=False)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    td = rb.batchify(env.reset(rb.TensorDictLite({k: v.to(dev) for k, v in raw.items()}, batch_size=[B])), S)
    for _ in range(3):
        a = torch.multinomial(td["action_mask"].float().cpu(), 1, generator=g).squeeze(1).to(dev)
        td.set("action", a)
        td = env.step(td)["next"]
    torch.cuda.synchronize()
    print(name, n, "env steps ok", int(td["action_mask"].sum()))
# gather (fp64 and fp32 source, with / without normalisation)
city = CityOnDevice(synth.make_city(3, 200), dev)
idx = torch.from_numpy(np.array([np.random.RandomState(i).choice(200, 41, replace=False) for i in range(9)]))
for src in (city.distance, city.distance_f32):
    a = gather_submatrix(src, idx)
    b = gather_submatrix(src, idx, normalize=True)[0]
torch.cuda.synchronize()
print("gather ok", float(a.sum()), float(b.sum()))
# tiled any-N decoder (ragged start groups), a few decode steps
for name, n, B, S in (("atsp", 150, 2, 37), ("rcvrptw", 131, 2, 9)):
    raw = synth.make_instances(name, B, n, seed=n)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    td = env.reset(rb.TensorDictLite({k: v.to(dev) for k, v in raw.items()}, batch_size=[B]))
    N = td["action_mask"].shape[-1]
    row, col = synth.random_embeddings(B, N, seed=2)
    row, col = row.to(dev), col.to(dev)

    class Enc(torch.nn.Module):
        def forward(self, td, phase=None):
            return row, col
    pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
    pol.decoder.load_state_dict(omodel.init_decoder_params(name, seed=1234))
    cache = pol.decoder._precompute_cache((row, col))
    out = rb.stepwise_rollout(pol.decoder, cache, env, td, S, True, "greedy", t_cap=None)
    torch.cuda.synchronize()
    print(name, n, "stepwise rollout ok", tuple(out["actions"].shape))

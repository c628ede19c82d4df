"""Small rollouts through the key-tiled fused kernel (CTA pairs, DSMEM exchange, cp.async bias staging), the lean kernel and
the encoder's NAB kernel for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_tiled_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from oracle import synth, model as omodel  # noqa: E402  (input generator + default-initialised weights only)

dev = torch.device("cuda", 0)
cases = (("atsp", 150, 2, "greedy", 40, 1.0), ("rcvrp", 133, 1, "sampling", 9, 1.0), ("rcvrptw", 130, 1, "greedy", 5, 1.0),
         ("atsp", 140, 1, "greedy", 6, 2.5),   # scaled embeddings: the exact-shift sweep
         ("rcvrp", 20, 3, "greedy", None, 1.0))
if os.environ.get("CASES"):
    cases = [cases[int(i)] for i in os.environ["CASES"].split(",")]
for name, n, B, kind, S, scale in cases:
    raw = synth.make_instances(name, B, n, seed=1)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    td = env.reset(rb.TensorDictLite(dict(raw), batch_size=[B]))
    N = td["action_mask"].shape[-1]
    row, col = synth.random_embeddings(B, N, seed=2)
    row, col = (scale * row).to(dev), (scale * col).to(dev)

    class Enc(torch.nn.Module):
        def forward(self, td, phase=None):
            return row, col
    pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
    pol.decoder.load_state_dict(omodel.init_decoder_params(name, seed=1234))
    cache = pol.decoder._precompute_cache((row, col))
    S = env.get_num_starts(td) if S is None else S
    t_cap = int(os.environ.get("T_CAP", 12))  # a few decode steps are enough for the sanitizer (truncation is reported, not an error)
    out = rb.fused_rollout(pol.decoder, cache, env, td, S, True, kind, check=False, t_cap=t_cap)
    torch.cuda.synchronize()
    print(name, n, "ok", tuple(out["actions"].shape), float(out["reward"].mean()), flush=True)
m = rb.DistAngleFusion(128).to(dev)
with torch.no_grad():
    o = m(torch.rand(2, 37, 2, device=dev), torch.rand(2, 37, 37, device=dev).transpose(1, 2))
torch.cuda.synchronize()
print("nab ok", float(o.mean()))
with torch.no_grad():
    y = rb.aft_nab(torch.randn(2, 37, 128, device=dev), torch.randn(2, 37, 128, device=dev), torch.randn(2, 37, 128, device=dev),
                   torch.rand(2, 37, 2, device=dev), torch.rand(2, 37, 37, device=dev), m, scale=0.7)
    md = rb.DistAngleFusion(128, use_duration_matrix=True).to(dev)
    od = md(torch.rand(3, 29, 2, device=dev), torch.rand(3, 29, 29, device=dev), torch.rand(3, 29, 29, device=dev).transpose(1, 2).contiguous().transpose(1, 2))
torch.cuda.synchronize()
print("aft_nab ok", float(y.mean()), "| duration gate ok", float(od.mean()))

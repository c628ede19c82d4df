"""Per-step drop-in API at the bench workload shape (RCVRP n=100, x8 aug, 101 starts): upstream's own decode loop
(policy.py:203-243) calling RRNetDecoder.forward + select + env.step once per step, vs the fused rollout kernel.
   python tools/per_step_probe.py [instances]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from rrnco_b200 import _lib  # noqa: E402
from oracle import synth, model as omodel  # noqa: E402  (input generator + default-initialised weights only)

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
A, S, n = 8, 101, 100
raw = synth.make_instances("rcvrp", B, n, seed=1)
env = rb.get_env("rcvrp", generator_params={"num_loc": n}, check_solution=False)
td = rb.batchify(env.reset(rb.TensorDictLite({k: v.to(dev) for k, v in raw.items()}, batch_size=[B])), A)
row, col = synth.random_embeddings(A * B, n + 1, seed=2)
row, col = row.to(dev), col.to(dev)


class Enc(torch.nn.Module):
    def forward(self, td, phase=None):
        return row, col


pol = rb.RRNetPolicy(encoder=Enc(), env_name="rcvrp").to(dev)
pol.decoder.load_state_dict(omodel.init_decoder_params("rcvrp", seed=1234))
cache = pol.decoder._precompute_cache((row, col))


def timeit(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


dt_f, out_f = timeit(lambda: rb.fused_rollout(pol.decoder, cache, env, td, S, True, "greedy", check=False))
print(f"fused rollout kernel          : {dt_f * 1e3:8.1f} ms  {B / dt_f:8.1f} instances/s")
for label, min_tiled in (("per-step, rrnco_decoder_logits (N <= 128 kernel)", 1 << 30),
                         ("per-step, rrnco_decoder_logits_large (any-N kernels)", _lib.MIN_STARTS_TILED)):
    _lib.MIN_STARTS_TILED = min_tiled
    dt, out = timeit(lambda: rb.stepwise_rollout(pol.decoder, cache, env, td, S, True, "greedy", check=False), reps=1)
    same = (out["actions"][:, :out_f["actions"].shape[1]] == out_f["actions"][:, :out["actions"].shape[1]]).all(1).float().mean().item()
    print(f"{label:52s}: {dt * 1e3:8.1f} ms  {B / dt:8.1f} instances/s  T={out['actions'].shape[1]}  tours equal to fused: {same:.4f}")

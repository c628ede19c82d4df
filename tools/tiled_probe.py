"""Key-tiled fused rollout kernel (rollout_tiled.cu, 128 < N <= 1024) against the per-step kernel pipeline on the same
inputs, and its throughput at BASELINE config C4 (ATSP n=1000, batch 64, 100 starts, greedy).
   python tools/tiled_probe.py            # agreement on small cases, then C4 timing
   CASES=0 B=64 python tools/tiled_probe.py
"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from oracle import synth, model as omodel  # noqa: E402  (input generator + default-initialised weights only)

dev = torch.device("cuda", 0)


def setup(name, n, B, seed=1):
    raw = synth.make_instances(name, B, n, seed=seed)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    td = env.reset(rb.TensorDictLite(dict(raw), batch_size=[B]))
    N = td["action_mask"].shape[-1]
    row, col = synth.random_embeddings(B, N, seed=seed + 1)
    row, col = row.to(dev), col.to(dev)

    class Enc(torch.nn.Module):
        def forward(self, td, phase=None):
            return row, col
    pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
    pol.decoder.load_state_dict(omodel.init_decoder_params(name, seed=1234))
    return env, td, pol


def compare(name, n, B, S, kind="greedy"):
    env, td, pol = setup(name, n, B, seed=n)
    S = env.get_num_starts(td) if S is None else S
    outs = {}
    for path in ("stepwise", "fused"):
        pol.large_n_path = path
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs[path] = pol(td, env, phase="val", decode_type=f"multistart_{kind}", num_starts=S, seed=7)
        torch.cuda.synchronize(); outs[path]["dt"] = time.perf_counter() - t0
    a, b = outs["stepwise"], outs["fused"]
    T = max(a["actions"].shape[1], b["actions"].shape[1])
    pa = torch.nn.functional.pad(a["actions"], (0, T - a["actions"].shape[1]))
    pb = torch.nn.functional.pad(b["actions"], (0, T - b["actions"].shape[1]))
    same = (pa == pb).all(1)
    dr = ((a["reward"] - b["reward"]).abs() / a["reward"].abs())[same].max().item() if same.any() else float("nan")
    dl = (a["log_likelihood"] - b["log_likelihood"]).abs()[same].max().item() if same.any() else float("nan")
    print(f"{name:8s} n={n:4d} B={B:3d} S={S:4d} {kind:8s}: same tours {same.float().mean().item():.4f}  max rel reward diff {dr:.2e}  "
          f"max |dLL| {dl:.2e}  T {a['actions'].shape[1]}/{b['actions'].shape[1]}  stepwise {a['dt']*1e3:.0f} ms, fused {b['dt']*1e3:.0f} ms",
          flush=True)


def c4(B, reps=5, warm=3):
    env, td, pol = setup("atsp", 1000, B, seed=1)
    for path in ("fused",) if os.environ.get("FUSED_ONLY") else ("fused", "stepwise"):
        pol.large_n_path = path
        for i in range(reps + warm):
            if i == warm:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            out = pol(td, env, phase="val", decode_type="multistart_greedy", num_starts=100)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"C4 atsp n=1000 B={B} starts=100 greedy [{path}]: {dt*1e3:9.1f} ms/batch  {B/dt:8.1f} instances/s  "
              f"mean best cost {-out['reward'].view(100, B).amax(0).mean().item():.4f}", flush=True)


if __name__ == "__main__":
    rb._lib.lib().rrnco_set_start_split(int(os.environ.get("SPLIT", 0)))
    if int(os.environ.get("CASES", 1)):
        compare("atsp", 150, 2, None)
        compare("atsp", 129, 1, 7)
        compare("rcvrp", 140, 2, None)
        compare("rcvrptw", 130, 2, 40)
        compare("atsp", 300, 2, 150)
        compare("atsp", 150, 2, None, "sampling")
        compare("rcvrp", 200, 2, 64, "sampling")
        compare("atsp", 1000, 2, 100)
    c4(int(os.environ.get("B", 64)))
    from rrnco_b200 import models
    print("fallbacks to the per-step pipeline:", models.FALLBACKS)

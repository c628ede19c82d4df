"""Throughput of the BASELINE.json configs other than the bench line (parity-test cases; reported for context)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from oracle import synth, model as omodel  # noqa: E402  (input generator + default-initialised weights only)

dev = torch.device("cuda", 0)


def run(name, n, B, A, S, kind, reps=3):
    raw = synth.make_instances(name, B, n, seed=1)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    td = env.reset(rb.TensorDictLite(dict(raw), batch_size=[B]))
    if A > 1:
        td = rb.batchify(td, A)
    N = td["action_mask"].shape[-1]
    row, col = synth.random_embeddings(A * B, N, seed=2)
    row, col = row.to(dev), col.to(dev)

    class Enc(torch.nn.Module):
        def forward(self, td, phase=None):
            return row, col
    pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
    pol.decoder.load_state_dict(omodel.init_decoder_params(name, seed=1234))
    S = env.get_num_starts(td) if S is None else S
    for i in range(reps + 1):
        if i == 1:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        out = pol(td, env, phase="val", decode_type=f"multistart_{kind}", num_starts=S)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:8s} n={n:4d} B={B:5d} aug={A} starts={S:4d} {kind:8s}: {dt*1e3:9.1f} ms/batch  {B/dt:10.1f} instances/s  "
          f"T={out['actions'].shape[1]}  mean best cost {-out['reward'].view(S, A, B).amax(0).amax(0).mean().item():.3f}")


if __name__ == "__main__":
    run("atsp", 100, 32, 8, 100, "greedy")          # C1 (reference test.py path: batch 32, x8 aug, 100 starts)
    run("rcvrp", 100, 1024, 8, 101, "greedy")       # C2 (the bench line)
    run("rcvrptw", 100, 1024, 1, 100, "sampling")   # C3
    run("atsp", 1000, 64, 1, 100, "greedy", reps=1) # C4 (key-streaming per-step path)
    run("rcvrp", 100, 4096, 1, 101, "sampling")     # C5 per-GPU share is 512; full 4096 here on one GPU

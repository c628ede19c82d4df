"""Training hand-off at a C5-like per-GPU shape (RCVRP n=100, 512 instances x 101 starts, sampling, no augmentation):
fused sampling rollout + differentiable batched replay + backward (rrnco_b200/training.py).
   python tools/train_step_probe.py [instances]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from oracle import synth, model as omodel  # noqa: E402  (input generator + default-initialised weights only)

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n, S = 100, 101
raw = synth.make_instances("rcvrp", B, n, seed=1)
env = rb.get_env("rcvrp", generator_params={"num_loc": n}, check_solution=False)
td = env.reset(rb.TensorDictLite({k: v.to(dev) for k, v in raw.items()}, batch_size=[B]))
row, col = synth.random_embeddings(B, n + 1, seed=2)
row, col = row.to(dev).requires_grad_(True), col.to(dev).requires_grad_(True)


class Enc(torch.nn.Module):
    def forward(self, td, phase=None):
        return row, col


pol = rb.RRNetPolicy(encoder=Enc(), env_name="rcvrp").to(dev)
pol.decoder.load_state_dict(omodel.init_decoder_params("rcvrp", seed=1234))


def sync_time():
    torch.cuda.synchronize()
    return time.perf_counter()


from rrnco_b200 import training as _tr  # noqa: E402
_prof = None
for it in range(int(os.environ.get("ITERS", 5))):
    if os.environ.get("PROFILE") and it == int(os.environ.get("ITERS", 5)) - 1:
        from torch.profiler import profile, ProfilerActivity
        _prof = profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU])
        _prof.__enter__()
    ac = torch.bfloat16 if it >= 3 and os.environ.get("BF16") else None
    _tr.ATTENTION_IMPL = os.environ.get("ATTN", "sdpa")
    _tr.REPLAY_IMPL = os.environ.get("IMPL", "fused")
    torch.backends.cuda.matmul.allow_tf32 = bool(int(os.environ.get("TF32", 0)))
    torch.cuda.reset_peak_memory_stats()
    t0 = sync_time()
    with torch.no_grad():
        out = pol(td, env, phase="train", decode_type="multistart_sampling", num_starts=S)
    t1 = sync_time()
    if os.environ.get("SPLIT"):   # separate timings of the env replay and the forward pass (a host sync in between)
        with torch.no_grad():
            inputs = rb.collect_decode_inputs(pol.decoder, env, td, out["actions"], S)
    else:                         # product form: the replay generator is consumed chunk by chunk, kernels overlap the host loop
        inputs = rb.iter_decode_inputs(pol.decoder, env, td, out["actions"], S)
    t2 = sync_time()
    logp = rb.batched_logprobs(pol.decoder, row, col, td["distance_matrix"].float(), None, inputs, out["actions"], S,
                              autocast_dtype=ac)
    ll = logp.sum(1)
    loss = rb.pomo_shared_baseline_loss(out["reward"], ll, S)
    t3 = sync_time()
    loss.backward()
    t4 = sync_time()
    err = (ll - out["log_likelihood"]).abs().max().item()
    print(f"iter {it} ({'bf16 autocast' if ac else 'fp32'}, {_tr.REPLAY_IMPL if not ac else 'aten'} replay): B={B} S={S} T={out['actions'].shape[1]}  sample (fused kernel) {1e3*(t1-t0):7.1f} ms | env replay "
          f"{1e3*(t2-t1):7.1f} | batched logprobs fwd {1e3*(t3-t2):7.1f} | bwd {1e3*(t4-t3):7.1f} | total {1e3*(t4-t0):7.1f} ms "
          f"= {B/(t4-t0):7.1f} instances/s | max |ll - kernel ll| {err:.1e} | peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    pol.zero_grad(set_to_none=True)
    row.grad = col.grad = None
if _prof is not None:
    _prof.__exit__(None, None, None)
    print(_prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))

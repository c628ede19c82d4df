"""Per-phase cycles of thread 0 of CTA 0 of the lean rollout kernel (needs RRNCO_PHASE_STAMPS=1 python rrnco_b200/build.py)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from bench import host_instances, stand_in_embeddings, N_LOC, N_START  # noqa: E402

NAMES = {0: "vote + step go", 1: "A: mask + query", 2: "sync (mask visible)", 3: "wait scores (4 heads)", 4: "softmax shift", 5: "exp pass + arrive P",
         6: "wait P V", 7: "glimpse write (incl. loop)", 9: "wait GEMM1 (4 chunks)", 10: "epilogue 1", 11: "wait GEMM2(3)", 12: "output epilogue",
         13: "sync", 14: "bias gather", 15: "sync", 16: "wait logits", 17: "select pass A", 18: "sync", 19: "select pass B", 20: "sync",
         21: "select pass C", 22: "sync", 23: "winner + transition", 24: "loop overhead"}
dev = torch.device("cuda", 0)
env = rb.RCVRPEnv(generator_params={"num_loc": N_LOC}, check_solution=False, device=dev)
torch.manual_seed(1234)
dec = rb.RRNetDecoder(env_name="rcvrp").to(dev)
L = rb._lib.lib()
L.rrnco_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
for B in (148, 296):
    td = env.reset(rb.TensorDictLite(host_instances(B, 7), batch_size=[B]))
    row, col = stand_in_embeddings(B, 8)
    cache = dec._precompute_cache((row.to(dev), col.to(dev)))
    rb.fused_rollout(dec, cache, env, td, N_START, True, "greedy", check=False)
    torch.cuda.synchronize()
    L.rrnco_debug_phase_cycles(None, 1)
    out = rb.fused_rollout(dec, cache, env, td, N_START, True, "greedy", check=False)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    L.rrnco_debug_phase_cycles(buf, 0)
    T = out["actions"].shape[1] - 1
    tot = sum(buf)
    print(f"--- {B} CTAs ({'one' if B == 148 else 'two'} per SM), {T} steps, {tot / T:.0f} cycles per step")
    for i in range(32):
        if buf[i]:
            print(f"  {NAMES.get(i, i):32s} {buf[i] / T:8.0f}  {100 * buf[i] / tot:5.1f}%")

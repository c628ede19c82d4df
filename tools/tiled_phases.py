"""Per-phase cycles of thread 0 of CTA 0 of the key-tiled rollout kernel (needs RRNCO_PHASE_STAMPS=1 python rrnco_b200/build.py)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from tools.tiled_probe import setup  # noqa: E402

NAMES = {0: "vote + step go", 1: "A: mask + query", 2: "sync (mask visible)", 3: "wait scores (all pairs)", 4: "exp pass + arrive P",
         5: "pair loop overhead / shift", 6: "sync + wait P V", 7: "glimpse write", 8: "wait GEMM1 (4 chunks)", 9: "epilogue 1",
         10: "wait GEMM2(3)", 11: "output epilogue", 12: "wait logits (tiles)", 13: "select pass (tiles)", 14: "bias load / exchange write",
         15: "loop overhead", 16: "sync", 17: "winner + transition", 18: "exact-shift sweep (max)"}
L = rb._lib.lib()
L.rrnco_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
n = int(os.environ.get("N", 1000))
L.rrnco_set_start_split(int(os.environ.get("SPLIT", 0)))
for B in [int(x) for x in os.environ.get("BS", "1,16,64").split(",")]:
    env, td, pol = setup("atsp", n, B, seed=1)
    row, col = pol.encoder(td)
    cache = pol.decoder._precompute_cache((row, col))
    rb.fused_rollout(pol.decoder, cache, env, td, 100, True, "greedy", check=False)  # (a loose softmax bound would raise)
    torch.cuda.synchronize()
    L.rrnco_debug_phase_cycles(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = rb.fused_rollout(pol.decoder, cache, env, td, 100, True, "greedy", check=False)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    L.rrnco_debug_phase_cycles(buf, 0)
    T = out["actions"].shape[1] - 1
    tot = sum(buf)
    print(f"--- atsp n={n}, {B} instances x {L.rrnco_rollout_tile_rows(0, n, B, 100)} starts per tile, {T} steps, {tot / T:.0f} cycles per step, {e0.elapsed_time(e1):.1f} ms per call")
    for i in range(32):
        if buf[i]:
            print(f"  {NAMES.get(i, i):32s} {buf[i] / T:8.0f}  {100 * buf[i] / tot:5.1f}%")

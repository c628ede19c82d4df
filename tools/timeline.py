"""Event timeline of one decode step of CTA 0 (development aid; needs a build with RRNCO_PHASE_STAMPS=1):
   RRNCO_PHASE_STAMPS=1 python rrnco_b200/build.py && gpurun -- python tools/timeline.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from bench import host_instances, stand_in_embeddings, N_LOC, N_START  # noqa: E402

NAMES = {1: "step start", 2: "mask+query done", 3: "mask words written", 64: "transition done", 200: "kernel start", 201: "staging done", 202: "state init done",
         203: "K/V/Lk packed", 204: "entering step loop"}
for h in range(8):
    NAMES.update({10 + h: f"grp waits scores h{h}", 20 + h: f"grp sees scores h{h}", 30 + h: f"grp wrote P h{h}",
                  110 + h: f"  issuer: QK h{h} issued", 120 + h: f"  issuer: P h{h} seen", 130 + h: f"  issuer: PV h{h} issued",
                  150 + h: f"  issuer0: FFN job {h} issued"})
NAMES.update({40: "grp0 all PV seen", 41: "grp1 all PV seen", 42: "grp0 glimpse written", 43: "grp1 glimpse written",
              58: "FFN output seen", 59: "g' written", 60: "logits seen", 63: "select passes done", 160: "  issuer0: logits issued"})
for c in range(4):
    NAMES.update({50 + c: f"hidden chunk {c} seen", 54 + c: f"epilogue-1 chunk {c} done"})
for i in range(3):
    NAMES.update({100 + i: f"  issuer{i}: Q seen", 140 + i: f"  issuer{i}: glimpse seen"})

dev = torch.device("cuda", 0)
B = 148
env = rb.RCVRPEnv(generator_params={"num_loc": N_LOC}, check_solution=False, device=dev)
torch.manual_seed(1234)
dec = rb.RRNetDecoder(env_name="rcvrp").to(dev)
td = env.reset(rb.TensorDictLite(host_instances(B, 7), batch_size=[B]))
row, col = stand_in_embeddings(B, 8)
cache = dec._precompute_cache((row.to(dev), col.to(dev)))
L = rb._lib.lib()
L.rrnco_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
buf, n = (C.c_longlong * 512)(), C.c_int(0)
for _ in range(2):
    rb.fused_rollout(dec, cache, env, td, N_START, True, "greedy", check=False)
    torch.cuda.synchronize()
    rc = L.rrnco_debug_timeline(buf, C.byref(n))
assert rc == 0, "build with RRNCO_PHASE_STAMPS=1"
ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(min(n.value, 256)))
t0 = ev[0][0]
for clk, tag in ev:
    print(f"{clk - t0:8d}  {NAMES.get(tag, tag)}")

import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rrnco_b200 import _lib
from rrnco_b200._lib import call, ptr, stream_ptr

dev = "cuda"
torch.manual_seed(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.randn(M, 128, device=dev)
lin1 = torch.nn.Linear(128, 512).to(dev)
lin2 = torch.nn.Linear(512, 128).to(dev)
w1, b1, w2, b2 = [t.detach().contiguous() for t in (lin1.weight, lin1.bias, lin2.weight, lin2.bias)]
out = torch.empty_like(g)
ws = torch.empty(_lib.lib().rrnco_pointer_ffn_workspace_bytes(), dtype=torch.uint8, device=dev)
call("rrnco_pointer_ffn", M, ptr(g), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(out), ptr(ws), stream_ptr())
torch.cuda.synchronize()
g64 = g.double()
ref = (torch.relu(g64 @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double() + g64)
ref32 = torch.relu(g @ w1.t() + b1) @ w2.t() + b2 + g
err = (out.double() - ref).abs()
print("tcgen05 ffn: max abs err vs fp64", err.max().item(), "mean", err.mean().item(), "| torch fp32 err", (ref32.double() - ref).abs().max().item())
print("rows with err>1e-3:", (err.max(1)[0] > 1e-3).sum().item(), "of", M, " first bad cols:", (err[0] > 1e-3).nonzero().flatten()[:10].tolist())
print("out[0,:6]", out[0, :6].tolist(), "\nref[0,:6]", ref[0, :6].tolist())
# timing
M2 = 148 * 128
g2 = torch.randn(M2, 128, device=dev); o2 = torch.empty_like(g2)
for _ in range(3):
    call("rrnco_pointer_ffn", M2, ptr(g2), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(o2), ptr(ws), stream_ptr())
torch.cuda.synchronize(); t0 = time.time()
for _ in range(20):
    call("rrnco_pointer_ffn", M2, ptr(g2), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(o2), ptr(ws), stream_ptr())
torch.cuda.synchronize(); dt = (time.time() - t0) / 20
print(f"one wave (148 CTAs x 128 rows): {dt*1e6:.1f} us per launch (incl. weight split kernels)")

import ctypes
_lib.lib().rrnco_debug_ffn_stamps.argtypes = [ctypes.c_void_p]
for mode in (0, 1):
    _lib.lib().rrnco_debug_ffn_mode(mode)
    call("rrnco_pointer_ffn", M2, ptr(g2), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(o2), ptr(ws), stream_ptr())
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    _lib.lib().rrnco_debug_ffn_stamps(buf)
    st = list(buf); t0 = st[0]
    print(f"mode {mode}: setup {st[1]-t0}; per-half issue_done/acc_ready:", [(st[2+3*i]-t0, st[3+3*i]-t0) for i in range(3)], "end", st[26]-t0)
_lib.lib().rrnco_debug_ffn_mode(0)

"""Standalone check of the tcgen05 FFN kernel built alone into tools/_scratch/libffn_probe.so (development aid):
   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static \
        rrnco_b200/csrc/ffn_tc_kernel.cu -o tools/_scratch/libffn_probe.so"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = C.CDLL(os.path.join(ROOT, "tools/_scratch/libffn_probe.so"))
L.rrnco_pointer_ffn_workspace_bytes.restype = C.c_int64
L.rrnco_pointer_ffn.restype = C.c_int
L.rrnco_pointer_ffn.argtypes = [C.c_int64] + [C.c_void_p] * 8
L.rrnco_debug_ffn_stamps.argtypes = [C.c_void_p]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
E, F = 128, 512
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
w1 = (torch.rand(F, E, device=dev) * 2 - 1) / E ** 0.5
w2 = (torch.rand(E, F, device=dev) * 2 - 1) / F ** 0.5
b1 = (torch.rand(F, device=dev) * 2 - 1) / E ** 0.5
b2 = (torch.rand(E, device=dev) * 2 - 1) / F ** 0.5
ws = torch.empty(L.rrnco_pointer_ffn_workspace_bytes(), dtype=torch.uint8, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for M in (128, 300, 128 * 148 * 4):
    g = torch.randn(M, E, device=dev) * scale
    out = torch.empty_like(g)
    rc = L.rrnco_pointer_ffn(M, p(g), p(w1), p(b1), p(w2), p(b2), p(out), p(ws), st)
    torch.cuda.synchronize()
    assert rc == 0, rc
    ref = (torch.relu(g.double() @ w1.double().T + b1.double()) @ w2.double().T + b2.double() + g.double())
    ref32 = torch.relu(g @ w1.T + b1) @ w2.T + b2 + g
    err = (out.double() - ref).abs()
    print(f"M={M}: max abs err vs fp64 {err.max().item():.3e} mean {err.mean().item():.3e} | torch fp32 err "
          f"{(ref32.double() - ref).abs().max().item():.3e} | |out| max {ref.abs().max().item():.2f}")
M = 128 * 148 * 8
g = torch.randn(M, E, device=dev)
out = torch.empty_like(g)
for _ in range(3):
    L.rrnco_pointer_ffn(M, p(g), p(w1), p(b1), p(w2), p(b2), p(out), p(ws), st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.rrnco_pointer_ffn(M, p(g), p(w1), p(b1), p(w2), p(b2), p(out), p(ws), st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"M={M}: {ms:.3f} ms per call, {ms * 1e3 / 8:.1f} us per 128-row tile-wave, {2 * 2 * M * E * F / ms / 1e9:.1f} algorithmic TFLOP/s")
buf = (C.c_longlong * 32)()
L.rrnco_debug_ffn_stamps(buf)
s = list(buf)
print("stamps (cycles from kernel start): setup", s[1] - s[0], "| epi starts/ends", [x - s[0] for x in s[2:10]], "| out", s[10] - s[0], "end", s[26] - s[0])
L.rrnco_debug_ffn_mode.argtypes = [C.c_int]
for mode, M2 in ((1, 128), (0, 128)):
    L.rrnco_debug_ffn_mode(mode)
    g2, o2 = g[:M2].contiguous(), out[:M2].contiguous()
    for _ in range(3):
        L.rrnco_pointer_ffn(M2, p(g2), p(w1), p(b1), p(w2), p(b2), p(o2), p(ws), st)
    torch.cuda.synchronize()
    L.rrnco_debug_ffn_stamps(buf)
    s = list(buf)
    print(f"mode {mode} M={M2}: setup", s[1] - s[0], "| epi starts/ends", [x - s[0] for x in s[2:10]], "| out", s[10] - s[0], "end", s[26] - s[0], "| producer issued 4/8/12/16 at", [x - s[0] for x in s[12:16]])
L.rrnco_debug_ffn_mode(0)

"""Times the training hand-off kernels (rrnco_b200/train_ops.py) on their own at the C5 per-GPU shape, one 32-step chunk:
512 instances x (32 steps x 101 starts) rows, 101 nodes.   python tools/train_kernels_probe.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rrnco_b200 import train_ops  # noqa: E402

dev = torch.device("cuda", 0)
n_inst, L, N = 512, 32 * 101, 101
rows = n_inst * L
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps=3):
    if os.environ.get("ONCE"):   # one launch of each kernel (for an ncu capture)
        fn()
        torch.cuda.synchronize()
        return float("nan")
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x = torch.randn(rows, 128, device=dev, generator=g)
dy = torch.randn(rows, 128, device=dev, generator=g) * 1e-6
w1 = (torch.rand(512, 128, device=dev, generator=g) * 2 - 1) / 128 ** 0.5
w2 = (torch.rand(128, 512, device=dev, generator=g) * 2 - 1) / 512 ** 0.5
b1, b2 = torch.zeros(512, device=dev), torch.zeros(128, device=dev)
h = train_ops.lib()
st = train_ops.status_word(dev)
P, S = train_ops._p, train_ops._stream
packed = train_ops._pack(w1, w2, dev)
packed_t = train_ops._pack(w2.t().contiguous(), w1.t().contiguous(), dev)
mask = torch.empty(rows, 16, dtype=torch.int32, device=dev)
y = torch.empty_like(x)
hid = torch.empty(rows, 512, device=dev)
dhid = torch.empty(rows, 512, device=dev)
sdy = train_ops.pow2_scale(dy)
c = torch.zeros(512, 128, device=dev)
cs = torch.zeros(512, device=dev)
flop = 2 * 2 * rows * 128 * 512
t = timed(lambda: h.rrnco_train_ffn(0, rows, P(x), P(packed), P(b1), P(b2), None, P(mask), None, P(y), P(st), S(dev)))
print(f"ffn forward (mask out)          {t:7.2f} ms  {flop / t / 1e9:7.1f} TFLOP/s algorithmic   rows {rows}")
t = timed(lambda: h.rrnco_train_ffn(0, rows, P(x), P(packed), P(b1), P(b2), None, None, P(hid), None, P(st), S(dev)))
print(f"ffn forward (hidden out)        {t:7.2f} ms")
t = timed(lambda: h.rrnco_train_ffn(1, rows, P(dy), P(packed_t), None, None, P(sdy), P(mask), P(dhid), P(y), P(st), S(dev)))
print(f"ffn backward-data (dhidden out) {t:7.2f} ms")
t = timed(lambda: h.rrnco_train_xty(rows, P(hid), 1, P(x), None, None, P(c), P(cs), None, P(st), S(dev)))
print(f"xty [rows,512]^T [rows,128]     {t:7.2f} ms  {2 * rows * 512 * 128 / t / 1e9:7.1f} TFLOP/s algorithmic, {rows * 640 * 4 / t / 1e6:7.1f} GB/s read")
ffn = lambda: train_ops.fused_ffn(x.requires_grad_(True), w1.requires_grad_(True), b1.requires_grad_(True), w2.requires_grad_(True), b2.requires_grad_(True))
def fb():
    out = ffn()
    out.backward(dy)
t = timed(fb)
print(f"fused_ffn forward + backward (autograd Function, incl. packing, scales, recompute) {t:7.2f} ms")
del hid, dhid, y, mask
q = torch.randn(n_inst, L, 128, device=dev, generator=g)
k = torch.randn(n_inst, N, 128, device=dev, generator=g)
v = torch.randn(n_inst, N, 128, device=dev, generator=g)
m = torch.rand(n_inst, L, N, device=dev, generator=g) < 0.5
m[..., 0] = True
out = torch.empty_like(q)
lse = torch.empty(n_inst, L, 8, device=dev)
m8 = m.view(torch.uint8)
t = timed(lambda: h.rrnco_train_attention_fwd(n_inst, L, N, P(q), P(k), P(v), P(m8), 1, P(out), P(lse), S(dev)))
print(f"attention forward               {t:7.2f} ms")
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
t = timed(lambda: h.rrnco_train_attention_bwd(n_inst, L, N, P(q), P(k), P(v), P(m8), P(out), 1, P(lse), P(dy.view_as(q)), P(dq), P(dk), P(dv), S(dev)))
print(f"attention backward              {t:7.2f} ms")
lk = torch.randn(n_inst, N, 128, device=dev, generator=g)
pk = torch.empty(h.rrnco_train_inst_packed_bytes(n_inst), dtype=torch.uint8, device=dev)
h.rrnco_train_inst_pack(n_inst, N, P(lk), 0, P(pk), P(st), S(dev))
zz = torch.empty(n_inst, L, 128, device=dev)
t = timed(lambda: h.rrnco_train_inst_gemm(n_inst, L, P(q), P(pk), None, None, P(zz), P(st), S(dev)))
print(f"pointer scores (inst_gemm)      {t:7.2f} ms  {2 * rows * 128 * 128 / t / 1e9:7.1f} TFLOP/s on the padded tile")
dlk = torch.zeros(n_inst, N, 128, device=dev)
t = timed(lambda: h.rrnco_train_inst_xty(n_inst, L, N, P(zz), P(q), None, None, None, P(dlk), P(st), S(dev)))
print(f"pointer dLk (xty_inst)          {t:7.2f} ms")
z = torch.randn(n_inst, L, N, device=dev, generator=g) * 20
dist = torch.rand(n_inst, N, N, device=dev, generator=g)
cur = torch.randint(0, N, (n_inst, L), device=dev, generator=g)
act = torch.zeros(n_inst, L, dtype=torch.int64, device=dev)
alpha = torch.ones(1, device=dev)
lp, da = torch.empty(n_inst, L, device=dev), torch.empty(n_inst, L, device=dev)
t = timed(lambda: h.rrnco_train_logits_tail(rows, L, N, N, P(z), P(dist), None, P(cur), P(m8), P(act), P(alpha), None, 128 ** -0.5, 10.0, 1.0, P(lp), P(da), None, S(dev)))
print(f"logits tail                     {t:7.2f} ms  {rows * N * 9 / t / 1e6:7.1f} GB/s (z in/out + mask)")
train_ops.check_status(dev)

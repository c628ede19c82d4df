"""Small forward + backward passes through every kernel of librrnco_b200_train.so for compute-sanitizer (memcheck / racecheck /
synccheck):   compute-sanitizer --tool racecheck python tools/sanitize_train_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rrnco_b200 import train_ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
n_inst, L, N = 2, 75, 37


def leaf(*shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(dev).requires_grad_(True)


# residual FFN (tcgen05 chain, relu mask, X^T Y weight gradients): 3 tiles with a ragged tail
x = leaf(300, 128, scale=2.0)
w1, b1, w2, b2 = leaf(512, 128, scale=0.08), leaf(512, scale=0.08), leaf(128, 512, scale=0.04), leaf(128, scale=0.04)
y = train_ops.fused_ffn(x, w1, b1, w2, b2)
(y * 1e-6).sum().backward()
print("ffn ok", float(y.sum()), float(w1.grad.abs().max()))
# attention
q, k, v = leaf(n_inst, L, 128), leaf(n_inst, N, 128), leaf(n_inst, N, 128)
mask = (torch.rand(n_inst, L, N, generator=g) < 0.5).to(dev)
mask[..., 0] = True
o = train_ops.fused_attention(q, k, v, mask)
(o * 1e-5).sum().backward()
print("attention ok", float(o.sum()), float(k.grad.abs().max()))
# context query
ta = leaf(n_inst, N, 128)
ia = torch.randint(0, N, (n_inst, L), generator=g).to(dev)
st = torch.rand(n_inst, L, 1, generator=g).to(dev)
sw = leaf(128, 1)
qq = train_ops.context_query(ta, ia, None, None, st, sw.t())
qq.sum().backward()
print("context ok", float(qq.sum()))
# pointer scores + tail (one node) and the separate forms
gg, lk = leaf(n_inst, L, 128), leaf(n_inst, N, 128)
dist = torch.rand(n_inst, N, N, generator=g).to(dev)
alpha = torch.ones(1, device=dev, requires_grad=True)
cur = torch.randint(0, N, (n_inst, L), generator=g).to(dev)
act = torch.randint(0, N, (n_inst, L), generator=g).to(dev)
mask.scatter_(-1, act.unsqueeze(-1), True)
lp = train_ops.pointer_logprob(gg, lk, alpha, None, dist, None, cur, mask, act, 10.0, 1.0)
(lp * 1e-4).sum().backward()
z = train_ops.pointer_scores(gg, lk)
lp2 = train_ops.fused_logits_tail(z, alpha, None, dist, None, cur, mask, act, 10.0, 1.0)
(lp2 * 1e-4).sum().backward()
torch.cuda.synchronize()
train_ops.check_status(dev)
print("pointer ok", float(lp.sum()), float((lp - lp2).abs().max()))

"""Encoder hot spot: DistAngleFusion / AFTFull of the attention-free block.

CPU: the oracle restatement reproduces the fixture recorded from the UNMODIFIED reference module
(tests/golden/encoder_nab.npz <- tests/golden/make_golden.py encoder), and the algebraic collapse the CUDA kernel evaluates
(four E-vectors per module) is the same function.  GPU: the kernel against the fixture and against the oracle at n = 100.
"""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oenc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "encoder_nab.npz")


def _fixture():
    z = np.load(GOLDEN)
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    params = {tag: {k[len(tag) + 7:]: v for k, v in t.items() if k.startswith(tag + ".param.")} for tag in ("nodur", "dur", "aft")}
    params["aft"] = {k[10:]: v for k, v in t.items() if k.startswith("aft.param.")}
    return t, params


def test_oracle_matches_reference_fixture():
    t, params = _fixture()
    for tag in ("nodur", "dur"):
        dur = t["dur"] if tag == "dur" else None
        out = oenc.dist_angle_fusion(params[tag], t["coords"], t["cost"], dur)
        assert torch.equal(out, t[tag + ".bias"])  # same torch ops in the same order: bit-exact on the CPU
        outT = oenc.dist_angle_fusion(params[tag], t["coords"], t["cost"].transpose(1, 2), None if dur is None else dur.transpose(1, 2))
        assert torch.equal(outT, t[tag + ".bias_T"])
    out = oenc.aft_full(params["aft"], t["aft.x"], t["aft.y"], t["nodur.bias"] * 1.7)
    assert torch.equal(out, t["aft.out"])


def collapsed_bias(p, coords, cost):
    """What rrnco_nab_gating computes, in fp64 torch: the module collapsed to four E-vectors (encoder_kernels.cu header)."""
    d = {k: v.double() for k, v in p.items()}
    E = d["out_lin.weight"].shape[1]
    wg, wo = d["gate.0.weight"][0], d["out_lin.weight"][0]
    ug_d, uo_d = d["dist_emb.2.weight"].t() @ wg[:E], d["dist_emb.2.weight"].t() @ wo
    ug_a, uo_a = d["angle_emb.2.weight"].t() @ wg[E:], d["angle_emb.2.weight"].t() @ wo
    cg = d["gate.0.bias"][0] + wg[:E] @ d["dist_emb.2.bias"] + wg[E:] @ d["angle_emb.2.bias"]
    hd = torch.relu(cost.double().unsqueeze(-1) * d["dist_emb.0.weight"][:, 0] + d["dist_emb.0.bias"])
    ha = torch.relu(oenc.pairwise_angles(coords).double().unsqueeze(-1) * d["angle_emb.0.weight"][:, 0] + d["angle_emb.0.bias"])
    g = torch.sigmoid(hd @ ug_d + ha @ ug_a + cg)
    return g * (hd @ uo_d + wo @ d["dist_emb.2.bias"]) + (1 - g) * (ha @ uo_a + wo @ d["angle_emb.2.bias"]) + d["out_lin.bias"][0]


def test_collapsed_form_is_the_same_function():
    t, params = _fixture()
    got = collapsed_bias(params["nodur"], t["coords"], t["cost"])
    assert (got - t["nodur.bias"].double()).abs().max() < 2e-5  # fp32 reference vs fp64 collapse (values up to 11)


def test_mirror_module_has_the_reference_parameter_names():
    import rrnco_b200 as rb
    _, params = _fixture()
    m = rb.DistAngleFusion(128)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in params["nodur"].items()}
    m.load_state_dict(params["nodur"], strict=True)
    with pytest.raises(NotImplementedError):
        rb.DistAngleFusion(128, use_duration_matrix=True)


@pytest.mark.gpu
def test_nab_kernel_matches_reference_fixture_and_oracle():
    import rrnco_b200 as rb
    dev = "cuda"
    t, params = _fixture()
    m = rb.DistAngleFusion(128).to(dev)
    m.load_state_dict(params["nodur"], strict=True)
    coords, cost = t["coords"].to(dev), t["cost"].to(dev)
    with torch.no_grad():
        for variant in (0, 1):  # 0 = piecewise-linear segment tables (default), 1 = sum over the hidden units
            out = m(coords, cost, variant=variant)
            outT = m(coords, cost.transpose(1, 2), variant=variant)  # the col-encoding block's transposed view: read through its base
            out2 = m(coords, cost, scale=0.37, variant=variant)
            assert (out.cpu() - t["nodur.bias"]).abs().max() < 2e-5, variant
            assert (outT.cpu() - t["nodur.bias_T"]).abs().max() < 2e-5, variant
            assert torch.allclose(out2, out * 0.37, rtol=1e-6, atol=1e-7)
    with pytest.raises(NotImplementedError):  # forward-only
        m(coords, cost)
    # n = 100 customers + depot, default initialisation (what a freshly built encoder holds), ragged tail of the pair chunks
    torch.manual_seed(3)
    m2 = rb.DistAngleFusion(128).to(dev)
    p2 = {k: v.detach().cpu() for k, v in m2.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    coords = torch.rand(5, 101, 2, generator=g)
    cost = torch.rand(5, 101, 101, generator=g) * 1.4
    want = oenc.dist_angle_fusion(p2, coords, cost)
    with torch.no_grad():
        got = m2(coords.to(dev), cost.to(dev))
        got1 = m2(coords.to(dev), cost.to(dev), variant=1)
        # arguments exactly ON breakpoints and far outside their range (every segment incl. the two unbounded ones)
        w, bb = m2.dist_emb[0].weight[:, 0], m2.dist_emb[0].bias
        edge = torch.cat([-bb / w, torch.tensor([-1e6, -50.0, 0.0, 50.0, 1e6], device=dev)])
        ce = edge[torch.randint(0, edge.numel(), (2, 101, 101), device=dev, generator=torch.Generator(device=dev).manual_seed(5))]
        e0, e1 = m2(coords[:2].to(dev), ce, variant=0), m2(coords[:2].to(dev), ce, variant=1)
    assert (got.cpu() - want).abs().max() < 2e-6 * max(1.0, want.abs().max().item())
    assert (got1.cpu() - want).abs().max() < 2e-6 * max(1.0, want.abs().max().item())
    assert torch.isfinite(e0).all() and ((e0 - e1).abs() <= 1e-5 * e1.abs() + 1e-5).all()
    # a parameter update invalidates the packed cache
    with torch.no_grad():
        m2.out_lin.bias.add_(1.0)
        got2 = m2(coords.to(dev), cost.to(dev))
    assert (got2 - got - 1.0).abs().max() < 1e-5


@pytest.mark.gpu
def test_patch_encoder_swaps_the_gating_modules():
    import rrnco_b200 as rb
    from torch import nn

    class RefFusion(nn.Module):  # structure of the reference module (attn_freenet.py:201-240), torch forward = the oracle's
        def __init__(self):
            super().__init__()
            self.embed_dim = 128
            self.dist_emb = nn.Sequential(nn.Linear(1, 128), nn.ReLU(), nn.Linear(128, 128))
            self.angle_emb = nn.Sequential(nn.Linear(1, 128), nn.ReLU(), nn.Linear(128, 128))
            self.gate = nn.Sequential(nn.Linear(256, 1), nn.Sigmoid())
            self.out_lin = nn.Linear(128, 1)

        def forward(self, coords, cost_mat, duration_mat=None):
            return oenc.dist_angle_fusion({k: v for k, v in self.state_dict().items()}, coords, cost_mat)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.angle_distance_fusion = RefFusion()

    enc = nn.ModuleList([Block(), Block()]).to("cuda")
    g = torch.Generator().manual_seed(8)
    coords, cost = torch.rand(2, 30, 2, generator=g).cuda(), torch.rand(2, 30, 30, generator=g).cuda()
    with torch.no_grad():
        want = [b.angle_distance_fusion(coords, cost.transpose(1, 2)) for b in enc]
        assert rb.patch_encoder(enc) == 2 and rb.patch_encoder(enc) == 0
        got = [b.angle_distance_fusion(coords, cost.transpose(1, 2)) for b in enc]
    for a, b in zip(got, want):
        assert (a - b).abs().max() < 5e-6

"""Parity statistics at the size the 99.9 % bar is meaningful on (development aid; the asserting form lives in
tests/test_gpu_parity.py): greedy multistart rollouts of B instances at n=100, CUDA path vs the fp32 oracle, with the
fp32-vs-fp64 oracle disagreement (the noise floor of the discontinuous argmax) measured on the same inputs, and the
run-to-run flip rate of the CUDA path on identical inputs.

    gpurun -- python tools/parity_1024.py [B] [chunk]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb  # noqa: E402
from oracle import envs as oenvs, model as omodel, synth  # noqa: E402
from oracle.td import TD  # noqa: E402

dev = "cuda"


def lite(td):
    return rb.TensorDictLite({k: v.to(dev) for k, v in td.items()}, batch_size=list(td.batch_size))


def make_policy(name, p, row, col):
    class Enc(torch.nn.Module):
        def forward(self, td, phase=None):
            return row, col
    pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
    pol.decoder.load_state_dict(p, strict=True)
    return pol


def chunked_oracle(p, oenv, raw, row, col, S, chunk, dtype=torch.float32):
    best, acts = [], []
    B = raw.batch_size[0]
    pp = omodel.cast_params(p, dtype) if dtype != torch.float32 else p
    for i in range(0, B, chunk):
        sub = TD({k: v[i:i + chunk] for k, v in raw.items()}, batch_size=[min(chunk, B - i)])
        otd = oenv.reset(sub)
        o = omodel.policy_forward(pp, oenv, otd, row[i:i + chunk].to(dtype), col[i:i + chunk].to(dtype),
                                  decode_type="multistart_greedy", num_starts=S)
        b = sub.batch_size[0]
        best.append(o["reward"].float().view(S, b).max(0)[0])
        acts.append(o["actions"].view(S, b, -1))
    return torch.cat(best), acts


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    n = 100
    torch.set_num_threads(os.cpu_count() or 1)
    for name in ("rcvrp", "atsp", "rcvrptw"):
        raw = synth.make_instances(name, B, n, seed=2025)
        oenv = oenvs.make_env(name, n, check_solution=False)
        S = oenv.get_num_starts(oenv.reset(TD({k: v[:2] for k, v in raw.items()}, batch_size=[2])))
        N = n if name == "atsp" else n + 1
        row, col = synth.random_embeddings(B, N, seed=77)
        p = omodel.init_decoder_params(name, seed=1234)
        t0 = time.time()
        obest, oacts = chunked_oracle(p, oenv, raw, row, col, S, chunk)
        t32 = time.time() - t0
        B64 = min(B, 256)
        raw64 = TD({k: v[:B64] for k, v in raw.items()}, batch_size=[B64])
        t0 = time.time()
        o64best, o64acts = chunked_oracle(p, oenv, raw64, row[:B64], col[:B64], S, chunk, torch.float64)
        t64 = time.time() - t0
        env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
        pol = make_policy(name, p, row.to(dev), col.to(dev))
        outs = []
        for rep in range(2):
            out = pol(env.reset(lite(raw)), env, phase="val", decode_type="multistart_greedy", num_starts=S)
            outs.append({k: out[k].cpu() for k in ("actions", "reward", "log_likelihood")})
        best = outs[0]["reward"].view(S, B).max(0)[0]
        rel = (best - obest).abs() / obest.abs()
        rel64 = (obest[:B64] - o64best).abs() / o64best.abs()
        # rollout-level tour agreement
        ga = outs[0]["actions"].view(S, B, -1)
        same, total, same64, total64 = 0, 0, 0, 0
        for ci, oa in enumerate(oacts):
            b = oa.shape[1]
            g = ga[:, ci * chunk: ci * chunk + b]
            T = max(g.shape[-1], oa.shape[-1])
            g = torch.nn.functional.pad(g, (0, T - g.shape[-1]))
            o = torch.nn.functional.pad(oa, (0, T - oa.shape[-1]))
            same += (g == o).all(-1).sum().item()
            total += S * b
            if ci * chunk < B64:
                o6 = o64acts[ci]
                T = max(o6.shape[-1], oa.shape[-1])
                same64 += (torch.nn.functional.pad(o6, (0, T - o6.shape[-1])) == torch.nn.functional.pad(oa, (0, T - oa.shape[-1]))).all(-1).sum().item()
                total64 += S * b
        rr_same = torch.equal(outs[0]["actions"], outs[1]["actions"]) if outs[0]["actions"].shape == outs[1]["actions"].shape else False
        rr_bits = torch.equal(outs[0]["reward"], outs[1]["reward"]) and torch.equal(outs[0]["log_likelihood"], outs[1]["log_likelihood"])
        print(f"{name}: B={B} S={S}  instances within 1e-4 of the fp32 oracle: {(rel < 1e-4).float().mean():.5f} "
              f"(max rel {rel.max():.2e}); rollouts with identical tours {same / total:.5f} | noise floor on {B64} instances "
              f"(fp32 oracle vs fp64 oracle): instances within 1e-4 {(rel64 < 1e-4).float().mean():.5f}, identical tours "
              f"{same64 / total64:.5f} | run-to-run: tours identical {rr_same}, reward+loglik bitwise {rr_bits} | "
              f"oracle {t32:.1f} s fp32, {t64:.1f} s fp64 ({B64})", flush=True)


if __name__ == "__main__":
    main()

"""GPU diagnostic: runs every kernel path against the oracle / golden fixtures and PRINTS the deviations
(no asserts), so that one gpurun call surfaces as many problems as possible.  Not part of the product."""
import glob
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rrnco_b200 as rb  # noqa: E402
from oracle import envs as oenvs, model as omodel, synth  # noqa: E402
from oracle.td import TD, batchify as obatchify  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
dev = "cuda"


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].shape != () else z[k].item() for k in z.files}


def to_lite(td, device=dev):
    return rb.TensorDictLite({k: v.to(device) for k, v in td.items()}, batch_size=list(td.batch_size))


def section(title):
    print(f"\n=== {title} ===", flush=True)


def guarded(fn):
    def wrap(*a, **k):
        try:
            return fn(*a, **k)
        except Exception:
            traceback.print_exc()
            torch.cuda.synchronize()
    return wrap


@guarded
def diag_env_golden():
    section("env forced sequences vs golden (reference code)")
    for path in sorted(glob.glob(os.path.join(GOLDEN, "env_*.npz"))):
        fname = os.path.basename(path)
        z = load(fname)
        name = fname[4:-4].split("_")[0]
        raw = TD({k[3:]: v for k, v in z.items() if k.startswith("in.")}, batch_size=[z["actions"].shape[0]])
        n = raw["distance_matrix"].shape[-1] - (0 if name == "atsp" else 1)
        env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
        td = env.reset(to_lite(raw, "cpu"))
        bad = []
        if not torch.equal(td["action_mask"].cpu(), z["reset.action_mask"]):
            bad.append("reset.mask")
        if not torch.equal(td["distance_matrix"].cpu(), z["reset.distance_matrix"]):
            bad.append(f"reset.dm maxdiff={(td['distance_matrix'].cpu() - z['reset.distance_matrix']).abs().max():.3g}")
        acts = z["actions"].to(dev)
        for t in range(acts.shape[1]):
            td.set("action", acts[:, t])
            td = env.step(td)["next"]
            for k in [k for k in z if k.startswith("step.")]:
                got, want = td[k[5:]].cpu(), z[k][t]
                if got.shape != want.shape or got.dtype != want.dtype:
                    bad.append(f"{k}@{t} shape/dtype {tuple(got.shape)}/{got.dtype} vs {tuple(want.shape)}/{want.dtype}")
                elif not torch.equal(got, want):
                    bad.append(f"{k}@{t} ndiff={(got != want).sum().item()}")
        real, norm = env.get_reward(td, acts)
        e1 = ((real.cpu() - z["reward.real"]).abs() / z["reward.real"].abs()).max().item()
        e2 = ((norm.cpu() - z["reward.norm"]).abs() / z["reward.norm"].abs()).max().item()
        print(f"{fname:28s} mismatches={len(bad)} {bad[:4]} reward relerr real={e1:.2e} norm={e2:.2e}")


def build_policy(name, p, row, col):
    class Enc(torch.nn.Module):
        def forward(self, td, phase=None):
            return row, col
    pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
    missing = pol.decoder.load_state_dict({k: v for k, v in p.items()}, strict=True)
    return pol


@guarded
def diag_policy_golden():
    section("decoder logits / greedy rollout / evaluate vs golden (reference code)")
    for name in ["atsp", "rcvrp", "rcvrptw"]:
        z = load(f"policy_{name}.npz")
        B = z["row_emb"].shape[0]
        raw = TD({k[3:]: v for k, v in z.items() if k.startswith("in.")}, batch_size=[B])
        p = {k[6:]: v for k, v in z.items() if k.startswith("param.")}
        n = raw["distance_matrix"].shape[-1] - (0 if name == "atsp" else 1)
        S = z["num_starts"]
        env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
        row, col = z["row_emb"].to(dev), z["col_emb"].to(dev)
        pol = build_policy(name, p, row, col)
        td0 = env.reset(to_lite(raw))
        # per-step decoder at the mid-rollout state
        td = rb.batchify(env.reset(to_lite(raw)), S)
        for t in range(4):
            td.set("action", z["greedy.actions"][:, t].to(dev))
            td = env.step(td)["next"]
        cache = pol.decoder._precompute_cache((row, col))
        ok = omodel.precompute_cache(p, z["row_emb"], z["col_emb"])
        for a, b_ in [("glimpse_key", cache.glimpse_key), ("glimpse_val", cache.glimpse_val), ("logit_key", cache.logit_key)]:
            print(f"  {name} cache {a} maxabs err {(b_.cpu() - ok[a]).abs().max():.2e}")
        logits, mask = pol.decoder(td, cache, S)
        fin = torch.isfinite(z["mid.logits"])
        err = (logits.cpu() - z["mid.logits"])[fin].abs().max().item()
        print(f"  {name} mid logits maxabs err {err:.3e} mask equal {torch.equal(mask.cpu(), z['mid.mask'])}")
        for passes in (3, 1):
            rb.set_precision(passes)
            out = pol(td0.clone(), env, phase="val", decode_type="multistart_greedy", num_starts=S)
            acts = out["actions"].cpu()
            want = z["greedy.actions"]
            same_shape = acts.shape == want.shape
            match = (acts == want).all(1).float().mean().item() if same_shape else -1
            rerr = ((out["reward"].cpu() - z["greedy.reward"]).abs() / z["greedy.reward"].abs()).max().item()
            lerr = (out["log_likelihood"].cpu() - z["greedy.log_likelihood"]).abs().max().item()
            print(f"  {name} passes={passes} greedy: shape {tuple(acts.shape)} vs {tuple(want.shape)} rollout match "
                  f"{match:.4f} reward relerr {rerr:.2e} ll abserr {lerr:.2e}")
        rb.set_precision(3)
        # evaluate: replay the golden actions through the fused kernel
        out2 = pol(td0.clone(), env, phase="val", num_starts=S, actions=z["greedy.actions"][:, 1:].to(dev))
        lerr = (out2["log_likelihood"].cpu() - z["evaluate.log_likelihood"]).abs().max().item()
        rerr = ((out2["reward"].cpu() - z["evaluate.reward"]).abs() / z["evaluate.reward"].abs()).max().item()
        print(f"  {name} evaluate: ll abserr {lerr:.2e} reward relerr {rerr:.2e} actions equal "
              f"{torch.equal(out2['actions'].cpu(), z['greedy.actions'])}")


@guarded
def diag_rollout_vs_oracle(name, B, n, seed=3):
    section(f"fused greedy rollout vs oracle: {name} n={n} B={B}")
    raw = synth.make_instances(name, B, n, seed=seed)
    oenv = oenvs.make_env(name, n, check_solution=False)
    otd = oenv.reset(raw)
    N = otd["action_mask"].shape[-1]
    S = oenv.get_num_starts(otd)
    row, col = synth.random_embeddings(B, N, seed=seed + 1)
    p = omodel.init_decoder_params(name, seed=seed)
    t0 = time.time()
    oout = omodel.policy_forward(p, oenv, otd, row, col, decode_type="multistart_greedy", num_starts=S)
    t_cpu = time.time() - t0
    p64 = omodel.cast_params(p, torch.float64)
    otd64 = oenv.reset(raw)
    o64 = omodel.policy_forward(p64, oenv, otd64, row.double(), col.double(), decode_type="multistart_greedy", num_starts=S)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    pol = build_policy(name, p, row.to(dev), col.to(dev))
    td0 = env.reset(to_lite(raw))
    for passes in (3, 1):
        rb.set_precision(passes)
        torch.cuda.synchronize()
        t0 = time.time()
        out = pol(td0.clone(), env, phase="val", decode_type="multistart_greedy", num_starts=S)
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        acts, want = out["actions"].cpu(), oout["actions"]
        T = min(acts.shape[1], want.shape[1])
        match = (acts[:, :T] == want[:, :T]).all(1).float().mean().item()
        cost = -out["reward"].cpu().view(S, B).max(0)[0] * -1
        best = out["reward"].cpu().view(S, B).max(0)[0]
        obest = oout["reward"].view(S, B).max(0)[0]
        o64best = o64["reward"].float().view(S, B).max(0)[0]
        rel = ((best - obest).abs() / obest.abs())
        rel64 = ((obest - o64best).abs() / o64best.abs())
        m64 = (oout["actions"][:, :min(T, o64['actions'].shape[1])] == o64["actions"][:, :min(T, o64['actions'].shape[1])]).all(1).float().mean().item()
        print(f"  passes={passes} T gpu/cpu {acts.shape[1]}/{want.shape[1]} rollout match {match:.4f} "
              f"(noise floor fp32-vs-fp64 oracle {m64:.4f}); inst best-cost within 1e-4: {(rel < 1e-4).float().mean():.4f} "
              f"(floor {(rel64 < 1e-4).float().mean():.4f}); mean cost gpu {-best.mean():.4f} cpu {-obest.mean():.4f}; "
              f"time gpu {t_gpu*1e3:.1f} ms cpu {t_cpu:.1f} s")
    rb.set_precision(3)


@guarded
def diag_gather():
    section("gather / normalise")
    city = synth.make_city(3, 300)
    rng = np.random.RandomState(0)
    idx = np.array([rng.choice(300, 101, replace=False) for _ in range(64)])
    want = city["distance"][idx[:, :, None], idx[:, None, :]].astype(np.float32)
    from rrnco_b200.sampler import CityOnDevice, gather_submatrix
    c = CityOnDevice(city)
    got = gather_submatrix(c.distance, torch.from_numpy(idx))
    print("  gather equal:", np.array_equal(got.cpu().numpy(), want))
    got_n, mn, mx = gather_submatrix(c.distance, torch.from_numpy(idx), normalize=True)
    w = torch.from_numpy(want)
    lo, hi = w.amin((1, 2), keepdim=True), w.amax((1, 2), keepdim=True)
    print("  fused normalise equal:", torch.equal(got_n.cpu(), (w - lo) / (hi - lo + 1e-6)),
          torch.equal(mn.cpu(), lo.flatten()), torch.equal(mx.cpu(), hi.flatten()))


@guarded
def diag_speed(name="rcvrp", B=256, n=100, A=1):
    section(f"speed probe {name} n={n} B={B}")
    raw = synth.make_instances(name, B, n, seed=5)
    env = rb.get_env(name, generator_params={"num_loc": n}, check_solution=False)
    td0 = env.reset(to_lite(raw))
    N = td0["action_mask"].shape[-1]
    S = env.get_num_starts(td0)
    row, col = synth.random_embeddings(B, N, seed=6)
    p = omodel.init_decoder_params(name, seed=5)
    pol = build_policy(name, p, row.to(dev), col.to(dev))
    for passes in (3, 1):
        rb.set_precision(passes)
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            out = pol(td0, env, phase="val", decode_type="multistart_greedy", num_starts=S)
            torch.cuda.synchronize()
            dt = time.time() - t0
        T = out["actions"].shape[1]
        import ctypes
        buf = (ctypes.c_longlong * 16)()
        L = rb._lib.lib()
        L.rrnco_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.rrnco_debug_phase_cycles(None, 1)
        out = pol(td0, env, phase="val", decode_type="multistart_greedy", num_starts=S)
        torch.cuda.synchronize()
        L.rrnco_debug_phase_cycles(buf, 0)
        cyc = [c / max(T - 1, 1) for c in list(buf)[:16]]
        names = ["q+sync", "attention", "ffn:convert", "ffn:gemm+epi1", "ffn:out-epi", "logits gemm", "transition+tail",
                 "sel:tmem+bias+clip", "sel:softmax-sum", "sel:argmax", "sel:chosen",
                 "att:wait-scores", "att:max-pass", "att:exp-pass", "att:wait-PV", "top+mask"]
        print("   cycles/step (CTA 0): " + ", ".join(f"{n}={c:.0f}" for n, c in zip(names, cyc)) + f"  total={sum(cyc):.0f}")
        print(f"  passes={passes}: {dt*1e3:.1f} ms for {B} instances x {S} starts, T={T} -> {B/dt:.0f} inst/s; "
              f"per CTA-step {dt/ (B*T/148) *1e6:.1f} us")
    rb.set_precision(3)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "abi", rb._lib.lib().rrnco_abi_version())
    quick = len(sys.argv) > 1 and sys.argv[1] in ("quick", "speed")
    if not quick:
        diag_env_golden()
        diag_gather()
    speed_only = len(sys.argv) > 1 and sys.argv[1] == "speed"
    for engine in (1, 0):
        rb.set_ffn_engine(engine)
        print(f"\n######## FFN engine {engine} ({'tcgen05' if engine else 'mma.sync'}) ########")
        if not speed_only:
            diag_policy_golden()
            for name in ["rcvrp", "atsp", "rcvrptw"]:
                diag_rollout_vs_oracle(name, 4, 20)
            diag_rollout_vs_oracle("rcvrp", 4, 100)
        diag_speed("rcvrp", 296, 100)

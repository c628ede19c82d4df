"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run here (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden.py

It puts `oracle/shims` (stand-ins for rl4co / tensordict / torchrl, see oracle/shims/README.md) and
`/root/reference` on sys.path, imports the reference's own modules and records what THEY compute:

  env_<name>.npz      reset + forced random feasible action sequence: per-step masks / state / done,
                      final (real, normalised) reward           -> rrnco/envs/*/env.py
  decoder_<name>.npz  RRNetDecoder.forward logits + mask at a mid-rollout state, with the state_dict
                      and the embeddings that produced them     -> rrnco/models/decoder.py
  policy_<name>.npz   RRNetPolicy.forward multistart-greedy rollout (encoder replaced by fixed
                      embeddings): actions, reward, log-likelihood -> rrnco/models/policy.py, decoding.py
  sampler.npz         Real_World_Sampler.sample on a small synthetic city -> rrnco/envs/*/sampler.py
  augment.npz         StateAugmentation(dihedral8)              -> rrnco/models/utils/transforms.py
  generator_*.npz     Lazy{RCVRP,ATSP,RMTVRP}Generator._process_real_world_data / subsample_problems with the
                      uniform draws they consumed               -> rrnco/envs/*/generator_lazy.py, rmtvrp/generator.py
  sampler_outliers.npz Real_World_Sampler.sample on a city with > 1e5 entries -> rrnco/envs/rmtvrp/sampler.py:41-60
  encoder_nab.npz     DistAngleFusion (both gate variants) + AFTFull of the encoder's attention-free block
                                                                -> rrnco/models/nn/attn_freenet.py:201-327

Inputs are produced with oracle.synth (only as an input generator, nothing of the oracle's
arithmetic is recorded).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference", ROOT]

from tensordict import TensorDict  # noqa: E402  (shim)

from rrnco.envs.atsp.env import ATSPEnv  # noqa: E402
from rrnco.envs.rcvrp.env import RCVRPEnv  # noqa: E402
from rrnco.envs.rmtvrp.env import RMTVRPEnv  # noqa: E402
from rrnco.envs.rcvrp.sampler import Real_World_Sampler as SamplerC  # noqa: E402
from rrnco.envs.rmtvrp.sampler import Real_World_Sampler as SamplerTW  # noqa: E402
from rrnco.models.decoder import RRNetDecoder  # noqa: E402
from rrnco.models.policy import RRNetPolicy  # noqa: E402
from rrnco.models.utils.transforms import StateAugmentation  # noqa: E402
from rl4co.utils.ops import batchify  # noqa: E402

from oracle import synth  # noqa: E402


def to_ref_td(td):
    return TensorDict({k: v.clone() for k, v in td.items()}, batch_size=list(td.batch_size))


def make_env(name, n):
    gp = {"num_loc": n}
    if name == "atsp":
        return ATSPEnv(generator_params=gp, check_solution=True)
    if name == "rcvrp":
        return RCVRPEnv(generator_params=gp, check_solution=True)
    return RMTVRPEnv(generator_params=gp, check_solution=False)


STATE_KEYS = {
    "atsp": ["action_mask", "current_node", "first_node", "i", "done"],
    "rcvrp": ["action_mask", "current_node", "used_capacity", "visited", "done"],
    "rcvrptw": ["action_mask", "current_node", "current_time", "current_route_length",
                "used_capacity_linehaul", "used_capacity_backhaul", "visited", "done"],
}


def variant_inputs(td, variant, g):
    """RMTVRP O / L / B / MB variants on top of the TW instance (rmtvrp/env.py:225-262)."""
    B, n = td["demand_linehaul"].shape
    if "O" in variant:
        td["open_route"] = torch.ones(B, 1, dtype=torch.bool)
    if "L" in variant:
        td["distance_limit"] = torch.full((B, 1), 2.8)  # in normalised-distance units
    if "B" in variant:
        is_b = torch.rand(B, n, generator=g) < 0.3
        dl = td["demand_linehaul"]
        td["demand_backhaul"] = dl * is_b
        td["demand_linehaul"] = dl * ~is_b
        td["backhaul_class"] = torch.full((B, 1), 2.0 if "M" in variant else 1.0)
    if "noTW" in variant:
        del td["time_windows"], td["service_time"]
    return td


def gen_env(name, n, B, seed, variant=""):
    g = torch.Generator().manual_seed(seed)
    raw = synth.make_instances(name, B, n, seed=seed, integer_demand=(seed % 2 == 0))
    if variant and name == "rcvrptw":
        raw = variant_inputs(raw, variant, g)
    env = make_env(name, n)
    td = env.reset(to_ref_td(raw))
    rec = {f"in.{k}": v.numpy() for k, v in raw.items()}
    rec["reset.action_mask"] = td["action_mask"].numpy()
    rec["reset.distance_matrix"] = td["distance_matrix"].clone().numpy()
    rec["reset.min_distance"] = td["min_distance"].numpy()
    rec["reset.max_distance"] = td["max_distance"].numpy()
    steps, actions = {k: [] for k in STATE_KEYS[name]}, []
    t = 0
    while not td["done"].all():
        a = torch.multinomial(td["action_mask"].float(), 1, generator=g).squeeze(1)
        td.set("action", a)
        td = env.step(td)["next"]
        actions.append(a)
        for k in STATE_KEYS[name]:
            steps[k].append(td[k].clone())
        t += 1
        assert t < 10 * n
    actions = torch.stack(actions, 1)
    rec["actions"] = actions.numpy()
    for k, v in steps.items():
        rec[f"step.{k}"] = torch.stack(v, 0).numpy()
    real, norm = env.get_reward(td, actions)
    rec["reward.real"], rec["reward.norm"] = real.numpy(), norm.numpy()
    rec["after_reward.distance_matrix"] = td["distance_matrix"].clone().numpy()  # rmtvrp/env.py:433 mutates col 0
    np.savez_compressed(os.path.join(HERE, f"env_{name}{('_' + variant) if variant else ''}.npz"), **rec)
    print("env", name, variant, "T =", actions.shape[1])


class FixedEncoder(torch.nn.Module):
    def __init__(self, row, col):
        super().__init__()
        self.row, self.col = row, col

    def forward(self, td, phase=None):
        return self.row, self.col


def gen_decoder_and_policy(name, n, B, seed):
    torch.manual_seed(seed)
    raw = synth.make_instances(name, B, n, seed=seed)
    env = make_env(name, n)
    td0 = env.reset(to_ref_td(raw))
    N = td0["action_mask"].shape[-1]
    row, col = synth.random_embeddings(B, N, seed=seed + 1)
    dec = RRNetDecoder(embed_dim=128, num_heads=8, env_name=name)
    with torch.no_grad():  # non-trivial alpha / beta so that the bias scale is exercised
        dec.alpha.fill_(0.8)
        if name == "rcvrptw":
            dec.beta.fill_(1.3)
    policy = RRNetPolicy(encoder=FixedEncoder(row, col), decoder=dec, env_name=name).eval()
    S = env.get_num_starts(td0)
    rec = {f"in.{k}": v.numpy() for k, v in raw.items()}
    rec.update({f"param.{k}": v.detach().numpy() for k, v in dec.state_dict().items()})
    rec["row_emb"], rec["col_emb"] = row.numpy(), col.numpy()
    rec["num_starts"] = np.int64(S)
    with torch.inference_mode():
        out = policy(td0.clone(), env, phase="val", decode_type="multistart_greedy", num_starts=S,
                     return_actions=True)
        rec["greedy.actions"] = out["actions"].numpy()
        rec["greedy.reward"] = out["reward"].numpy()
        rec["greedy.normalized_reward"] = out["normalized_reward"].numpy()
        rec["greedy.log_likelihood"] = out["log_likelihood"].numpy()

        # one decoder call at a mid-rollout state (after the forced start + 3 greedy steps)
        td = batchify(env.reset(to_ref_td(raw)), S)
        acts = out["actions"]
        for t in range(4):
            td.set("action", acts[:, t])
            td = env.step(td)["next"]
        _, _, cache = dec.pre_decoder_hook(td, env, (row, col), S)
        logits, mask = dec(td, cache, S)
        rec["mid.logits"], rec["mid.mask"] = logits.numpy(), mask.numpy()

        # evaluate path: replay with given actions (decoding.py:386-399), flat (no multistart) td
        td_flat = batchify(env.reset(to_ref_td(raw)), S)
        pol2 = RRNetPolicy(encoder=FixedEncoder(batchify(row, S), batchify(col, S)), decoder=dec,
                           env_name=name).eval()
        out2 = pol2(td_flat, env, phase="val", actions=acts, return_actions=True)
        rec["evaluate.log_likelihood"] = out2["log_likelihood"].numpy()
        rec["evaluate.reward"] = out2["reward"].numpy()
    np.savez_compressed(os.path.join(HERE, f"policy_{name}.npz"), **rec)
    print("policy", name, "actions", tuple(out["actions"].shape), "mean cost", -out["reward"].mean().item())


def gen_sampler():
    city = synth.make_city(7, length=60)
    rec = {f"city.{k}": v for k, v in city.items()}
    np.random.seed(4321)
    s = SamplerC().sample(city, batch=5, num_sample=11)
    rec["c.points"], rec["c.distance_matrix"] = s["points"], s["distance_matrix"]
    np.random.seed(4321)
    s = SamplerTW().sample(city, batch=5, num_sample=11)
    rec["tw.points"], rec["tw.distance_matrix"], rec["tw.duration_matrix"] = (
        s["points"], s["distance_matrix"], s["duration_matrix"])
    np.random.seed(4321)
    rec["indices"] = SamplerC().uniform_sample(5, 60, 11)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **rec)
    print("sampler ok")


def gen_augment():
    g = torch.Generator().manual_seed(5)
    td = TensorDict({"locs": torch.rand(3, 6, 2, generator=g), "distance_matrix": torch.rand(3, 6, 6, generator=g)},
                    batch_size=[3])
    aug = StateAugmentation(num_augment=8, augment_fn="dihedral8", first_aug_identity=True, no_aug_coords=False)
    out = aug(td)
    np.savez_compressed(os.path.join(HERE, "augment.npz"), locs_in=td["locs"].numpy(),
                        dm_in=td["distance_matrix"].numpy(), locs_out=out["locs"].numpy(),
                        dm_out=out["distance_matrix"].numpy())
    print("augment ok")


def gen_generators():
    """Lazy*Generator._process_real_world_data (+ subsample_problems) of the UNMODIFIED reference on sub-matrices drawn by
    the reference's own sampler.  The uniform draws the laws consume are recorded by replaying the same torch RNG calls
    after the same seed (rand / uniform_ share one stream), so that a re-implementation can be checked bit for bit."""
    from rrnco.envs.rcvrp.generator_lazy import LazyRCVRPGenerator
    from rrnco.envs.rmtvrp.generator_lazy import LazyRMTVRPGenerator
    from rrnco.envs.atsp.generator_lazy import LazyATSPGenerator
    city = synth.make_city(11, length=80)
    city["duration"][3] = city["duration"][3, 0]  # (rows with a constant duration exist in OSRM tables)
    B, n = 6, 12

    def save(name, rec):
        np.savez_compressed(os.path.join(HERE, name), **rec)
        print(name, "ok")

    # rcvrp
    np.random.seed(77)
    chunk = SamplerC().sample(city, batch=B, num_sample=n + 1)
    gen = LazyRCVRPGenerator(num_loc=n)
    torch.manual_seed(123)
    td = gen._process_real_world_data(dict(chunk), [B])
    torch.manual_seed(123)
    rec = {"in.points": chunk["points"], "in.distance_matrix": chunk["distance_matrix"], "draw.demand": torch.rand(B, n).numpy(),
           "capacity_value": np.float32(gen.capacity)}
    rec.update({f"out.{k}": td[k].numpy() for k in td.keys()})
    save("generator_rcvrp.npz", rec)

    # atsp
    np.random.seed(78)
    chunk = SamplerC().sample(city, batch=B, num_sample=n)
    td = LazyATSPGenerator(num_loc=n)._process_real_world_data(dict(chunk), [B])
    rec = {"in.points": chunk["points"], "in.distance_matrix": chunk["distance_matrix"]}
    rec.update({f"out.{k}": td[k].numpy() for k in td.keys()})
    save("generator_atsp.npz", rec)

    # rcvrptw (config preset "vrptw") and the all-features preset; one instance gets a constant duration matrix
    for preset in ("vrptw", "ovrpbltw"):
        np.random.seed(79)
        chunk = SamplerTW().sample(city, batch=B, num_sample=n + 1)
        chunk["duration_matrix"][2] = 7.5  # zero range -> the np.where guard
        gen = LazyRMTVRPGenerator(num_loc=n, variant_preset=preset)
        torch.manual_seed(321)
        td = gen._process_real_world_data({k: v.copy() for k, v in chunk.items()}, [B])
        torch.manual_seed(321)
        rec = {"in.points": chunk["points"], "in.distance_matrix": chunk["distance_matrix"],
               "in.duration_matrix": chunk["duration_matrix"], "capacity_value": np.float32(gen.capacity)}
        for key, shape in (("linehaul", (B, n)), ("backhaul", (B, n)), ("is_linehaul", (B, n)), ("service_time", (B, n)),
                           ("tw_length", (B, n)), ("tw_start", (B, n)), ("distance_limit", (B,))):
            rec["draw." + key] = torch.rand(*shape).numpy()
        rec.update({f"out.{k}": td[k].numpy() for k in td.keys()})
        sub = gen.subsample_problems(TensorDict({k: v.clone() for k, v in td.items()}, batch_size=[B]))
        rec.update({f"sub.{k}": sub[k].numpy() for k in sub.keys()})
        save(f"generator_rcvrptw_{preset}.npz", rec)


def gen_outliers():
    """Real_World_Sampler.sample on a city with unreachable pairs (> 1e5): the row / column removal of sampler.py:41-60."""
    city = synth.make_city(12, length=40)
    for r, c in ((5, 9), (5, 17), (5, 30), (22, 9)):
        city["distance"][r, c] = 2.0e5
        city["duration"][r, c] = 2.0e5
    rec = {f"city.{k}": v for k, v in city.items()}
    np.random.seed(99)
    s = SamplerTW().sample(city, batch=4, num_sample=9)
    rec["points"], rec["distance_matrix"], rec["duration_matrix"] = s["points"], s["distance_matrix"], s["duration_matrix"]
    np.savez_compressed(os.path.join(HERE, "sampler_outliers.npz"), **rec)
    print("sampler outliers ok", s["distance_matrix"].max())


def gen_encoder_nab():
    """DistAngleFusion (gating neural adaptive bias, both gate variants), the block's `alpha` scaling and AFTFull of the
    UNMODIFIED reference (rrnco/models/nn/attn_freenet.py:201-327, 417-432) on a small instance batch: parameters
    (state_dict), inputs and outputs.  The col-encoding block sees the TRANSPOSED cost matrix with the same coords
    (attn_freenet.py:480-486): recorded too."""
    from rrnco.models.nn.attn_freenet import AFTFull, DistAngleFusion
    torch.manual_seed(77)
    B, N, E = 3, 13, 128
    g = torch.Generator().manual_seed(78)
    coords = torch.rand(B, N, 2, generator=g)
    cost = torch.rand(B, N, N, generator=g) * (1 - torch.eye(N))
    dur = torch.rand(B, N, N, generator=g) * (1 - torch.eye(N))
    rec = {"coords": coords.numpy(), "cost": cost.numpy(), "dur": dur.numpy()}
    with torch.no_grad():
        for tag, use_dur in (("nodur", False), ("dur", True)):
            m = DistAngleFusion(E, use_duration_matrix=use_dur)
            for q in m.parameters():  # default init gives a nearly constant bias: widen it so that every term matters
                q.mul_(3.0)
            for k, v in m.state_dict().items():
                rec[f"{tag}.param.{k}"] = v.numpy()
            rec[f"{tag}.bias"] = m(coords, cost, dur if use_dur else None).numpy()
            rec[f"{tag}.bias_T"] = m(coords, cost.transpose(1, 2), dur.transpose(1, 2) if use_dur else None).numpy()
        aft = AFTFull(dim=E, hidden_dim=E)
        for k, v in aft.state_dict().items():
            rec[f"aft.param.{k}"] = v.numpy()
        x, y = torch.randn(B, N, E, generator=g), torch.randn(B, N, E, generator=g)
        rec["aft.x"], rec["aft.y"] = x.numpy(), y.numpy()
        rec["aft.out"] = aft(x, y=y, adapt_bias=torch.from_numpy(rec["nodur.bias"]) * 1.7).numpy()
    np.savez_compressed(os.path.join(HERE, "encoder_nab.npz"), **rec)
    print("encoder_nab", {k: v.shape for k, v in rec.items() if not k.startswith(("nodur.param", "dur.param", "aft.param"))})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":  # fixture of the encoder hot spot (round 2, second half)
        gen_encoder_nab()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "new":  # fixtures added in round 2 (the others regenerate bit-identically)
        gen_generators()
        gen_outliers()
        sys.exit(0)
    gen_env("atsp", 9, 6, 10)
    gen_env("rcvrp", 12, 6, 11)   # continuous demand law
    gen_env("rcvrp", 12, 6, 12, variant="int")  # integer demand law (seed even)
    gen_env("rcvrptw", 12, 6, 13)
    for v in ["O", "L", "B", "MB", "OLB", "noTW"]:
        gen_env("rcvrptw", 12, 6, 14, variant=v)
    for name in ["atsp", "rcvrp", "rcvrptw"]:
        gen_decoder_and_policy(name, 10, 3, 20)
    gen_sampler()
    gen_augment()
    gen_generators()
    gen_outliers()
    gen_encoder_nab()

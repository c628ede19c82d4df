"""Oracle (oracle/) vs golden fixtures recorded from the UNMODIFIED reference (tests/golden/make_golden.py).

Integer / bool / index quantities must be bit-exact; fp32 state identical too on CPU (same torch ops in
the same order), rewards and logits to 1e-6.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import envs, model, sampler
from oracle.td import TD, batchify

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].shape != () else z[k].item() for k in z.files}


def _inputs(z):
    d = {k[3:]: v for k, v in z.items() if k.startswith("in.")}
    B = next(iter(d.values())).shape[0]
    return TD(d, batch_size=[B])


ENV_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "env_*.npz")))


@pytest.mark.parametrize("fname", ENV_FILES)
def test_env_forced_sequence(fname):
    z = _load(fname)
    name = fname[4:-4].split("_")[0]
    raw = _inputs(z)
    n = raw["distance_matrix"].shape[-1] - (0 if name == "atsp" else 1)
    env = envs.make_env(name, n, check_solution=(name != "rcvrptw"))
    td = env.reset(raw)
    assert torch.equal(td["action_mask"], z["reset.action_mask"])
    assert torch.equal(td["distance_matrix"], z["reset.distance_matrix"])
    assert torch.equal(td["min_distance"], z["reset.min_distance"])
    acts = z["actions"]
    for t in range(acts.shape[1]):
        td["action"] = acts[:, t]
        td = env.step(td)["next"]
        for k in [k for k in z if k.startswith("step.")]:
            got, want = td[k[5:]], z[k][t]
            assert got.shape == want.shape, (k, t, got.shape, want.shape)
            assert torch.equal(got, want), (k, t)
    assert td["done"].all()
    real, norm = env.get_reward(td, acts)
    torch.testing.assert_close(real, z["reward.real"], rtol=1e-6, atol=0)
    torch.testing.assert_close(norm, z["reward.norm"], rtol=1e-6, atol=0)
    assert torch.equal(td["distance_matrix"], z["after_reward.distance_matrix"])  # open-route column-0 quirk


@pytest.mark.parametrize("name", ["atsp", "rcvrp", "rcvrptw"])
def test_policy_greedy_and_decoder(name):
    z = _load(f"policy_{name}.npz")
    raw = _inputs(z)
    p = {k[6:]: v for k, v in z.items() if k.startswith("param.")}
    n = raw["distance_matrix"].shape[-1] - (0 if name == "atsp" else 1)
    env = envs.make_env(name, n, check_solution=(name != "rcvrptw"))
    S = z["num_starts"]
    row, col = z["row_emb"], z["col_emb"]
    td0 = env.reset(raw)
    assert env.get_num_starts(td0) == S
    out = model.policy_forward(p, env, td0.clone(), row, col, decode_type="multistart_greedy", num_starts=S)
    assert torch.equal(out["actions"], z["greedy.actions"])
    torch.testing.assert_close(out["reward"], z["greedy.reward"], rtol=1e-6, atol=0)
    torch.testing.assert_close(out["normalized_reward"], z["greedy.normalized_reward"], rtol=1e-6, atol=0)
    torch.testing.assert_close(out["log_likelihood"], z["greedy.log_likelihood"], rtol=1e-5, atol=1e-6)

    # decoder logits at the mid-rollout state
    td = batchify(env.reset(raw), S)
    for t in range(4):
        td["action"] = z["greedy.actions"][:, t]
        td = env.step(td)["next"]
    cache = model.precompute_cache(p, row, col)
    logits, mask = model.decoder_forward(p, name, td, cache, S)
    assert torch.equal(mask, z["mid.mask"])
    torch.testing.assert_close(logits, z["mid.logits"], rtol=1e-5, atol=1e-6)

    # evaluate path on a flat (already batchified) td
    td_flat = batchify(env.reset(raw), S)
    out2 = model.policy_forward(p, env, td_flat, batchify(row, S), batchify(col, S),
                                actions=z["greedy.actions"])
    torch.testing.assert_close(out2["log_likelihood"], z["evaluate.log_likelihood"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out2["reward"], z["evaluate.reward"], rtol=1e-6, atol=0)


def test_sampler_matches_reference():
    z = np.load(os.path.join(GOLDEN, "sampler.npz"))
    city = {k[5:]: z[k] for k in z.files if k.startswith("city.")}
    np.random.seed(4321)
    idx = sampler.uniform_sample(5, 60, 11)
    assert np.array_equal(idx, z["indices"])
    np.random.seed(4321)
    s = sampler.sample(city, 5, 11)
    assert np.array_equal(s["distance_matrix"], z["c.distance_matrix"])
    assert np.array_equal(s["points"], z["c.points"])
    np.random.seed(4321)
    s = sampler.sample(city, 5, 11, with_duration=True)
    assert np.array_equal(s["duration_matrix"], z["tw.duration_matrix"])
    assert s["distance_matrix"].dtype == np.float64


def test_sampler_errors():
    city = {"points": np.zeros((5, 2)), "distance": np.zeros((5, 5)), "duration": np.zeros((5, 5))}
    with pytest.raises(ValueError):
        sampler.sample(city, 0, 3)
    with pytest.raises(ValueError):
        sampler.sample(city, 2, 6)

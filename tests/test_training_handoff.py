"""Training hand-off (rrnco_b200/training.py): the batched, differentiable teacher-forced decoder replay against the
oracle's per-step decoder (decoder.py:151-206,281-326 + decoding.py:311-361), values and gradients, on the CPU."""
import pytest
import torch

from oracle import envs as oenvs, model as omodel, synth
from oracle.td import batchify as obatchify


def oracle_rollout_and_inputs(name, n, B, S, seed):
    raw = synth.make_instances(name, B, n, seed=seed)
    oenv = oenvs.make_env(name, n, check_solution=False)
    otd = oenv.reset(raw)
    N = otd["action_mask"].shape[-1]
    row, col = synth.random_embeddings(B, N, seed=seed + 1)
    p = omodel.init_decoder_params(name, seed=seed + 2)
    g = torch.Generator().manual_seed(seed)
    out = omodel.policy_forward(p, oenv, otd, row, col, decode_type="multistart_sampling", num_starts=S, generator=g)
    acts = out["actions"]
    # replay on the oracle env, recording what the decoder saw at every call
    td = obatchify(oenv.reset(raw), S)
    td["action"] = acts[:, 0]
    td = oenv.step(td)["next"]
    rec = {"current_node": [], "action_mask": [], "first_node": [], "ctx_state": []}
    for t in range(1, acts.shape[1]):
        rec["current_node"].append(td["current_node"].reshape(-1).clone())
        rec["action_mask"].append(td["action_mask"].clone())
        if name == "atsp":
            rec["first_node"].append(td["first_node"].reshape(-1).clone())
        else:
            rec["ctx_state"].append(omodel._state_embedding(name, td).clone())
        td["action"] = acts[:, t]
        td = oenv.step(td)["next"]
    inputs = {k: torch.stack(v) for k, v in rec.items() if v}
    return raw, otd, row, col, p, out, inputs


@pytest.mark.parametrize("name,n,B,S", [("rcvrp", 12, 3, 5), ("atsp", 9, 2, 9), ("rcvrptw", 10, 3, 4)])
def test_batched_replay_matches_oracle_stepwise_logprobs_and_grads(name, n, B, S):
    import rrnco_b200 as rb
    raw, otd, row, col, p, out, inputs = oracle_rollout_and_inputs(name, n, B, S, seed=11)
    dec = rb.RRNetDecoder(env_name=name)
    dec.load_state_dict(p, strict=True)
    row_g, col_g = row.clone().requires_grad_(True), col.clone().requires_grad_(True)
    dur = otd["duration_matrix"] if name == "rcvrptw" else None
    logp = rb.batched_logprobs(dec, row_g, col_g, otd["distance_matrix"], dur, inputs, out["actions"], S, step_chunk=3)
    want = out["logprobs"][:, 1:]
    assert logp.shape == want.shape
    assert (logp - want).abs().max().item() < 2e-5
    assert (logp.sum(1) - out["log_likelihood"]).abs().max().item() < 1e-4
    # gradients: the same scalar through the oracle's sequential evaluate-mode loop with autograd
    loss = rb.pomo_shared_baseline_loss(out["reward"], logp.sum(1), S)
    loss.backward()
    p_g = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    row_o, col_o = row.clone().requires_grad_(True), col.clone().requires_grad_(True)
    oenv = oenvs.make_env(name, n, check_solution=False)
    o2 = omodel.policy_forward(p_g, oenv, oenv.reset(raw), row_o, col_o, num_starts=S, actions=out["actions"][:, 1:])
    r = out["reward"].view(S, B)
    adv = (r - r.mean(0, keepdim=True)).reshape(-1)
    (-(adv * o2["log_likelihood"]).mean()).backward()
    pairs = [(dec.pointer.ffn.lins[0].weight.grad, p_g["pointer.ffn.lins.0.weight"].grad),
             (dec.project_node_embeddings.weight.grad, p_g["project_node_embeddings.weight"].grad),
             (dec.context_embedding.project_context.weight.grad, p_g["context_embedding.project_context.weight"].grad),
             (dec.alpha.grad, p_g["alpha"].grad), (row_g.grad, row_o.grad), (col_g.grad, col_o.grad)]
    for got, ref in pairs:
        assert got is not None and ref is not None
        assert (got - ref).abs().max().item() <= 1e-4 * max(ref.abs().max().item(), 1e-3), (got - ref).abs().max()


def test_shared_baseline_loss_is_zero_mean_advantage():
    import rrnco_b200 as rb
    S, B = 4, 3
    reward = torch.randn(S * B)
    ll = torch.randn(S * B, requires_grad=True)
    loss = rb.pomo_shared_baseline_loss(reward, ll, S)
    loss.backward()
    # d loss / d ll = -(advantage) / (S B); advantages of one instance sum to zero over its starts
    assert torch.allclose(ll.grad.view(S, B).sum(0), torch.zeros(B), atol=1e-6)
    r = reward.view(S, B)
    assert torch.allclose(ll.grad, -((r - r.mean(0, keepdim=True)).reshape(-1)) / (S * B), atol=1e-7)

"""CPU oracle of the RRNet decoder, the decoding strategies and the policy loop
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows:
  rrnco/models/decoder.py:123-232 (RRNetDecoder), :281-326 (RRNet_PointerAttention)
  rrnco/models/env_embeddings/context.py:7-70 + rl4co VRPContext / TSPContext (App. A)
  rrnco/models/decoding.py:157-399 (hooks, process_logits, greedy / sampling / evaluate)
  rrnco/models/policy.py:175-255 (decode loop, reward, log-likelihood)
Parameters are a flat dict keyed like RRNetDecoder.state_dict().
`dtype=torch.float64` gives the fp64 twin used to measure the fp32 noise floor.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .td import TD, batchify, gather_by_index, unbatchify

CTX_EXTRA = {"rcvrp": 1, "rcvrptw": 4}  # state scalars appended to the current-node embedding


def init_decoder_params(env_name: str, embed_dim: int = 128, seed: int = 1234) -> dict:
    """Default-initialised decoder parameters (torch nn.Linear init, fixed creation order)."""
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(int(torch.randint(0, 2**31 - 1, (1,), generator=g)))
    E = embed_dim
    ctx_in = 2 * E if env_name == "atsp" else E + CTX_EXTRA[env_name]
    mods = {
        "context_embedding.project_context": nn.Linear(ctx_in, E, bias=False),
        "project_node_embeddings": nn.Linear(E, 3 * E, bias=False),
        "project_fixed_context": nn.Linear(E, E, bias=False),  # unused (graph_context = 0)
        "pointer.project_out": nn.Linear(E, E, bias=False),  # unused (decoder.py:295)
        "pointer.ffn.lins.0": nn.Linear(E, 4 * E),
        "pointer.ffn.lins.1": nn.Linear(4 * E, E),
    }
    p = {}
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            p[f"{name}.{k}"] = v.detach().clone()
    if env_name == "atsp":
        p["context_embedding.W_placeholder"] = torch.empty(2 * E).uniform_(-1, 1)
    p["alpha"] = torch.tensor([1.0])
    if env_name == "rcvrptw":
        p["beta"] = torch.tensor([1.0])
    torch.random.set_rng_state(state)
    return p


def cast_params(p: dict, dtype) -> dict:
    return {k: v.to(dtype) for k, v in p.items()}


# ---------------------------------------------------------------------------
# decoder
# ---------------------------------------------------------------------------

def precompute_cache(p: dict, row_emb, col_emb) -> dict:
    """decoder.py:214-232: K, V, Lk = Linear(E->3E)(col_emb).chunk(3); node_embeddings = row_emb."""
    k, v, lk = F.linear(col_emb, p["project_node_embeddings.weight"]).chunk(3, dim=-1)
    return {"node_embeddings": row_emb, "graph_context": 0, "glimpse_key": k, "glimpse_val": v,
            "logit_key": lk}


def _state_embedding(env_name: str, td):
    if env_name == "rcvrp":  # rl4co VRPContext
        return td["vehicle_capacity"] - td["used_capacity"]
    # context.py:51-70 (MTVRPContextEmbedding)
    used = torch.where(td["used_capacity_backhaul"] == 0, td["used_capacity_linehaul"],
                       td["used_capacity_backhaul"])
    remaining = torch.nan_to_num(td["distance_limit"] - td["current_route_length"], posinf=10)
    return torch.cat((td["vehicle_capacity"] - used, td["current_time"], td["open_route"].float(),
                      remaining), -1)


def context_embedding(p: dict, env_name: str, emb, td):
    W = p["context_embedding.project_context.weight"]
    if env_name == "atsp":  # rl4co TSPContext
        bsz = emb.size(0)
        node_dim = (-1,) if td["first_node"].dim() == 1 else (td["first_node"].size(-1), -1)
        if int(td["i"].flatten()[0]) < 1:
            ph = p["context_embedding.W_placeholder"]
            ctx = ph[None, :].expand(bsz, ph.size(-1)) if len(td.batch_size) < 2 else \
                ph[None, None, :].expand(bsz, td.batch_size[1], ph.size(-1))
        else:
            idx = torch.stack([td["first_node"], td["current_node"]], -1).view(bsz, -1)
            ctx = gather_by_index(emb, idx).view(bsz, *node_dim)
        return F.linear(ctx.to(W.dtype), W)
    cur = gather_by_index(emb, td["current_node"])  # context.py:18-21
    state = _state_embedding(env_name, td).to(cur.dtype)
    return F.linear(torch.cat([cur, state], -1), W)  # context.py:27-31


def pointer_attention(p: dict, q, k, v, lk, mask, num_heads: int = 8):
    """decoder.py:281-326: SDPA(H heads, bool mask) -> +q -> FFN + residual -> g.Lk^T / sqrt(E)."""
    def heads(x):  # "... g (h s) -> ... h g s"
        return x.unflatten(-1, (num_heads, -1)).transpose(-2, -3)

    am = mask.unsqueeze(1) if mask.ndim == 3 else mask.unsqueeze(1).unsqueeze(2)
    h = F.scaled_dot_product_attention(heads(q), heads(k), heads(v), attn_mask=am)
    h = h.transpose(-2, -3).flatten(-2)  # "... h n g -> ... n (h g)"
    g = h + q
    f = F.linear(F.relu(F.linear(g, p["pointer.ffn.lins.0.weight"], p["pointer.ffn.lins.0.bias"])),
                 p["pointer.ffn.lins.1.weight"], p["pointer.ffn.lins.1.bias"])
    g = f + g
    logits = torch.bmm(g, lk.transpose(-2, -1)).squeeze(-2) / math.sqrt(g.size(-1))
    assert not torch.isnan(logits).any(), "Logits contain NaNs"  # decoder.py:303-304
    return logits


def decoder_forward(p: dict, env_name: str, td, cache: dict, num_starts: int = 0):
    """decoder.py:151-206 -> (logits [R,N] fp32, mask [R,N] bool) in (s b) order."""
    if num_starts > 1:
        td = unbatchify(td, num_starts)
    q = context_embedding(p, env_name, cache["node_embeddings"], td) + cache["graph_context"]
    q = q.unsqueeze(1) if q.ndim == 2 else q
    mask = td["action_mask"]
    logits = pointer_attention(p, q, cache["glimpse_key"], cache["glimpse_val"], cache["logit_key"], mask)
    dt = logits.dtype
    bias = p["alpha"] * gather_by_index(td["distance_matrix"].to(dt), td["current_node"], dim=-2)
    if env_name == "rcvrptw":
        bias = bias + p["beta"] * gather_by_index(td["duration_matrix"].to(dt), td["current_node"], dim=-2)
    if dt != torch.float64:
        logits, bias = logits.to(torch.float32), bias.to(torch.float32)
    logits = torch.log(torch.exp(logits - bias) + 1e-6)  # decoder.py:198
    if num_starts > 1:
        logits = logits.permute(1, 0, 2).reshape(-1, logits.size(-1))  # "b s l -> (s b) l"
        mask = mask.permute(1, 0, 2).reshape(-1, mask.size(-1))
    return logits, mask


# ---------------------------------------------------------------------------
# decoding strategies
# ---------------------------------------------------------------------------

def process_logits(logits, mask, temperature=1.0, tanh_clipping=10.0, mask_logits=True):
    """decoding.py:311-361 with top_k = top_p = 0."""
    if tanh_clipping > 0:
        logits = torch.tanh(logits) * tanh_clipping
    if mask_logits:
        logits[~mask] = float("-inf")
    logits = logits / temperature
    return F.log_softmax(logits, dim=-1)


class Strategy:
    """Greedy / Sampling / Evaluate (+ multistart pre-hook), decoding.py:68-399."""

    def __init__(self, kind: str, multistart: bool, num_starts=None, temperature=1.0, tanh_clipping=10.0,
                 mask_logits=True, generator=None, gumbel_noise=None):
        assert kind in ("greedy", "sampling", "evaluate")
        self.kind, self.multistart, self.num_starts = kind, multistart, num_starts
        self.temperature, self.tanh_clipping, self.mask_logits = temperature, tanh_clipping, mask_logits
        self.generator = generator
        self.gumbel_noise = gumbel_noise  # optional callable(step, shape)->noise: Gumbel-max twin of the kernel
        self.actions, self.logprobs = [], []

    def pre_decoder_hook(self, td, env):  # decoding.py:157-205
        if self.multistart:
            if self.num_starts is None:
                self.num_starts = env.get_num_starts(td)
        else:
            self.num_starts = 0
        if self.num_starts >= 1 and self.multistart:
            action = env.select_start_nodes(td, num_starts=self.num_starts)
            td = batchify(td, self.num_starts)
            td["action"] = action
            td = env.step(td)["next"]
            self.logprobs.append(torch.zeros_like(action, dtype=torch.float32))
            self.actions.append(action)
        return td, env, self.num_starts

    def step(self, logits, mask, td, action=None):  # decoding.py:219-270
        logp = process_logits(logits, mask, self.temperature, self.tanh_clipping, self.mask_logits)
        if self.kind == "greedy":
            sel = logp.argmax(dim=-1)  # decoding.py:272-282
        elif self.kind == "sampling":
            if self.gumbel_noise is not None:
                sel = (logp + self.gumbel_noise(len(self.actions), logp.shape)).argmax(dim=-1)
            else:  # decoding.py:284-298
                sel = torch.multinomial(logp.exp(), 1, generator=self.generator).squeeze(1)
        else:
            sel = action
        if self.kind != "evaluate":
            assert not (~mask).gather(1, sel.unsqueeze(-1)).any(), "infeasible action selected"
        td["action"] = sel
        self.actions.append(sel)
        self.logprobs.append(gather_by_index(logp, sel, dim=1))
        return td

    def post_decoder_hook(self, td, env):  # decoding.py:207-217
        assert len(self.logprobs) > 0, "No logprobs were collected because all environments were done"
        return torch.stack(self.logprobs, 1), torch.stack(self.actions, 1), td, env


def get_log_likelihood(logprobs, return_sum=True):
    assert (logprobs > -1000).all(), "Logprobs should not be -inf, check sampling procedure!"
    return logprobs.sum(1) if return_sum else logprobs


# ---------------------------------------------------------------------------
# policy loop
# ---------------------------------------------------------------------------

def policy_forward(p: dict, env, td, row_emb, col_emb, decode_type="greedy", num_starts=None,
                   actions=None, calc_reward=True, max_steps=1_000_000, generator=None,
                   gumbel_noise=None, trace=None) -> dict:
    """policy.py:175-255 with the encoder output passed in (the encoder stays upstream PyTorch)."""
    env_name = env.name
    if actions is not None:
        decode_type = "evaluate"
    multistart = "multistart" in decode_type  # decoding.py:31-32
    if num_starts is not None:  # decoding.py:117-118: an explicit num_starts decides
        multistart = num_starts > 1
    kind = decode_type.replace("multistart_", "")
    strat = Strategy(kind, multistart, num_starts if multistart else None, generator=generator,
                     gumbel_noise=gumbel_noise)
    td, env, S = strat.pre_decoder_hook(td, env)
    cache = precompute_cache(p, row_emb, col_emb)
    step = 0
    while not td["done"].all():
        logits, mask = decoder_forward(p, env_name, td, cache, S)
        if trace is not None:
            trace.append({"logits": logits.clone(), "mask": mask.clone()})
        td = strat.step(logits, mask, td, action=actions[..., step] if actions is not None else None)
        td = env.step(td)["next"]
        step += 1
        if step > max_steps:
            break
    logprobs, acts, td, env = strat.post_decoder_hook(td, env)
    out = {}
    if calc_reward:
        if env.normalize:
            real, norm = env.get_reward(td, acts)
            td["reward"] = real
            out["normalized_reward"] = norm
        else:
            td["reward"] = env.get_reward(td, acts)
    out["reward"] = td["reward"]
    out["log_likelihood"] = get_log_likelihood(logprobs)
    out["actions"] = acts
    out["logprobs"] = logprobs
    out["td"] = td
    return out

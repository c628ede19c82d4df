"""rl4co.utils.ops subset (batchify / unbatchify / gather_by_index / ...), recalled from rl4co 0.6.0."""
import torch
from tensordict import TensorDict


def _batchify_single(x, repeats):
    s = x.shape
    return x.expand(repeats, *s).contiguous().view(s[0] * repeats, *s[1:])


def batchify(x, shape):
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _batchify_single(x, s) if s > 0 else x
    return x


def _unbatchify_single(x, repeats):
    s = x.shape
    return x.view(repeats, s[0] // repeats, *s[1:]).permute(1, 0, *range(2, len(s) + 1))


def unbatchify(x, shape):
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _unbatchify_single(x, s) if s > 0 else x
    return x


def gather_by_index(src, idx, dim=1, squeeze=True):
    expanded_shape = list(src.shape)
    expanded_shape[dim] = -1
    idx = idx.view(idx.shape + (1,) * (src.dim() - idx.dim())).expand(expanded_shape)
    squeeze = idx.size(dim) == 1 and squeeze
    return src.gather(dim, idx).squeeze(dim) if squeeze else src.gather(dim, idx)


def unbatchify_and_gather(x, idx, n):
    x = unbatchify(x, n)
    return gather_by_index(x, idx, dim=idx.dim())


def get_distance(x, y):
    return (x - y).norm(p=2, dim=-1)


def get_distance_matrix(locs):
    return (locs[..., :, None, :] - locs[..., None, :, :]).norm(p=2, dim=-1)


def calculate_entropy(logprobs):
    logprobs = torch.nan_to_num(logprobs, nan=0.0)
    entropy = -(logprobs.exp() * logprobs).sum(dim=-1)
    entropy = entropy.sum(dim=1)
    assert entropy.isfinite().all(), "Entropy is not finite"
    return entropy


def get_num_starts(td, env_name=None):
    num_starts = td["action_mask"].shape[-1]
    if env_name == "pdp":
        num_starts = (num_starts - 1) // 2
    elif env_name in ["cvrp", "cvrptw", "sdvrp", "mtsp", "op", "pctsp", "spctsp"]:
        num_starts = num_starts - 1
    return num_starts


def select_start_nodes(td, env, num_starts):
    num_loc = env.generator.num_loc if hasattr(env.generator, "num_loc") else 0xFFFFFFFF
    if env.name in ["tsp", "atsp", "flp", "mcp"]:
        selected = torch.arange(num_starts, device=td.device).repeat_interleave(td.shape[0]) % num_loc
    else:
        selected = torch.arange(num_starts, device=td.device).repeat_interleave(td.shape[0]) % num_loc + 1
    return selected

"""rl4co context embeddings used by the reference registry (rrnco/models/env_embeddings/__init__.py:43-47)."""
import torch
import torch.nn as nn
from rl4co.utils.ops import gather_by_index


class EnvContext(nn.Module):
    def __init__(self, embed_dim, step_context_dim=None, linear_bias=False):
        super().__init__()
        self.embed_dim = embed_dim
        step_context_dim = step_context_dim if step_context_dim is not None else embed_dim
        self.project_context = nn.Linear(step_context_dim, embed_dim, bias=linear_bias)

    def _cur_node_embedding(self, embeddings, td):
        return gather_by_index(embeddings, td["current_node"])

    def _state_embedding(self, embeddings, td):
        raise NotImplementedError

    def forward(self, embeddings, td):
        cur = self._cur_node_embedding(embeddings, td)
        state = self._state_embedding(embeddings, td)
        return self.project_context(torch.cat([cur, state], -1))


class TSPContext(EnvContext):
    def __init__(self, embed_dim):
        super().__init__(embed_dim, 2 * embed_dim)
        self.W_placeholder = nn.Parameter(torch.Tensor(2 * self.embed_dim).uniform_(-1, 1))

    def forward(self, embeddings, td):
        batch_size = embeddings.size(0)
        node_dim = (-1,) if td["first_node"].dim() == 1 else (td["first_node"].size(-1), -1)
        if td["i"][(0,) * td["i"].dim()].item() < 1:
            if len(td.batch_size) < 2:
                ctx = self.W_placeholder[None, :].expand(batch_size, self.W_placeholder.size(-1))
            else:
                ctx = self.W_placeholder[None, None, :].expand(
                    batch_size, td.batch_size[1], self.W_placeholder.size(-1))
        else:
            ctx = gather_by_index(
                embeddings, torch.stack([td["first_node"], td["current_node"]], -1).view(batch_size, -1)
            ).view(batch_size, *node_dim)
        return self.project_context(ctx)


class VRPContext(EnvContext):
    def __init__(self, embed_dim):
        super().__init__(embed_dim=embed_dim, step_context_dim=embed_dim + 1)

    def _state_embedding(self, embeddings, td):
        return td["vehicle_capacity"] - td["used_capacity"]


class VRPTWContext(VRPContext):
    def __init__(self, embed_dim):
        EnvContext.__init__(self, embed_dim=embed_dim, step_context_dim=embed_dim + 2)

    def _state_embedding(self, embeddings, td):
        return torch.cat([td["vehicle_capacity"] - td["used_capacity"], td["current_time"]], -1)

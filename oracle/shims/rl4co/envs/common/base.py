"""RL4COEnvBase surface used by the reference hot path (reset / step / get_reward / start nodes)."""
import torch
from rl4co.utils import ops


class RL4COEnvBase:
    name = "base"

    def __init__(self, *, check_solution=True, device="cpu", **kwargs):
        self.check_solution = check_solution
        self.device = device

    def reset(self, td=None, batch_size=None):
        if batch_size is None:
            batch_size = td.batch_size
        batch_size = list(batch_size)
        if td is None or len(td.keys()) == 0:
            td = self.generator(batch_size)
        self.device = td.device
        out = self._reset(td, batch_size=batch_size)
        out.set("done", torch.zeros(*batch_size, 1, dtype=torch.bool, device=td.device))
        return out

    def step(self, td):
        td = self._step(td)
        return {"next": td}

    def get_reward(self, td, actions):
        if self.check_solution:
            self.check_solution_validity(td, actions)
        return self._get_reward(td, actions)

    def get_num_starts(self, td):
        return ops.get_num_starts(td, self.name)

    def select_start_nodes(self, td, num_starts):
        return ops.select_start_nodes(td, self, num_starts)

    def to(self, device):
        self.device = device
        return self

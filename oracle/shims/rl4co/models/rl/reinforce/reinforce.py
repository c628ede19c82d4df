import torch.nn as nn


class REINFORCE(nn.Module):
    """Placeholder so that rrnco.models.rl imports; the training module is out of scope."""

    def __init__(self, env, policy, baseline="shared", **kw):
        super().__init__()
        self.env, self.policy = env, policy

    def save_hyperparameters(self, *a, **k):
        pass

    def set_decode_type_multistart(self, phase):
        attr = f"{phase}_decode_type"
        val = getattr(self.policy, attr, None)
        if val is not None and "multistart" not in val:
            setattr(self.policy, attr, f"multistart_{val}")

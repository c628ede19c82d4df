import torch.nn as nn


class MLP(nn.Module):
    """rl4co MLP with defaults (no norms, dropout 0, identity out-act): lins.{i}."""

    def __init__(self, input_dim, output_dim, num_neurons=[64, 32], dropout_probs=None,
                 hidden_act="ReLU", out_act="Identity", input_norm="None", output_norm="None"):
        super().__init__()
        assert input_norm == "None" and output_norm == "None" and out_act == "Identity"
        self.hidden_act = getattr(nn, hidden_act)()
        self.out_act = nn.Identity()
        sizes = [input_dim] + list(num_neurons) + [output_dim]
        self.lins = nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))

    def forward(self, x):
        for lin in self.lins[:-1]:
            x = self.hidden_act(lin(x))
        return self.out_act(self.lins[-1](x))

"""The two vector-field MLPs the ANCDE models use, with the reference's parameter names so that its
state_dicts load unchanged (experiments/models/vector_fields.py:185-250).

Inside a solve these modules are never called: the fused kernel reads their weights
(`solver.mlp_params`).  `forward` exists for the API (shape checks, stand-alone use).
"""
import torch


class _TanhField(torch.nn.Module):
    def _mlp(self, z):
        for lin in self._hidden():
            z = torch.relu(lin(z))
        return self.linear_out(z)

    def extra_repr(self):
        return "input_channels: {}, hidden_channels: {}, hidden_hidden_channels: {}, num_hidden_layers: {}" \
               "".format(self.input_channels, self.hidden_channels, self.hidden_hidden_channels, self.num_hidden_layers)


class FinalTanh(_TanhField):
    """g: z (H) -> tanh(MLP(z)).view(H, C).  Reference: vector_fields.py:219-250."""

    def __init__(self, input_channels, hidden_channels, hidden_hidden_channels, num_hidden_layers):
        super(FinalTanh, self).__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.hidden_hidden_channels = hidden_hidden_channels
        self.num_hidden_layers = num_hidden_layers
        self.linear_in = torch.nn.Linear(hidden_channels, hidden_hidden_channels)
        self.linears = torch.nn.ModuleList(torch.nn.Linear(hidden_hidden_channels, hidden_hidden_channels)
                                           for _ in range(num_hidden_layers - 1))
        self.linear_out = torch.nn.Linear(hidden_hidden_channels, input_channels * hidden_channels)

    def _hidden(self):
        return [self.linear_in] + list(self.linears)

    def forward(self, z):
        out = self._mlp(z)
        return torch.tanh(out.view(*z.shape[:-1], self.hidden_channels, self.input_channels))


class FinalTanh_ff6(_TanhField):
    """f: a (C) -> tanh(MLP(a)).view(C, C), the attention NCDE's field.  Reference: vector_fields.py:185-218."""

    def __init__(self, input_channels, hidden_atten_channels, hidden_hidden_atten_channels, num_hidden_layers):
        super(FinalTanh_ff6, self).__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_atten_channels
        self.hidden_hidden_channels = hidden_hidden_atten_channels
        self.num_hidden_layers = num_hidden_layers
        self.linear_in = torch.nn.Linear(input_channels, hidden_hidden_atten_channels)
        self.linear_test = torch.nn.Linear(hidden_hidden_atten_channels, hidden_atten_channels)
        self.linears = torch.nn.ModuleList(torch.nn.Linear(hidden_atten_channels, hidden_atten_channels)
                                           for _ in range(num_hidden_layers - 1))
        self.linear_out = torch.nn.Linear(hidden_atten_channels, input_channels * input_channels)

    def _hidden(self):
        return [self.linear_in, self.linear_test] + list(self.linears)

    def forward(self, z):
        out = self._mlp(z)
        return torch.tanh(out.view(*z.shape[:-1], self.input_channels, self.input_channels))

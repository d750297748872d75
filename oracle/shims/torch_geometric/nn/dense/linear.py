import math
import torch
import torch.nn.functional as F
from torch.nn import Parameter
from .. import inits


class Linear(torch.nn.Module):
    """PyG dense Linear: F.linear with weight [out,in]; glorot weight, zero bias."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer = weight_initializer
        self.weight = Parameter(torch.empty(out_channels, in_channels))
        self.bias = Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            inits.glorot(self.weight)
        else:  # kaiming_uniform(a=sqrt(5)) like torch.nn.Linear
            torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            inits.zeros(self.bias)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)

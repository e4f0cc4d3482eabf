"""MLP builders of the reference (src/nets/nets.py:14-53), same class names, constructor
arguments and module structure (`.net` = Sequential of Linear/Tanh) so that pickled
checkpoints are interchangeable.  All three nets are the same stack:

    Linear(in, dim) Tanh [Linear(dim, dim) Tanh] * (num_layers - 1) Linear(dim, out)

orthogonal init with gain sqrt(2) on hidden layers, `action_std` on the head (0.01 actor,
1.0 critic), zero biases.  The `dropout` argument is accepted and ignored, as in the reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn


def layer_init(layer: nn.Linear, std: float = math.sqrt(2.0), bias_const: float = 0.0) -> nn.Linear:
    nn.init.orthogonal_(layer.weight, std)
    nn.init.constant_(layer.bias, bias_const)
    return layer


def _stack(in_features: int, dim: int, out_features: int, num_layers: int, head_std: float) -> nn.Sequential:
    widths = [int(in_features)] + [int(dim)] * int(num_layers)
    mods = []
    for a, b in zip(widths[:-1], widths[1:]):
        mods += [layer_init(nn.Linear(a, b)), nn.Tanh()]
    mods.append(layer_init(nn.Linear(int(dim), int(out_features)), head_std))
    return nn.Sequential(*mods)


class _Net(nn.Module):
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.net(input)


class discrete_net(_Net):
    def __init__(self, dim: int, input_dim, output_dim: int, num_layers: int, dropout: float, action_std: float = 0.01):
        super().__init__()
        self.net = _stack(np.array(input_dim).prod(), dim, output_dim, num_layers, action_std)


class continuous_net(_Net):
    def __init__(self, dim: int, input_dim, output_dim, num_layers: int, dropout: float, action_std: float = 0.01):
        super().__init__()
        self.net = _stack(np.array(input_dim).prod(), dim, np.prod(output_dim), num_layers, action_std)


class critic(_Net):
    def __init__(self, dim: int, input_dim, num_layers: int, dropout: float, action_std: float = 1.0):
        super().__init__()
        self.net = _stack(np.array(input_dim).prod(), dim, 1, num_layers, action_std)

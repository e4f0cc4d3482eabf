"""The reference's actor_critic (src/models/actor_critic.py:8-51): same constructor, attributes
and methods.  `evaluate(state, action=None)` -> (action, log_prob, entropy, value[B,1]) and
`value(state)` -> [B]; `get_action_and_value` / `get_value` / `act` are aliases kept for callers
that use the CleanRL names or the legacy test.py:51 call.

The module owns the parameters (and the checkpoint format); the training hot path reads them
through `flat_parameters()`, a single fp32 buffer in the order include/aur_ppo.h documents, of
which every nn.Parameter is a view -- the CUDA kernels and torch see the same memory.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import nn
from torch.distributions import Categorical, Normal

from ..nets.nets import continuous_net, critic, discrete_net


class actor_critic(nn.Module):
    def __init__(self, state_dim: int, action_dim, hidden_dim: int, num_layers: int, dropout, continuous: bool) -> None:
        super().__init__()
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.hidden_dim = hidden_dim
        self.continuous = continuous
        self.num_layers = num_layers
        self.dropout = dropout
        if continuous:
            self.actor = continuous_net(hidden_dim, state_dim, action_dim, num_layers, dropout)
            self.critic = critic(hidden_dim, state_dim, num_layers, dropout)
            self.actor_logstd = nn.Parameter(torch.zeros(1, int(np.prod(action_dim))))
        else:
            self.actor = discrete_net(hidden_dim, state_dim, action_dim, num_layers, dropout)
            self.critic = critic(hidden_dim, state_dim, num_layers, dropout)

    def forward(self):
        pass

    # ------------------------------------------------------------------ reference API
    def value(self, state: torch.Tensor) -> torch.Tensor:
        return self.critic(state).flatten()

    def evaluate(self, state: torch.Tensor, action: Optional[torch.Tensor] = None):
        if self.continuous:
            mean = self.actor(state)
            dist = Normal(mean, torch.exp(self.actor_logstd.expand_as(mean)))
            if action is None:
                action = dist.sample()
            log_prob, entropy = dist.log_prob(action).sum(1), dist.entropy().sum(1)
        else:
            dist = Categorical(logits=self.actor(state))
            if action is None:
                action = dist.sample()
            log_prob, entropy = dist.log_prob(action), dist.entropy()
        return action, log_prob, entropy, self.critic(state)

    # aliases (CleanRL names used by BASELINE.json; legacy test.py:51 `act`)
    def get_action_and_value(self, state, action=None):
        return self.evaluate(state, action)

    def get_value(self, state):
        return self.value(state)

    def act(self, state):
        action, log_prob, _, value = self.evaluate(state)
        return action, log_prob, value

    # --------------------------------------------------------------- kernel interface
    @staticmethod
    def _linears(seq: nn.Sequential) -> List[nn.Linear]:
        """Linear layers in order; legacy checkpoints interleave Dropout (indices 0,3,6), current
        code does not (0,2,4) -- both map by order."""
        return [m for m in seq if isinstance(m, nn.Linear)]

    def kernel_shape(self) -> Tuple[int, int, int, int, bool]:
        """(obs_dim, act_dim, hidden_dim, num_hidden_layers, continuous) as the C ABI wants them."""
        lin = self._linears(self.actor.net)
        return lin[0].in_features, lin[-1].out_features, lin[0].out_features, len(lin) - 1, bool(self.continuous)

    def _ordered_parameters(self) -> List[nn.Parameter]:
        out: List[nn.Parameter] = []
        for net in (self.actor.net, self.critic.net):
            for lin in self._linears(net):
                out += [lin.weight, lin.bias]
        if self.continuous:
            out.append(self.actor_logstd)
        return out

    def flat_parameters(self) -> torch.Tensor:
        """One contiguous fp32 buffer [actor | critic | logstd]; parameters become views of it."""
        flat = getattr(self, "_flat", None)
        ps = self._ordered_parameters()
        if flat is not None and flat.device == ps[0].device and all(p.data_ptr() == flat.data_ptr() + 4 * o
                                                                     for p, o in zip(ps, self._flat_offsets)):
            return flat
        n = sum(p.numel() for p in ps)
        flat = torch.empty(n, dtype=torch.float32, device=ps[0].device)
        offs, o = [], 0
        for p in ps:
            flat[o:o + p.numel()].copy_(p.detach().reshape(-1).float())
            p.data = flat[o:o + p.numel()].view_as(p)
            offs.append(o)
            o += p.numel()
        object.__setattr__(self, "_flat", flat)
        object.__setattr__(self, "_flat_offsets", offs)
        return flat

    def __getstate__(self):
        # keep the pickle free of the flat-buffer bookkeeping: the checkpoint is the plain module
        state = self.__dict__.copy()
        state.pop("_flat", None)
        state.pop("_flat_offsets", None)
        return state

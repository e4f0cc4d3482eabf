"""Checkpoint compatibility with the reference's whole-module pickles.

The reference saves `torch.save(self.policy, 'actor_critic_<L>.pt')` (src/ppo.py:296) from a
process whose import root is src/, so the pickle names the classes
`models.actor_critic.actor_critic`, `nets.nets.discrete_net`, `nets.nets.continuous_net` and
`nets.nets.critic` (or the same under `src.` when imported as a package, run_ppo.py:6).  To read
those files here -- and to write files the reference's test.py:49 `torch.load` can read -- the
same dotted names must resolve to classes with the same structure.  `install()` registers alias
modules under those names (leaving any real module of that name alone) and stamps `__module__`
on our classes so the pickles we write carry the reference's paths.
"""
from __future__ import annotations

import sys
import types

import torch

from .models.actor_critic import actor_critic as _actor_critic
from .nets.nets import continuous_net as _continuous_net, critic as _critic, discrete_net as _discrete_net
from .nets.nets import layer_init as _layer_init

_installed = False


def _alias(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__aur_alias__ = True
        sys.modules[name] = mod
    if getattr(mod, "__aur_alias__", False):
        for k, v in attrs.items():
            setattr(mod, k, v)
    return mod


def install() -> None:
    global _installed
    if _installed:
        return
    net_classes = dict(discrete_net=_discrete_net, continuous_net=_continuous_net, critic=_critic, layer_init=_layer_init)
    for root in ("", "src."):
        if root:
            _alias("src")
        models = _alias(root + "models")
        m_ac = _alias(root + "models.actor_critic", actor_critic=_actor_critic)
        nets = _alias(root + "nets", **{k: v for k, v in net_classes.items() if k != "layer_init"})
        n_nets = _alias(root + "nets.nets", **net_classes)
        if getattr(models, "__aur_alias__", False):
            models.actor_critic = m_ac
        if getattr(nets, "__aur_alias__", False):
            nets.nets = n_nets
        if root:
            src = sys.modules["src"]
            if getattr(src, "__aur_alias__", False):
                src.models, src.nets = models, nets
    if getattr(sys.modules["models.actor_critic"], "actor_critic", None) is _actor_critic:
        _actor_critic.__module__ = "models.actor_critic"
    if getattr(sys.modules["nets.nets"], "critic", None) is _critic:
        for cls in (_discrete_net, _continuous_net, _critic):
            cls.__module__ = "nets.nets"
    _installed = True


def save_policy(policy, path: str) -> None:
    """torch.save(self.policy, path) with the reference's class paths (src/ppo.py:296)."""
    install()
    torch.save(policy, path)


def load_policy(path: str, map_location="cpu"):
    """torch.load of a reference (or our) whole-module checkpoint; legacy Dropout-interleaved
    Sequentials (actor.net.{0,3,6}) and the current layout ({0,2,4}) both load."""
    install()
    return torch.load(path, map_location=map_location, weights_only=False)

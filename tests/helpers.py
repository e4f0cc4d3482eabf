"""Shared helpers for the GPU parity tests."""
import numpy as np
import torch

from oracle import ppo_ref as R


def flat_from_named(named: dict) -> np.ndarray:
    """Canonical flat parameter order of include/aur_ppo.h from the reference's names."""
    def net(prefix):
        idx = sorted({int(k.split(".")[2]) for k in named if k.startswith(prefix + ".net.")})
        out = []
        for i in idx:
            out += [np.asarray(named[f"{prefix}.net.{i}.weight"]).ravel(), np.asarray(named[f"{prefix}.net.{i}.bias"]).ravel()]
        return out
    parts = net("actor") + net("critic")
    if "actor_logstd" in named:
        parts.append(np.asarray(named["actor_logstd"]).ravel())
    return np.concatenate(parts).astype(np.float32)


def golden_policy(golden_dir, tag, file="model.npz", prefix="p"):
    import os
    g = np.load(os.path.join(golden_dir, file))
    names = [str(n) for n in g[f"{tag}_names"]]
    named = {n: g[f"{tag}_{prefix}_{n}"] for n in names}
    return R.RefPolicy(named, continuous="actor_logstd" in names), named, g


def random_policy(obs_dim, act_dim, hidden, num_layers, continuous, seed=0, scale=1.0):
    """Reference-shaped parameters with torch's default Linear init scale (deterministic)."""
    g = torch.Generator().manual_seed(seed)
    named = {}
    def net(prefix, out):
        dims = [obs_dim] + [hidden] * num_layers + [out]
        for li in range(len(dims) - 1):
            bound = scale / np.sqrt(dims[li])
            named[f"{prefix}.net.{2 * li}.weight"] = ((torch.rand(dims[li + 1], dims[li], generator=g) * 2 - 1) * bound).numpy()
            named[f"{prefix}.net.{2 * li}.bias"] = ((torch.rand(dims[li + 1], generator=g) * 2 - 1) * bound).numpy()
    net("actor", act_dim)
    net("critic", 1)
    if continuous:
        named["actor_logstd"] = (torch.rand(1, act_dim, generator=g) * 0.6 - 0.5).numpy()
    return R.RefPolicy(named, continuous), named

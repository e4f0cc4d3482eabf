"""Fixtures for the checkpoint consumer -- TEST INFRASTRUCTURE ONLY.

Copies the reference's shipped whole-module checkpoints (plots/actor_critic.pt == src/models/saved/actor_critic.pt,
src/models/saved/actor_critic_2.pt, src/models/saved/actor_critic_10.pt: model weights, not source) to
tests/golden/ckpt/ and records what the REFERENCE'S OWN classes compute from them: the pickles are loaded with
/root/reference/src on sys.path, so `models.actor_critic.actor_critic` / `nets.nets.*` resolve to the reference's
code (src/models/actor_critic.py:8-51, src/nets/nets.py:14-53), and `evaluate(state, action)` / `value(state)`
are run on fixed inputs.  Run here (the GPU box has no /root/reference):  python oracle/gen_golden_ckpt.py
"""
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = {"actor_critic": "plots/actor_critic.pt", "actor_critic_saved": "src/models/saved/actor_critic.pt",
         "actor_critic_2": "src/models/saved/actor_critic_2.pt", "actor_critic_10": "src/models/saved/actor_critic_10.pt"}


def main():
    sys.path.insert(0, REF)                                    # `src.nets.nets` (models/actor_critic.py:3)
    sys.path.insert(0, os.path.join(REF, "src"))               # `models.actor_critic` as the pickles name it
    out = {}
    dst = os.path.join(ROOT, "tests", "golden", "ckpt")
    os.makedirs(dst, exist_ok=True)
    for tag, rel in FILES.items():
        src = os.path.join(REF, rel)
        if tag != "actor_critic_saved":                       # byte-identical to plots/actor_critic.pt (asserted below)
            shutil.copyfile(src, os.path.join(dst, tag + ".pt"))
        m = torch.load(src, map_location="cpu", weights_only=False)
        assert type(m).__module__ == "models.actor_critic" and "reference" in sys.modules["models.actor_critic"].__file__
        sd = m.state_dict()
        obs_dim = sd["actor.net.0.weight"].shape[1]
        n_act = [v for k, v in sd.items() if k.startswith("actor.net.") and k.endswith(".bias")][-1].shape[0]
        g = torch.Generator().manual_seed(7)
        obs = (torch.rand(64, obs_dim, generator=g) * 2 - 1) * torch.tensor([2.4, 3.0, 0.21, 3.0] * (obs_dim // 4))
        act = torch.randint(0, n_act, (64,), generator=g)
        with torch.no_grad():
            a, lp, ent, val = m.evaluate(obs, act)
            logits = m.actor(obs)
            v2 = m.value(obs)
        out[f"{tag}_obs"], out[f"{tag}_act"] = obs.numpy(), act.numpy()
        out[f"{tag}_logp"], out[f"{tag}_entropy"], out[f"{tag}_value"] = lp.numpy(), ent.numpy(), val.numpy().reshape(-1)
        out[f"{tag}_logits"], out[f"{tag}_value2"] = logits.numpy(), v2.numpy()
        out[f"{tag}_keys"] = np.array(list(sd.keys()))
    a = open(os.path.join(REF, FILES["actor_critic"]), "rb").read()
    b = open(os.path.join(REF, FILES["actor_critic_saved"]), "rb").read()
    out["saved_equals_plots"] = np.array(a == b)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ckpt.npz"), **out)
    print("wrote tests/golden/ckpt.npz and", sorted(os.listdir(dst)))


if __name__ == "__main__":
    main()

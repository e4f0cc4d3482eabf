"""oracle/cnn_ref.py (restated plain CNN actor-critic) against tests/golden/plain_cnn.npz, which was produced by the
reference's OWN base_actor / base_critic classes (oracle/gen_golden_cnn.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import cnn_ref as C


def test_restated_plain_cnn_matches_reference_classes(golden_dir):
    g = np.load(os.path.join(golden_dir, "plain_cnn.npz"))
    p = {k: v.clone().requires_grad_(True) for k, v in C.formula_params(C.param_shapes()).items()}
    obs, state, action, adv, ret = C.golden_inputs(2)
    lp, ent, val = C.evaluate(p, state, obs, action)
    np.testing.assert_allclose(lp.detach().numpy(), g["logp"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(ent.detach().numpy(), g["entropy"], rtol=1e-6)
    np.testing.assert_allclose(val.detach().numpy(), g["value"], rtol=2e-5, atol=1e-6)
    loss, st = C.update_loss(p, state, obs, action, torch.from_numpy(g["oldlp"]), adv, ret, torch.from_numpy(g["vold"]))
    assert abs(st["policy_loss"] - float(g["policy_loss"])) < 1e-5 and abs(st["value_loss"] - float(g["value_loss"])) < 1e-5
    assert abs(st["loss"] - float(g["loss"])) < 1e-5
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    assert names == list(p.keys())                       # the reference modules' own state_dict names
    for i, n in enumerate(names):
        gr = p[n].grad
        np.testing.assert_allclose(float(gr.norm()), g["grad_norm"][i], rtol=2e-4, atol=1e-7, err_msg=n)
        probe = gr.reshape(-1)[[0, gr.numel() // 3, gr.numel() // 2, -1]].numpy()
        np.testing.assert_allclose(probe, g["grad_probe"][i], rtol=2e-3, atol=1e-6 + 2e-4 * float(np.abs(g["grad_probe"][i]).max()), err_msg=n)


def test_quantised_checker_stays_close_to_fp32():
    p = C.formula_params(C.param_shapes())
    obs, state, action, adv, ret = C.golden_inputs(2)
    a = C.evaluate(p, state, obs, action)
    b = C.evaluate(p, state, obs, action, quant=True)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x.numpy(), y.numpy(), rtol=5e-2, atol=5e-2)

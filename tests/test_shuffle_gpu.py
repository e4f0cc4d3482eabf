"""Device minibatch shuffle (aur_shuffle_indices) vs its numpy restatement: bit-exact, a true permutation, and
statistically a shuffle (np.random.shuffle of ppo.py:214-215 is replaced by a keyed bijection)."""
import numpy as np
import pytest
import torch

from aur_ppo_b200 import kernels
from oracle import ppo_ref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 3, 17, 256, 1000, 65536, 65537, 1 << 20, 2097152 + 5])
def test_shuffle_matches_oracle_bit_exact(n):
    got = kernels.shuffle_indices(n, seed=1, stream_id=7).cpu().numpy()
    np.testing.assert_array_equal(got, R.feistel_shuffle(n, 1, 7))


def test_shuffle_is_a_permutation_at_full_size():
    n = 8388608  # BASELINE config B batch
    p = kernels.shuffle_indices(n, seed=3, stream_id=0)
    assert torch.equal(torch.sort(p.long()).values, torch.arange(n, device="cuda"))
    q = kernels.shuffle_indices(n, seed=3, stream_id=1)
    assert (p == q).float().mean().item() < 1e-4              # a different epoch is a different permutation
    assert torch.equal(p, kernels.shuffle_indices(n, seed=3, stream_id=0))   # deterministic


def test_shuffle_statistics():
    n = 1 << 20
    p = kernels.shuffle_indices(n, seed=11, stream_id=5).double()
    i = torch.arange(n, device="cuda", dtype=torch.double)
    # rank correlation with the identity ~ N(0, 1/n); mean displacement of a uniform permutation = n/3
    corr = torch.corrcoef(torch.stack([p, i]))[0, 1].item()
    assert abs(corr) < 5 / np.sqrt(n)
    assert abs((p - i).abs().mean().item() / n - 1 / 3) < 0.01
    # every minibatch-sized slice covers the index range evenly (chi-square over 64 buckets, 4 slices)
    for sl in p.reshape(4, -1):
        h = torch.histc(sl, bins=64, min=0, max=n)
        e = sl.numel() / 64
        chi2 = ((h - e) ** 2 / e).sum().item()
        assert chi2 < 140, chi2                                  # 63 dof: P(chi2 > 140) ~ 1e-7


def test_shuffle_rejects_bad_arguments():
    from aur_ppo_b200 import _lib
    with pytest.raises(_lib.AurError):
        kernels.shuffle_indices(8, 0, 0, out=torch.empty(8, dtype=torch.int64, device="cuda"))

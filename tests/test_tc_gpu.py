"""tcgen05 GEMM core vs torch (GPU)."""
import pytest
import torch

from aur_ppo_b200 import kernels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (256, 384, 576), (1000, 200, 72), (4096, 512, 1152), (128 * 400, 128, 64)])   # last: 400 CTAs = several waves per SM
def test_gemm_bf16_matches_fp32_matmul_of_bf16_inputs(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    b = torch.randn(N, K, generator=g, device="cuda").bfloat16()
    c = kernels.tc_gemm_bf16(a, b)
    ref = a.float() @ b.float().T          # exact products of bf16 inputs, fp32 accumulation
    torch.testing.assert_close(c, ref, rtol=1e-4, atol=1e-3 * (K ** 0.5) * 1e-1)

"""tcgen05 engine self test: plain GEMMs through the TMA + UMMA + TMEM pipeline against torch fp32
matmul of the same 16-bit-rounded operands (so the only difference is fp32 summation order)."""
import pytest
import torch

from constant_memory_waveglow_b200 import ops

pytestmark = pytest.mark.gpu


def _rand(shape, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(dtype).cuda()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,variant", [
    (128, 256, 64, 0), (128, 128, 64, 2), (256, 256, 256, 0), (384, 512, 832, 0),
    (200, 128, 128, 2),      # ragged M: TMA zero fill + masked epilogue
    (1000, 512, 768, 0), (4096, 256, 1024, 0),
])
def test_kmajor_gemm(dtype, M, N, K, variant):
    a = _rand((M, K), dtype, 1)
    b = _rand((N, K), dtype, 2)
    d = ops.selftest_tc_gemm(a, b, M, N, K, variant)
    ref = a.float() @ b.float().t()
    err = (d - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("dtype", [torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 128), (256, 256, 512), (512, 256, 2000), (192, 128, 300)])
def test_mnmajor_wgrad_gemm(dtype, M, N, K):
    a = _rand((K, M), dtype, 3)   # A[t][m]
    b = _rand((K, N), dtype, 4)   # B[t][n]
    d = ops.selftest_tc_gemm(a, b, M, N, K, 1)
    ref = a.float().t() @ b.float()
    err = (d - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err

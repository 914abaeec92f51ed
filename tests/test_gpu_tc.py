"""tcgen05 kernels vs fp32 torch (bf16-rounded operands, fp32 accumulation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("mode,K,N", [(0, 64, 256), (0, 256, 128), (1, 64, 128), (1, 128, 256), (2, 128, 64), (2, 128, 256)])
def test_tc_selftest_tile(mode, K, N):
    from ncf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(mode * 100 + K + N)
    if mode == 0:
        A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
        ref = _bf(A) @ _bf(B).t()
    elif mode == 1:
        A, B = torch.randn(128, K, generator=g), torch.randn(K, N, generator=g)
        ref = _bf(A) @ _bf(B)
    else:
        A, B = torch.randn(128, 128, generator=g), torch.randn(128, N, generator=g)
        ref = _bf(A).t() @ _bf(B)
    dA, dB = A.cuda(), B.cuda()
    D = torch.zeros(128, N, device="cuda")
    _lib.check(lib.ncf_tc_selftest(mode, K, N, _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(D), None))
    torch.cuda.synchronize()
    err = float((D.cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"mode {mode} K {K} N {N}: rel err {err}"

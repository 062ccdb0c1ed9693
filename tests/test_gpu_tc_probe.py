"""tcgen05 bring-up: the shared-memory / instruction descriptor conventions used by the tensor-core
convolution kernels, checked against torch.matmul on the GPU (tf32: 10-bit mantissa inputs, fp32 accumulate)."""
import pytest
import torch

from endo_b200 import _lib

pytestmark = pytest.mark.gpu
TF32, BF16 = 2, 1


def _run(a_rows, n, k, shift, fmt, a_mn, b_mn):
    g = torch.Generator().manual_seed(a_rows * 7 + n * 3 + k + shift)
    a = torch.randn(a_rows, k, generator=g).cuda()
    b = torch.randn(n, k, generator=g).cuda()
    d = torch.full((128, n), float("nan"), device="cuda")
    _lib.check(_lib.lib().endo_tc_probe(a.data_ptr(), b.data_ptr(), d.data_ptr(), a_rows, n, k, shift, fmt, a_mn, b_mn,
                                        _lib.stream_ptr(a.device)), "tc_probe")
    torch.cuda.synchronize()
    if fmt == BF16:
        ar, br = a.bfloat16().double(), b.bfloat16().double()
        tol = 1e-5
    else:
        # tf32 keeps 10 mantissa bits of each input (truncation or rounding is implementation-defined)
        ar, br = a.double(), b.double()
        tol = 2e-3
    ref = ar[shift:shift + 128] @ br.t()
    err = float((d.double() - ref).abs().max() / ref.abs().max())
    return err, tol


@pytest.mark.parametrize("n,k,shift", [(16, 8, 0), (16, 32, 0), (48, 16, 0), (48, 16, 5), (48, 16, 35), (192, 64, 34),
                                       (256, 16, 1)])
def test_tf32_k_major(n, k, shift):
    err, tol = _run(168, n, k, shift, TF32, 0, 0)
    assert err < tol, err


@pytest.mark.parametrize("n,k,shift", [(16, 16, 0), (48, 32, 3), (192, 64, 34)])
def test_bf16_k_major(n, k, shift):
    err, tol = _run(168, n, k, shift, BF16, 0, 0)
    assert err < tol, err


@pytest.mark.parametrize("fmt,k", [(TF32, 8), (TF32, 32), (BF16, 16), (BF16, 48)])
@pytest.mark.parametrize("a_mn,b_mn", [(1, 0), (0, 1), (1, 1)])
def test_mn_major(fmt, k, a_mn, b_mn):
    err, tol = _run(160, 64, k, 8, fmt, a_mn, b_mn)
    assert err < tol, err

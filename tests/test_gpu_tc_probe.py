"""tcgen05 bring-up: the shared-memory / instruction descriptor conventions used by the tensor-core
convolution kernels, checked against torch.matmul on the GPU (tf32: 10-bit mantissa inputs, fp32 accumulate)."""
import pytest
import torch

from endo_b200 import _lib

pytestmark = pytest.mark.gpu
TF32, BF16 = 2, 1


def _run(a_rows, n, k, shift, fmt, a_mn, b_mn, swz=0, reps=1, rotate=1):
    g = torch.Generator().manual_seed(a_rows * 7 + n * 3 + k + shift)
    a = torch.randn(a_rows, k, generator=g).cuda()
    b = torch.randn(n, k, generator=g).cuda()
    d = torch.full((128, n), float("nan"), device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().endo_tc_probe(a.data_ptr(), b.data_ptr(), d.data_ptr(), a_rows, n, k, shift, fmt, a_mn, b_mn,
                                        swz, reps, cyc.data_ptr(), rotate, _lib.stream_ptr(a.device)), "tc_probe")
    torch.cuda.synchronize()
    _run.cycles = int(cyc.item())
    if fmt == BF16:
        ar, br = a.bfloat16().double(), b.bfloat16().double()
        tol = 1e-5
    else:
        # tf32 keeps 10 mantissa bits of each input (truncation or rounding is implementation-defined)
        ar, br = a.double(), b.double()
        tol = 2e-3
    ref = (ar[shift:shift + 128] @ br.t()) * ((reps + rotate - 1) // rotate)
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


@pytest.mark.parametrize("fmt,k", [(BF16, 16), (BF16, 48)])      # tf32 MN-major returns zeros with this layout (measured): unused
@pytest.mark.parametrize("a_mn,b_mn", [(1, 0), (0, 1), (1, 1)])
def test_mn_major(fmt, k, a_mn, b_mn):
    err, tol = _run(160, 64, k, 8, fmt, a_mn, b_mn)
    assert err < tol, err


@pytest.mark.parametrize("fmt,n,k", [(TF32, 48, 32), (BF16, 48, 64)])
def test_swizzle128_k_major(fmt, n, k):
    """SWIZZLE_128B operands work when the start address is atom-aligned.  Sliding the start address by single rows
    (the trick the convolution kernels use with SWIZZLE_NONE) did NOT give correct results with base_offset =
    (start >> 7) & 7 on this hardware/driver, and the MMA cost is the same for both layouts (see below), so the
    kernels stay on SWIZZLE_NONE."""
    err, tol = _run(168, n, k, 0, fmt, 0, 0, swz=2)
    assert err < tol, err


def test_mma_cost_by_operand_layout():
    """Microbenchmark (printed with -s): cycles per 128xNx8 tf32 MMA when every MMA accumulates into the SAME TMEM
    tile.  Measured on B200: ~266 cycles for N = 48, 64 and 192, SWIZZLE_NONE and SWIZZLE_128B alike, i.e. a
    dependent accumulate is latency-bound.  The convolution kernels therefore interleave independent accumulators."""
    for n in (48, 64, 192):
        for swz in (0, 2):
            k, reps = 32, 64
            err, tol = _run(168, n, k, 0, TF32, 0, 0, swz=swz, reps=reps)
            n_mma = reps * k // 8
            print(f"N={n:3d} swizzle={swz}: {_run.cycles / n_mma:7.1f} cycles per MMA ({n_mma} MMAs), err {err:.1e}")
            assert err < 5e-3


def test_mma_cost_independent_accumulators():
    """Same microbenchmark, but consecutive MMAs rotate over several accumulator tiles (printed with -s): tells a
    dependent-accumulate latency from a per-instruction throughput limit."""
    for fmt, name, k in ((TF32, "tf32", 32), (BF16, "bf16", 64)):
        for n, rot in ((48, 1), (48, 8), (64, 8), (128, 4), (256, 2)):
            reps = 64
            err, tol = _run(168, n, k, 0, fmt, 0, 0, swz=0, reps=reps, rotate=rot)
            n_mma = reps * (k // (8 if fmt == TF32 else 16))
            print(f"{name} N={n:3d} rotate={rot}: {_run.cycles / n_mma:7.1f} cycles per MMA, err {err:.1e}")
            assert err < 5e-3


def _probe(a, b, n, k):
    d = torch.full((128, n), float("nan"), device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().endo_tc_probe(a.data_ptr(), b.data_ptr(), d.data_ptr(), a.shape[0], n, k, 0, TF32, 0, 0,
                                        0, 1, cyc.data_ptr(), 1, _lib.stream_ptr(a.device)), "tc_probe")
    torch.cuda.synchronize()
    return d


def _trunc_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def test_tf32_operand_conversion_and_split_accuracy():
    """Facts the 3xTF32 (fp32-accurate) convolution mode relies on: (1) how kind::tf32 converts fp32 operands
    (truncation of the low 13 mantissa bits vs rounding), (2) that hi/lo operand splitting
    a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo recovers fp32-level accuracy from tf32 MMAs with fp32 accumulation."""
    n, k = 48, 128
    g = torch.Generator().manual_seed(5)
    a = torch.randn(128, k, generator=g).cuda()
    b = torch.randn(n, k, generator=g).cuda()
    d = _probe(a, b, n, k).double()
    ref_trunc = _trunc_tf32(a).double() @ _trunc_tf32(b).double().t()
    ref_full = a.double() @ b.double().t()
    e_trunc = float((d - ref_trunc).abs().max() / ref_full.abs().max())
    e_full = float((d - ref_full).abs().max() / ref_full.abs().max())
    print(f"tf32 MMA vs truncated-operand product: {e_trunc:.2e}; vs exact product: {e_full:.2e}")
    a_hi, b_hi = _trunc_tf32(a), _trunc_tf32(b)
    a_lo, b_lo = a - a_hi, b - b_hi
    d3 = _probe(a_hi, b_hi, n, k) + (_probe(a_lo, b_hi, n, k) + _probe(a_hi, b_lo, n, k))
    e3 = float((d3.double() - ref_full).abs().max() / ref_full.abs().max())
    e32 = float(((a @ b.t()).double() - ref_full).abs().max() / ref_full.abs().max())
    print(f"3xTF32 split: {e3:.2e}   (fp32 FFMA matmul: {e32:.2e})")
    assert e3 < 2e-6, e3


@pytest.mark.parametrize("c0,x0,y0", [(0, 0, 0), (8, -1, -1), (40, 30, 50), (184, 300, 250)])
def test_tma_box_load_with_zero_fill(c0, x0, y0):
    """cp.async.bulk.tensor box of an NHWC level buffer: dense [h][w][c] block in shared memory, zero fill outside the
    tensor (negative start coordinates, overhang past the right / bottom edge and past the last channel)."""
    import ctypes
    from endo_b200 import _lib as L
    B, H, W, C = 2, 256, 320, 192
    bc, bw, bh = 8, 34, 18
    src = torch.randn(B, H, W, C, device="cuda")
    out = torch.full((bh, bw, bc), 7.0, device="cuda")
    L.check(L.lib().endo_tma_probe(src.data_ptr(), B, H, W, C, bc, bw, bh, c0, x0, y0, 1, out.data_ptr(),
                                   L.stream_ptr(src.device)), "tma_probe")
    torch.cuda.synchronize()
    ref = torch.zeros(bh, bw, bc, device="cuda")
    ys, xs, cs = range(y0, y0 + bh), range(x0, x0 + bw), range(c0, c0 + bc)
    yv = [i for i, y in enumerate(ys) if 0 <= y < H]
    xv = [i for i, x in enumerate(xs) if 0 <= x < W]
    cv = [i for i, c in enumerate(cs) if 0 <= c < C]
    if yv and xv and cv:
        ref[yv[0]:yv[-1] + 1, xv[0]:xv[-1] + 1, cv[0]:cv[-1] + 1] = src[1, ys[yv[0]]:ys[yv[-1]] + 1, xs[xv[0]]:xs[xv[-1]] + 1,
                                                                      cs[cv[0]]:cs[cv[-1]] + 1]
    assert torch.equal(out, ref)

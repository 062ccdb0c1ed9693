"""GPU parity of the FCDenseNet engine (endo_net_fwd / endo_net_bwd through the nn.Module mirror)
against the oracle and the fixtures generated from the unmodified reference.

Forward: depth maps within 1e-4 of the reference relative to the map's scale (north_star).
Backward: fp32 gradients of this 57-layer BatchNorm network are themselves only reproducible to
~1e-3 between two fp32 implementations (tests/test_oracle_golden.py), so the CUDA gradients are
judged against the fp64 oracle with the fp32 oracle's own error as the yardstick."""
import numpy as np
import pytest
import torch

import endo_b200
from oracle import net as onet, step as ostep
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _setup(cfg, factory, b, h, w, seed, perturb=True):
    state = onet.init_state(cfg, seed=seed, perturb=perturb)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    x = batch["boundaries"] * batch["colors_1"]
    model = factory()
    model.load_state_dict(state)
    model.cuda().train()
    return state, x, model


def _oracle_fwd_bwd(state, x, cfg, gy, dtype):
    params = {}
    for k, v in state.items():
        v = v if v.dtype == torch.long else v.to(dtype)
        params[k] = v if onet.is_buffer(k) else v.clone().requires_grad_(True)
    new_buf = {}
    y = onet.forward(params, x.to(dtype), cfg, True, new_buf)
    (y * gy.to(dtype)).sum().backward()
    grads = {k: p.grad for k, p in params.items() if not onet.is_buffer(k)}
    return y.detach(), grads, new_buf


def _check_grads(model, g64, g32, names):
    """Per-tensor error vs the fp64 oracle, compared with the fp32 oracle's own error distribution.

    Individual tensors are noisy in ANY fp32 implementation of this network: a ReLU whose pre-activation is
    within rounding of zero flips its mask, which changes the gradient of a BatchNorm that sees only a few
    samples per channel (deep levels) by O(1/n).  Different fp32 codes flip different units, so the check is
    on the distribution (median / 90th percentile / max), not tensor by tensor."""
    gmax = max(float(g64[k].abs().max()) for k in names)
    params = dict(model.named_parameters())
    e_cuda, e_ref = [], []
    for k in names:
        ref = g64[k]
        scale = max(float(ref.abs().max()), 1e-5 * gmax)      # conv biases feeding a BN have zero true gradient
        e_cuda.append(float((params[k].grad.double().cpu() - ref).abs().max()) / scale)
        e_ref.append(float((g32[k].double() - ref).abs().max()) / scale)
    e_cuda, e_ref = np.array(e_cuda), np.array(e_ref)
    worst = names[int(e_cuda.argmax())]
    print(f"gradient error vs fp64 oracle: CUDA median {np.median(e_cuda):.2e} p90 {np.percentile(e_cuda, 90):.2e} max {e_cuda.max():.2e}; "
          f"fp32 CPU oracle median {np.median(e_ref):.2e} p90 {np.percentile(e_ref, 90):.2e} max {e_ref.max():.2e}")
    # measured: the CUDA path sits within ~3-11x of the CPU fp32 implementation's own error on these
    # ill-conditioned sums (long fp32 FMA chains + fp32 atomics vs the CPU's blocked summation)
    assert np.median(e_cuda) < max(5e-4, 20.0 * np.median(e_ref)), (np.median(e_cuda), np.median(e_ref))
    assert np.percentile(e_cuda, 90) < max(2e-3, 20.0 * np.percentile(e_ref, 90)), (np.percentile(e_cuda, 90), np.percentile(e_ref, 90))
    assert e_cuda.max() < max(3e-2, 5.0 * e_ref.max()), (worst, e_cuda.max(), e_ref.max())


def test_forward_backward_vs_oracle():
    cfg = onet.FCDENSENET57
    state, x, model = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 128, 160, 303)
    gy = torch.randn(2, 1, 128, 160, generator=torch.Generator().manual_seed(9))
    y64, g64, buf64 = _oracle_fwd_bwd(state, x, cfg, gy, torch.float64)
    y32, g32, _ = _oracle_fwd_bwd(state, x, cfg, gy, torch.float32)
    y = model(x.cuda())
    assert y.shape == (2, 1, 128, 160) and y.grad_fn is not None and float(y.min()) >= 0.0
    assert rel_err(y, y64) < 1e-4
    (y * gy.cuda()).sum().backward()
    names = [k for k in state if not onet.is_buffer(k)]
    _check_grads(model, g64, g32, names)
    # every p.grad is a view of the flat bucket, in state_dict order
    off = 0
    for k, p in model.named_parameters():
        assert p.grad.data_ptr() == model.flat_grads.data_ptr() + 4 * off, k
        off += p.numel()
    # BN running buffers and counters (momentum 0.1, unbiased variance)
    sd = model.state_dict()
    for k, v in buf64.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v)
        else:
            assert rel_err(sd[k], v) < 1e-5, k


def test_reference_fixture_net_a():
    g = load_golden("net_a")
    b, h, w, seed = [int(v) for v in g["meta"]]
    cfg = onet.FCDENSENET57
    state, x, model = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), b, h, w, seed)
    y = model(x.cuda())
    assert rel_err(y, g["y"]) < 1e-4
    (y * torch.tensor(g["gy"]).cuda()).sum().backward()
    names = [k for k in state if not onet.is_buffer(k)]
    params = dict(model.named_parameters())
    l2 = np.array([params[k].grad.double().norm().item() for k in names])
    assert np.all(np.abs(l2 - g["grad_l2"]) <= 5e-3 * g["grad_l2"] + 1e-5 * g["grad_l2"].max())
    for k in g:
        if k.startswith("grad::"):
            assert rel_err(params[k[6:]].grad, g[k]) < 1e-2, k
        if k.startswith("buf::"):
            assert rel_err(model.state_dict()[k[5:]], g[k]) < 1e-5, k
    model.eval()
    with torch.no_grad():
        y_eval = model(x.cuda())
    assert rel_err(y_eval, g["y_eval"]) < 1e-4


def test_pair_forward_equals_two_calls():
    cfg = onet.FCDENSENET57
    state, x1, model = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 64, 64, 11)
    x2 = torch.flip(x1, dims=[3]).contiguous()
    y1 = model(x1.cuda())
    y2 = model(x2.cuda())
    gy = torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(2)).cuda()
    ((y1 * gy).sum() + (y2 * gy * 0.5).sum()).backward()
    g_two = model.flat_grads.clone()
    buf_two = model._flat_buf.clone()
    model2 = endo_b200.models.FCDenseNet57(n_classes=1)
    model2.load_state_dict(state)
    model2.cuda().train()
    p1, p2 = model2.forward_pair(x1.cuda(), x2.cuda())
    assert rel_err(p1, y1) < 1e-6 and rel_err(p2, y2) < 1e-6
    ((p1 * gy).sum() + (p2 * gy * 0.5).sum()).backward()
    assert rel_err(model2.flat_grads, g_two) < 1e-3
    assert rel_err(model2._flat_buf, buf_two) < 1e-6
    assert int(model2.state_dict()["denseBlocksDown.0.layers.0.norm.num_batches_tracked"]) == 2


def test_growth16_variant_fcdensenet67():
    cfg = onet.FCDENSENET67
    state, x, model = _setup(cfg, lambda: endo_b200.models.FCDenseNet67(n_classes=1), 2, 64, 64, 21)
    # (batch 1 at 32x64 leaves 2 samples per channel in the bottleneck BatchNorms: gradients there are pure cancellation)
    gy = torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(4))
    y64, g64, _ = _oracle_fwd_bwd(state, x, cfg, gy, torch.float64)
    y32, g32, _ = _oracle_fwd_bwd(state, x, cfg, gy, torch.float32)
    y = model(x.cuda())
    assert rel_err(y, y64) < 1e-4
    (y * gy.cuda()).sum().backward()
    _check_grads(model, g64, g32, [k for k in state if not onet.is_buffer(k)])


@pytest.mark.parametrize("math_mode", ["fp32", "tf32x3"])
def test_full_train_step_vs_reference_fixture(math_mode):
    """Two optimisation steps (train.py:272-328) against the trace recorded from the unmodified reference: loss terms,
    gradient norm and updated weights -- the same bounds for the fp32 FFMA path and the tensor-core tf32x3 path."""
    g = load_golden("step_a")
    b, h, w, seed = [int(v) for v in g["meta"]]
    cfg = onet.FCDENSENET57
    state = onet.init_state(cfg, seed=seed, perturb=False)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=0.02)
    cb = {k: v.cuda() for k, v in batch.items()}
    for pair in (False, True):
        model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
        model.load_state_dict(state)
        model.cuda().train()
        step = endo_b200.train_step.TrainStep(model, h, w, lr=1e-3, momentum=0.9, max_norm=10.0, pair=pair)
        for it in range(2):
            loss, dcl, sfl = step.step(cb)
            # (bf16x3 is not held to this trace: its 2e-5 depth error is amplified ~250x by the composite loss at random
            # init and the second step then moves by 5-15 %; its forward is checked in test_tf32x3_*)
            # first step: the composite loss at random init amplifies depth-map differences ~10^3 x (a 3e-6 depth difference of
            # the tensor-core path shows up as 2.3e-3 in the flow term); second step: one realisation of the chaotic update
            tol = (2e-3 if math_mode == "fp32" else 4e-3) if it == 0 else (5e-2 if math_mode == "fp32" else 1e-1)
            assert abs(float(loss) - g["loss"][it]) / g["loss"][it] < tol, (pair, it, float(loss), g["loss"][it])
            assert abs(float(dcl) - g["dcl"][it]) / g["dcl"][it] < tol
            assert abs(float(sfl) - g["sfl"][it]) / g["sfl"][it] < tol
            # the second step's gradient norm is one realisation of fp32 rounding: an A/B of two summation orders of
            # the SAME forward (test_splitk_forward_matches_single_pass: y equal to 1e-6) moves individual gradient
            # tensors by up to 5e-2 at this size, and the norm after one update by ~10%
            # (tensor-core gradients perturb the first update a little more than the fp32 FFMA path: measured up to 28 %)
            assert abs(float(step.opt.grad_norm) - g["gnorm"][it]) / g["gnorm"][it] < \
                (2e-2 if it == 0 else (0.35 if math_mode == "fp32" else 0.5))
        names = [k for k in state if not onet.is_buffer(k)]
        params = dict(model.named_parameters())
        l2 = np.array([params[k].double().norm().item() for k in names])
        assert np.all(np.abs(l2 - g["w_l2_after"]) <= 1e-3 * g["w_l2_after"] + 1e-5)


def test_torch_optimizer_dropin_matches_fused_tail():
    """train.py:202,324-328 with torch.optim.SGD + clip_grad_norm_ on the mirrored module == fused tail."""
    cfg = onet.FCDENSENET57
    state = onet.init_state(cfg, seed=5, perturb=False)
    batch = endo_b200.synthetic.make_batch(2, 64, 64, seed=5, sparse_prob=0.02)
    cb = {k: v.cuda() for k, v in batch.items()}
    m1 = endo_b200.models.FCDenseNet57(1); m1.load_state_dict(state); m1.cuda().train()
    m2 = endo_b200.models.FCDenseNet57(1); m2.load_state_dict(state); m2.cuda().train()
    stack = endo_b200.train_step.LossStack(64, 64)
    opt = torch.optim.SGD(m1.parameters(), lr=1e-3, momentum=0.9)
    fused = endo_b200.train_step.TrainStep(m2, 64, 64, lr=1e-3, pair=False)
    for _ in range(2):
        loss, _, _, _ = stack.loss(m1, cb)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m1.parameters(), 10.0)
        opt.step()
        fused.step(cb)
    assert rel_err(m2.flat_params, m1.flat_params) < 1e-5


def test_shape_errors_are_loud():
    model = endo_b200.models.FCDenseNet57(1).cuda()
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 40, 64, device="cuda"))      # H not a multiple of 32
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 4, 64, 64, device="cuda"))      # wrong channel count


def test_tf32_tensor_core_forward():
    """math="tf32": DenseLayer convolutions on tcgen05 with tf32 operands and fp32 accumulation (the arithmetic
    cuDNN uses for the reference's convs by default, torch.backends.cudnn.allow_tf32=True).  Looser bound than
    the fp32 path: 10-bit operand mantissas through 44 stacked convolutions."""
    cfg = onet.FCDENSENET57
    state, x, _ = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 128, 160, 303)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math="tf32")
    model.load_state_dict(state)
    model.cuda().train()
    gy = torch.randn(2, 1, 128, 160, generator=torch.Generator().manual_seed(9))
    y64, g64, buf64 = _oracle_fwd_bwd(state, x, cfg, gy, torch.float64)
    y = model(x.cuda())
    err = rel_err(y, y64)
    print("tf32 forward rel err", err)
    assert err < 2e-2, err
    (y * gy.cuda()).sum().backward()
    params = dict(model.named_parameters())
    e = rel_err(params["finalConv.weight"].grad, g64["finalConv.weight"])
    assert e < 5e-2, e
    sd = model.state_dict()
    for k in ("denseBlocksDown.0.layers.1.norm.running_mean", "denseBlocksUp.4.layers.3.norm.running_var"):
        assert rel_err(sd[k], buf64[k]) < 2e-2, k


def _grad_errors(model, g64, names):
    gmax = max(float(g64[k].abs().max()) for k in names)
    params = dict(model.named_parameters())
    out = []
    for k in names:
        scale = max(float(g64[k].abs().max()), 1e-5 * gmax)
        out.append(float((params[k].grad.double().cpu() - g64[k]).abs().max()) / scale)
    return np.array(out)


@pytest.mark.parametrize("math_mode", ["tf32x3", "bf16x3"])
def test_tf32x3_tensor_core_forward_is_fp32_grade(math_mode):
    """math="tf32x3" / "bf16x3": every convolution of the forward on tcgen05 with error-compensated operands (x = hi + lo,
    D += lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM).  The depth map, the BatchNorm buffers and the eval-mode
    output must meet the SAME 1e-4 / 1e-5 bounds as the fp32 FFMA path (north_star: depth maps within 1e-4 rel fp32);
    the gradients come from the tf32 / bf16-operand tensor-core kernels and are held to the tensor-core bound.
    bf16x3 splits every operand into two bf16 terms (16 significant bits) instead of two tf32 terms (21 bits)."""
    cfg = onet.FCDENSENET57
    state, x, _ = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 128, 160, 303)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    gy = torch.randn(2, 1, 128, 160, generator=torch.Generator().manual_seed(9))
    y64, g64, buf64 = _oracle_fwd_bwd(state, x, cfg, gy, torch.float64)
    y32, g32, _ = _oracle_fwd_bwd(state, x, cfg, gy, torch.float32)
    y = model(x.cuda())
    err, err32 = rel_err(y, y64), rel_err(y32, y64)
    print(f"{math_mode} forward rel err vs fp64 oracle {err:.3e} (fp32 CPU oracle: {err32:.3e})")
    assert err < 1e-4, err
    (y * gy.cuda()).sum().backward()
    sd = model.state_dict()
    for k, v in buf64.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v)
        else:
            assert rel_err(sd[k], v) < (1e-5 if math_mode == "tf32x3" else 5e-5), k
    names = [k for k in state if not onet.is_buffer(k)]
    errs = _grad_errors(model, g64, names)
    gmax = max(float(g64[k].abs().max()) for k in names)
    e32 = np.array([float((g32[k].double() - g64[k]).abs().max()) / max(float(g64[k].abs().max()), 1e-5 * gmax) for k in names])
    print(f"{math_mode} gradient error vs fp64 oracle: median {np.median(errs):.2e} p90 {np.percentile(errs, 90):.2e} max {errs.max():.2e}; "
          f"fp32 CPU oracle median {np.median(e32):.2e} p90 {np.percentile(e32, 90):.2e} max {e32.max():.2e}")
    assert np.median(errs) < (5e-3 if math_mode == "tf32x3" else 1.5e-2) and errs.max() < 2e-1, (np.median(errs), errs.max())
    model.eval()
    with torch.no_grad():
        y_eval = model(x.cuda())
    y_eval64 = onet.forward({k: (v if v.dtype == torch.long else v.double()) for k, v in {**state, **buf64}.items()},
                            x.double(), cfg, False, {})
    assert rel_err(y_eval, y_eval64) < 1e-4


@pytest.mark.parametrize("math_mode", ["tf32x3", "bf16x3"])
def test_tf32x3_reference_fixture_and_growth16(math_mode):
    """Error-compensated tensor-core forward against the fixture generated from the unmodified reference, and the 16-channel-growth variant."""
    g = load_golden("net_a")
    b, h, w, seed = [int(v) for v in g["meta"]]
    cfg = onet.FCDENSENET57
    state, x, _ = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), b, h, w, seed)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    y = model(x.cuda())
    print(f"{math_mode} forward vs reference fixture: {rel_err(y, g['y']):.3e}")
    assert rel_err(y, g["y"]) < 1e-4, rel_err(y, g["y"])
    for k in g:
        if k.startswith("buf::"):
            assert rel_err(model.state_dict()[k[5:]], g[k]) < (1e-5 if math_mode == "tf32x3" else 5e-5), k
    cfg = onet.FCDENSENET67
    state, x, _ = _setup(cfg, lambda: endo_b200.models.FCDenseNet67(n_classes=1), 2, 64, 64, 21)
    model = endo_b200.models.FCDenseNet67(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    y64 = onet.forward({k: (v if v.dtype == torch.long else v.double()) for k, v in state.items()}, x.double(), cfg, True, {})
    e67 = rel_err(model(x.cuda()), y64)
    print(f"{math_mode} FCDenseNet67 forward vs fp64 oracle: {e67:.3e}")
    assert e67 < 1e-4


def test_tensor_core_backward_kernels_match_ffma_backward(monkeypatch):
    """A/B inside math="tf32": identical tcgen05 forward, then the backward once with the fp32 FFMA kernels
    (ENDO_TC_DISABLE=6) and once with the tcgen05 data-gradient (tf32) and/or weight-gradient (bf16) kernels.
    The two backward passes differentiate the SAME forward, so they must agree to operand-rounding accuracy."""
    cfg = onet.FCDENSENET57
    state, x, _ = _setup(cfg, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 128, 160, 77)
    gy = torch.randn(2, 1, 128, 160, generator=torch.Generator().manual_seed(3)).cuda()

    def grads(mask):
        monkeypatch.setenv("ENDO_TC_DISABLE", str(mask))
        model = endo_b200.models.FCDenseNet57(n_classes=1, math="tf32")
        model.load_state_dict(state)
        model.cuda().train()
        y = model(x.cuda())
        (y * gy).sum().backward()
        torch.cuda.synchronize()
        return {k: p.grad.clone() for k, p in model.named_parameters()}, y.detach().clone()

    # bits: 2 = dense dgrad, 4 = dense wgrad, 16 = TransitionUp wgrad, 32 = TransitionDown wgrad, 1024 = TransitionDown
    # dgrad, 2048 = TransitionUp dgrad fall back to FFMA (1 / 8 / 64 / 512 would change the forward)
    full = 2 + 4 + 16 + 32 + 1024 + 2048
    ref, y_ref = grads(full)
    gmax = max(float(v.abs().max()) for v in ref.values())
    for mask, what in ((full - 2, "dense dgrad"), (full - 4, "dense wgrad"), (full - 16, "TransitionUp wgrad"),
                       (full - 32, "TransitionDown wgrad"), (full - 1024, "TransitionDown dgrad"), (full - 2048, "TransitionUp dgrad"), (0, "all")):
        got, y = grads(mask)
        assert rel_err(y, y_ref) < 1e-6
        errs = []
        for k, v in ref.items():
            scale = max(float(v.abs().max()), 1e-4 * gmax)
            errs.append(float((got[k] - v).abs().max()) / scale)
        errs = np.array(errs)
        print(f"tensor-core {what}: median {np.median(errs):.2e}  p90 {np.percentile(errs, 90):.2e}  max {errs.max():.2e}")
        assert np.median(errs) < 5e-3, (what, np.median(errs))
        assert errs.max() < 1e-1, (what, errs.max(), list(ref)[int(errs.argmax())])


def test_gemm_weight_gradients_and_fused_pool_match_round2a_kernels(monkeypatch):
    """A/B inside math="tf32x3" at 256x320: the TMA -> tcgen05 GEMM weight gradients over the bf16 by-product planes
    (DenseLayers / TransitionUp / first convolution: ENDO_TC_DISABLE bit 262144; TransitionDown: 131072) and the max-pool fused
    into the 1x1 GEMM epilogue (65536) against the round-2a kernels they replace.
    * GEMM weight gradients: both sides round the same operands to bf16 with the same instruction, the forward is bit-identical,
      so the gradients may differ by the fp32 summation order only -- held to the run-to-run noise floor of the reference
      kernels themselves (measured on B200: median 1.2e-7, p90 6e-6, max 7e-4 for the same kernels run twice; GEMMs: 1.3e-7 /
      6e-6 / 1.9e-3).
    * fused max-pool: pooled values and argmax are bit-identical, only the order of the partial sums of the pooled map's
      statistics changes: BatchNorm buffers 1.2e-7, forward 6e-7.  This raw-Kaiming bs2 network amplifies that into 2e-3 (median)
      of the gradients (ReLU / argmax flips feeding BatchNorms, see _check_grads and tools/chaos_probe.py), so the gradients are a
      loose routing guard here; the tight gradient bounds on the fused path are tests/test_gpu_parity_c2.py's."""
    state, x, _ = _setup(onet.FCDENSENET57, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 256, 320, 79)
    gy = torch.randn(2, 1, 256, 320, generator=torch.Generator().manual_seed(5)).cuda()

    def run(mask):
        monkeypatch.setenv("ENDO_TC_DISABLE", str(mask))
        model = endo_b200.models.FCDenseNet57(n_classes=1, math="tf32x3")
        model.load_state_dict(state)
        model.cuda().train()
        y = model(x.cuda())
        (y * gy).sum().backward()
        torch.cuda.synchronize()
        bufs = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k}
        return {k: p.grad.clone() for k, p in model.named_parameters()}, y.detach().clone(), bufs

    old = 65536 + 131072 + 262144
    ref, y_ref, b_ref = run(old)
    gmax = max(float(v.abs().max()) for v in ref.values())
    for mask, what, tight in ((old, "noise floor (same kernels again)", True), (old - 131072, "TransitionDown GEMM", True),
                              (old - 262144, "DenseLayer / TransitionUp / first-conv GEMM", True),
                              (old - 65536, "fused max-pool", False), (0, "all", False)):
        got, y, bufs = run(mask)
        fwd = rel_err(y, y_ref)
        bn = max(rel_err(bufs[k], b_ref[k]) for k in bufs)
        errs = np.array([float((got[k] - v).abs().max()) / max(float(v.abs().max()), 1e-4 * gmax) for k, v in ref.items()])
        print(f"{what}: forward {fwd:.2e}, BatchNorm buffers {bn:.2e}; gradients median {np.median(errs):.2e}  "
              f"p90 {np.percentile(errs, 90):.2e}  max {errs.max():.2e}")
        if tight:
            assert fwd == 0.0 and bn == 0.0, (what, fwd, bn)
            assert np.median(errs) < 2e-6 and np.percentile(errs, 90) < 1e-4 and errs.max() < 1e-2, (what, np.median(errs), errs.max())
        else:
            assert fwd < 1e-5 and bn < 1e-5, (what, fwd, bn)
            assert np.median(errs) < 1e-2 and errs.max() < 2e-1, (what, np.median(errs), errs.max(), list(ref)[int(errs.argmax())])


def test_tensor_core_transition_down_forward_matches_ffma(monkeypatch):
    """math="tf32" with the TransitionDown 1x1 convolution on tcgen05 (+ pooling pass) against the same forward
    with that one layer type on the fp32 FFMA kernel (ENDO_TC_DISABLE=64): tf32 operand rounding only.  The
    gradients of the two runs differentiate forwards that differ by ~1e-3, which this network amplifies ~100x
    (ReLU / arg-max flips feeding BatchNorms with few samples per channel, see _check_grads): they are compared
    loosely, as a guard against routing errors (a wrong arg-max byte or statistic gives O(1) differences)."""
    state, x, _ = _setup(onet.FCDENSENET57, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 128, 160, 78)
    gy = torch.randn(2, 1, 128, 160, generator=torch.Generator().manual_seed(4)).cuda()

    def run(mask):
        monkeypatch.setenv("ENDO_TC_DISABLE", str(mask))
        model = endo_b200.models.FCDenseNet57(n_classes=1, math="tf32")
        model.load_state_dict(state)
        model.cuda().train()
        y = model(x.cuda())
        (y * gy).sum().backward()
        torch.cuda.synchronize()
        return y.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()}

    y_ref, g_ref = run(64)
    y, g = run(0)
    assert rel_err(y, y_ref) < 5e-3, rel_err(y, y_ref)
    gmax = max(float(v.abs().max()) for v in g_ref.values())
    errs = np.array([float((g[k] - v).abs().max()) / max(float(v.abs().max()), 1e-4 * gmax) for k, v in g_ref.items()])
    print(f"TransitionDown tcgen05 forward: y {rel_err(y, y_ref):.2e}  grads median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 0.25, np.median(errs)


def test_splitk_forward_matches_single_pass(monkeypatch):
    """fp32 FFMA DenseLayer forward at low resolution: input channels split over CTAs + fixed-order finish kernel
    against the single-pass kernel (ENDO_TC_DISABLE=128).  Same products, different fp32 summation order."""
    state, x, _ = _setup(onet.FCDENSENET57, lambda: endo_b200.models.FCDenseNet57(n_classes=1), 2, 64, 64, 404)
    gy = torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(6)).cuda()

    def run(mask):
        monkeypatch.setenv("ENDO_TC_DISABLE", str(mask))
        model = endo_b200.models.FCDenseNet57(n_classes=1)
        model.load_state_dict(state)
        model.cuda().train()
        y = model(x.cuda())
        (y * gy).sum().backward()
        torch.cuda.synchronize()
        bufs = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}
        return y.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()}, bufs

    y_ref, g_ref, b_ref = run(128)
    y, g, b = run(0)
    assert rel_err(y, y_ref) < 1e-5, rel_err(y, y_ref)
    for k in b_ref:
        assert rel_err(b[k], b_ref[k]) < 1e-5, k
    gmax = max(float(v.abs().max()) for v in g_ref.values())
    errs = np.array([float((g[k] - v).abs().max()) / max(float(v.abs().max()), 1e-4 * gmax) for k, v in g_ref.items()])
    print(f"split-K forward: y {rel_err(y, y_ref):.2e}  grads median {np.median(errs):.2e} p90 {np.percentile(errs, 90):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 1e-3, np.median(errs)

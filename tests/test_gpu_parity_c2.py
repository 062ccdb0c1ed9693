"""GPU parity AT THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1]: bs8 256x320, FCDenseNet57, full loss
stack dcl 5 / sfl 20) and at the tolerance north_star states: loss scalars and depth maps within 1e-4.

Two kinds of evidence:
  * the well-conditioned step fixtures `step_b` (2x64x96) and `step_c` (8x256x320, the bench shape and seed),
    recorded from the UNMODIFIED reference by oracle/gen_golden.py (oracle.net.condition_state keeps the predicted
    depth away from zero, so that the composite loss does not amplify 1e-7 depth differences: fp32 and fp64
    oracles agree to <1e-6 on every loss term there, the gradient norm is O(10));
  * the network alone at 256x320 against the CPU oracle: forward of 16 images through forward_pair (groups = 2)
    and through two separate calls (fp32 oracle, 1e-4), forward + backward at bs2 (fp64 oracle, distribution bounds).

Gradient yardstick: two fp32 CPU implementations of this step (the oracle's written-out BatchNorm and the
reference's ATen BatchNorm) differ per tensor by up to 2.7e-3 of the tensor's max at bs8 256x320
(tests/test_oracle_golden.py::test_full_step_benchmark_configuration), 3e-5 in the median tensor norm."""
import numpy as np
import pytest
import torch

import endo_b200
from oracle import net as onet
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-4          # north_star: loss scalars within 1e-4 rel fp32
DEPTH_TOL = 1e-4         # north_star: depth maps within 1e-4 rel fp32


def _fixture_setup(tag):
    g = load_golden(tag)
    b, h, w, seed, stride = [int(v) for v in g["meta"]]
    cfg = onet.FCDENSENET57
    state = onet.condition_state(onet.init_state(cfg, seed=seed, perturb=(tag == "step_b")))
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=0.02 if tag == "step_b" else 0.005)
    return g, (b, h, w, stride), state, {k: v.cuda() for k, v in batch.items()}


# per-tensor gradient bounds (max |diff| / max |ref| per tensor; distribution over the 210 tensors), by math mode.
# fp32: the fp32 FFMA path must sit at the CPU-vs-CPU yardstick.  tf32x3: forward fp32-grade, data gradient with tf32
# operands, weight gradient with bf16 operands (fp32 accumulate) -- what cuDNN's default TF32 convolutions give the
# reference on a GPU, see bench.py --impl reference-gpu.
# Measured on B200 (profiles/r2_parity.md), bs8 256x320: fp32 4e-5 / 3e-4 / 2.4e-3, norm 8e-7; tf32x3 7e-5 / 4e-4 / 2.4e-3,
# norm 5e-6; single tensors up to 3.8e-3 (a BatchNorm gamma gradient) -- the bounds leave ~4x.
GRAD_BOUNDS = {"fp32": dict(median=2e-4, p90=1.5e-3, max=1.5e-2, norm=1e-4),
               "tf32x3": dict(median=3e-4, p90=2e-3, max=1.5e-2, norm=1e-4)}


@pytest.mark.parametrize("tag", ["step_b", "step_c"])
@pytest.mark.parametrize("math_mode", ["fp32", "tf32x3"])
@pytest.mark.parametrize("pair", [False, True])
def test_two_steps_vs_reference_trace(tag, math_mode, pair):
    """train.py:272-328, two iterations, against the trace of the unmodified reference: every loss term within 1e-4
    in BOTH iterations, depth / scaled / warped / flow maps within 1e-4, gradient norm, per-tensor gradient norms,
    weights after two clipped SGD-momentum updates within 1e-4, BatchNorm buffers within 1e-5."""
    g, (b, h, w, stride), state, cb = _fixture_setup(tag)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    step = endo_b200.train_step.TrainStep(model, h, w, lr=1e-3, momentum=0.9, max_norm=10.0, pair=pair)
    names = [k for k in state if not onet.is_buffer(k)]
    gb = GRAD_BOUNDS[math_mode]
    sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
    for it in range(2):
        for p in model.parameters():
            p.grad = None
        loss, dcl, sfl, ex = step.stack.loss(model, cb, pair=pair)
        loss.backward()
        for name, got in (("loss", loss), ("dcl", dcl), ("sfl", sfl)):
            e = abs(float(got) - g[name][it]) / g[name][it]
            print(f"{tag} {math_mode} pair={pair} it={it} {name}: {float(got):.7f} vs {g[name][it]:.7f} rel {e:.2e}")
            assert e < LOSS_TOL, (tag, math_mode, pair, it, name, float(got), g[name][it])
        if it == 0:
            assert rel_err(ex["depth_1"][sub], g["p1"]) < DEPTH_TOL
            assert rel_err(ex["depth_2"][sub], g["p2"]) < DEPTH_TOL
            assert rel_err(ex["scaled_1"][sub], g["s1"]) < DEPTH_TOL
            assert rel_err(ex["warped_2to1"][sub], g["w21"]) < DEPTH_TOL
            assert rel_err(ex["flow_1"][sub], g["f1"] * cb["boundaries"][sub].cpu().numpy()) < DEPTH_TOL
            # the mask thresholds a bilinear sum of the boundary mask at sampling positions that depend on the PREDICTED
            # depth: bit-exact given identical depths (tests/test_gpu_geometry.py); through the network a pixel whose sum
            # is within rounding of 0.9 may flip
            flips = int((ex["inter_1"][sub].cpu().numpy() != g["i1"]).sum())
            assert flips <= max(1, g["i1"].size // 100000), flips
            params = dict(model.named_parameters())
            l2 = np.array([params[k].grad.double().norm().item() for k in names])
            rn = np.abs(l2 - g["grad_l2"]) / (g["grad_l2"] + 1e-5 * g["grad_l2"].max())
            errs = []
            for k in g:
                if k.startswith("grad::"):
                    errs.append((k[6:], rel_err(params[k[6:]].grad, g[k])))
            print(f"{tag} {math_mode} pair={pair}: per-tensor gradient-norm rel err median {np.median(rn):.2e} "
                  f"p90 {np.percentile(rn, 90):.2e} max {rn.max():.2e} ({names[int(rn.argmax())]}); tensors {errs}")
            assert np.median(rn) < gb["median"], np.median(rn)
            assert np.percentile(rn, 90) < gb["p90"], np.percentile(rn, 90)
            assert rn.max() < gb["max"], (names[int(rn.argmax())], rn.max())
            for k, e in errs:
                assert e < gb["max"], (k, e)
        finite = torch.isfinite(loss.detach()).to(torch.float32).reshape(1)
        step.opt.step(finite_flag=finite)
        e = abs(float(step.opt.grad_norm) - g["gnorm"][it]) / g["gnorm"][it]
        print(f"{tag} {math_mode} pair={pair} it={it} gnorm rel {e:.2e}")
        assert e < gb["norm"], (it, float(step.opt.grad_norm), g["gnorm"][it])
    params = dict(model.named_parameters())
    l2 = np.array([params[k].double().norm().item() for k in names])
    assert np.all(np.abs(l2 - g["w_l2_after"]) <= 1e-4 * g["w_l2_after"] + 1e-6)
    sd = model.state_dict()
    for k in g:
        if k.startswith("after::"):
            assert rel_err(params[k[7:]], g[k]) < 1e-4, k
        if k.startswith("buf::"):
            assert rel_err(sd[k[5:]], g[k]) < 1e-5, k


@pytest.mark.parametrize("math_mode", ["fp32", "tf32x3"])
def test_network_forward_bs8_256x320_pair_and_separate(math_mode):
    """The exact launch configuration bench.py times: 16 images of 256x320 through forward_pair (groups = 2: 1280-CTA
    full-resolution grids, the split-K / low-resolution switch points of this size) against the fp32 CPU oracle, and
    against two separate net() calls."""
    cfg = onet.FCDENSENET57
    b, h, w, seed = 8, 256, 320, 10085
    state = onet.init_state(cfg, seed=seed, perturb=True)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    x1 = batch["boundaries"] * batch["colors_1"]
    x2 = batch["boundaries"] * batch["colors_2"]
    new_buf = {}
    with torch.no_grad():
        y1_ref = onet.forward(state, x1, cfg, True, new_buf)
        y2_ref = onet.forward(state, x2, cfg, True, new_buf)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    with torch.no_grad():
        p1, p2 = model.forward_pair(x1.cuda(), x2.cuda())
    e1, e2 = rel_err(p1, y1_ref), rel_err(p2, y2_ref)
    print(f"{math_mode} forward_pair bs8 256x320 vs fp32 oracle: {e1:.2e} {e2:.2e}")
    assert e1 < DEPTH_TOL and e2 < DEPTH_TOL
    sd = model.state_dict()
    for k, v in new_buf.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v) == 2
        else:
            assert rel_err(sd[k], v) < 1e-5, k
    model2 = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model2.load_state_dict(state)
    model2.cuda().train()
    with torch.no_grad():
        q1 = model2(x1.cuda())
        q2 = model2(x2.cuda())
    assert rel_err(q1, y1_ref) < DEPTH_TOL and rel_err(q2, y2_ref) < DEPTH_TOL
    assert rel_err(q1, p1) < 1e-5 and rel_err(q2, p2) < 1e-5
    assert rel_err(model2._flat_buf, model._flat_buf) < 1e-5


@pytest.mark.parametrize("math_mode", ["fp32", "tf32x3"])
def test_network_forward_backward_bs2_256x320_vs_fp64(math_mode):
    """Forward + backward at the benchmark resolution against the fp64 oracle, gradient errors judged as a distribution
    with the fp32 CPU oracle's own error as the yardstick (tests/test_gpu_net.py::_check_grads)."""
    cfg = onet.FCDENSENET57
    b, h, w, seed = 2, 256, 320, 777
    state = onet.init_state(cfg, seed=seed, perturb=True)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    x = batch["boundaries"] * batch["colors_1"]
    gy = torch.randn(b, 1, h, w, generator=torch.Generator().manual_seed(9))

    def oracle(dtype):
        params = {}
        for k, v in state.items():
            v = v if v.dtype == torch.long else v.to(dtype)
            params[k] = v if onet.is_buffer(k) else v.clone().requires_grad_(True)
        y = onet.forward(params, x.to(dtype), cfg, True, {})
        (y * gy.to(dtype)).sum().backward()
        return y.detach(), {k: p.grad for k, p in params.items() if not onet.is_buffer(k)}

    y64, g64 = oracle(torch.float64)
    y32, g32 = oracle(torch.float32)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    y = model(x.cuda())
    assert rel_err(y, y64) < DEPTH_TOL, rel_err(y, y64)
    (y * gy.cuda()).sum().backward()
    names = [k for k in state if not onet.is_buffer(k)]
    gmax = max(float(g64[k].abs().max()) for k in names)
    params = dict(model.named_parameters())
    e_cuda, e_ref = [], []
    for k in names:
        scale = max(float(g64[k].abs().max()), 1e-5 * gmax)
        e_cuda.append(float((params[k].grad.double().cpu() - g64[k]).abs().max()) / scale)
        e_ref.append(float((g32[k].double() - g64[k]).abs().max()) / scale)
    e_cuda, e_ref = np.array(e_cuda), np.array(e_ref)
    print(f"{math_mode} bs2 256x320 gradient error vs fp64 oracle: CUDA median {np.median(e_cuda):.2e} p90 {np.percentile(e_cuda, 90):.2e} "
          f"max {e_cuda.max():.2e} ({names[int(e_cuda.argmax())]}); fp32 CPU oracle median {np.median(e_ref):.2e} "
          f"p90 {np.percentile(e_ref, 90):.2e} max {e_ref.max():.2e}")
    if math_mode == "fp32":
        assert np.median(e_cuda) < max(5e-4, 20.0 * np.median(e_ref))
        assert np.percentile(e_cuda, 90) < max(2e-3, 20.0 * np.percentile(e_ref, 90))
        assert e_cuda.max() < max(3e-2, 5.0 * e_ref.max())
    else:
        assert np.median(e_cuda) < 5e-3 and e_cuda.max() < 1e-1


@pytest.mark.parametrize("math_mode", ["tf32x3", "bf16x3"])
def test_run_to_run_reproducibility_bs8_256x320(math_mode):
    """The benchmarked step twice from identical state.  The forward has no atomics on fp32 data (per-channel BatchNorm
    sums are fp64 atomics of per-CTA fp32 partials: order-dependent only below 1e-16 relative) so depth maps must repeat
    bit for bit up to that; parameter gradients accumulate with fp32 atomics (weight-gradient tiles, 4-tap scatter of the
    warp backward -- the one the reference has too, models.py:546) and may differ by summation order only.  A NaN or a
    large difference here means a race or an uninitialised read, not rounding."""
    cfg = onet.FCDENSENET57
    b, h, w, seed = 8, 256, 320, 10085
    state = onet.condition_state(onet.init_state(cfg, seed=seed, perturb=False))
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    cb = {k: v.cuda() for k, v in batch.items()}
    runs = []
    for _ in range(3):
        model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
        model.load_state_dict(state)
        model.cuda().train()
        stack = endo_b200.train_step.LossStack(h, w)
        loss, dcl, sfl, ex = stack.forward_backward(model, cb, pair=True)
        torch.cuda.synchronize()
        assert torch.isfinite(loss), float(loss)
        assert bool(torch.isfinite(model.flat_grads).all())
        runs.append((float(loss), ex["depth_1"].detach().clone(), ex["depth_2"].detach().clone(), model.flat_grads.clone(),
                     model._flat_buf.clone()))
    l0, d1, d2, g0, b0 = runs[0]
    gmax = float(g0.abs().max())
    for l, e1, e2, g, bb in runs[1:]:
        # depth maps: bit-identical up to the fp64 BatchNorm-sum order (observed: identical)
        assert rel_err(e1, d1) < 1e-6 and rel_err(e2, d2) < 1e-6, (rel_err(e1, d1), rel_err(e2, d2))
        assert rel_err(bb, b0) < 1e-6
        assert abs(l - l0) / abs(l0) < 1e-5, (l, l0)
        gd = float((g - g0).abs().max()) / gmax
        print(f"{math_mode} run-to-run: loss {l:.7f} vs {l0:.7f}; depth {rel_err(e1, d1):.1e}; grads max diff / max {gd:.2e}")
        assert gd < 2e-3, gd


# BASELINE.json configs[2]: "1xB200 bs32 256x320, bf16 tensor-core conv path, warp layers fp32, loss parity vs reference".
# Stated tolerances for the reduced-precision path (operands of the 3x3 convolutions rounded to bf16 = 8 significant bits,
# fp32 accumulate / BatchNorm statistics / master weights; geometric layers and losses fp32): depth maps 1e-2 of their
# scale, every loss term 1e-3 relative.  The fp32-grade tensor-core path must meet 1e-4 at this size as well.
C3_TOL = {"bf16": dict(depth=1e-2, loss=1e-3), "tf32x3": dict(depth=DEPTH_TOL, loss=LOSS_TOL)}   # measured: 1.7e-3 / 7.7e-5


@pytest.mark.parametrize("math_mode", ["bf16", "tf32x3"])
def test_config3_bs32_loss_parity(math_mode):
    g = load_golden("step_d")
    b, h, w, seed, stride = [int(v) for v in g["meta"]]
    assert (b, h, w) == (32, 256, 320)
    cfg = onet.FCDENSENET57
    state = onet.condition_state(onet.init_state(cfg, seed=seed, perturb=False))
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=0.005)
    cb = {k: v.cuda() for k, v in batch.items()}
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    stack = endo_b200.train_step.LossStack(h, w)
    loss, dcl, sfl, ex = stack.forward_backward(model, cb, pair=True)
    tol = C3_TOL[math_mode]
    sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
    e_depth = max(rel_err(ex["depth_1"][sub], g["p1"]), rel_err(ex["depth_2"][sub], g["p2"]))
    errs = {n: abs(float(v) - g[n][0]) / g[n][0] for n, v in (("loss", loss), ("dcl", dcl), ("sfl", sfl))}
    print(f"config 3 (bs32 256x320) {math_mode}: depth rel err {e_depth:.2e}; loss terms {errs}")
    assert e_depth < tol["depth"], e_depth
    assert max(errs.values()) < tol["loss"], errs
    assert bool(torch.isfinite(model.flat_grads).all())      # (the bs32 fixture is forward-only: 46 GB of autograd state on the CPU)


def test_cuda_graph_step_equals_eager_step():
    """GraphedTrainStep (whole step captured once, replayed; device-scalar learning rate) against the eager TrainStep:
    same loss trajectory and weights, also across a learning-rate change between replays (train.py:203 CyclicLR)."""
    g, (b, h, w, stride), state, cb = _fixture_setup("step_b")

    def fresh():
        m = endo_b200.models.FCDenseNet57(n_classes=1, math="tf32x3")
        m.load_state_dict(state)
        return m.cuda().train()

    m_e, m_g = fresh(), fresh()
    eager = endo_b200.train_step.TrainStep(m_e, h, w, lr=1e-3, pair=True)
    # capture WITHOUT consuming optimisation steps: warm-up steps would move the weights, so capture on a scratch copy of the
    # state and restore it afterwards
    graphed = endo_b200.train_step.GraphedTrainStep(m_g, h, w, cb, lr=1e-3, pair=True)
    m_g.load_state_dict(state)
    graphed.inner.opt.buf.zero_()
    losses_e, losses_g = [], []
    for it in range(3):
        lr = 1e-3 if it < 2 else 5e-4
        eager.opt.lr = lr
        graphed.set_lr(lr)
        le, _, _ = eager.step(cb)
        lg, _, _ = graphed.step(cb)
        losses_e.append(float(le)); losses_g.append(float(lg))
    print("eager", losses_e, "graph", losses_g, "launches per replay", graphed.launches_per_step)
    for a, c in zip(losses_e, losses_g):
        assert abs(a - c) / abs(a) < 1e-5, (losses_e, losses_g)
    assert abs(losses_g[0] - g["loss"][0]) / g["loss"][0] < LOSS_TOL
    assert abs(losses_g[1] - g["loss"][1]) / g["loss"][1] < LOSS_TOL
    assert rel_err(m_g.flat_params, m_e.flat_params) < 1e-5
    assert rel_err(m_g._flat_buf, m_e._flat_buf) < 1e-5
    assert int(m_g.state_dict()["denseBlocksDown.0.layers.0.norm.num_batches_tracked"]) == \
        int(m_e.state_dict()["denseBlocksDown.0.layers.0.norm.num_batches_tracked"]) == 6
    # host-batch path: prefetch / swap_in / replay
    host = {k: v.cpu().pin_memory() for k, v in cb.items() if k in graphed.keys}
    graphed.prefetch(host)
    graphed.swap_in()
    l4, _, _ = graphed.replay()
    l4e, _, _ = eager.step(cb)
    assert abs(float(l4) - float(l4e)) / abs(float(l4e)) < 1e-5


def test_successive_graphed_steps_in_one_process():
    """Two GraphedTrainStep objects one after the other (the first deleted, cache emptied), in the two-graph shape the
    multi-GPU path uses: round 2 found a crash here on N > 1 GPUs (a cached workspace living in the first graph's pool)."""
    import gc
    g, (b, h, w, stride), state, cb = _fixture_setup("step_b")
    losses = []
    for _ in range(2):
        m = endo_b200.models.FCDenseNet57(n_classes=1, math="tf32x3")
        m.load_state_dict(state)
        m.cuda().train()
        ts = endo_b200.train_step.GraphedTrainStep(m, h, w, cb, lr=1e-3, pair=True, split_graphs=True)
        m.load_state_dict(state)
        ts.inner.opt.buf.zero_()
        l0 = float(ts.replay()[0])            # (the returned tensors are the graph's static outputs: read before the next replay)
        l1 = float(ts.replay()[0])
        torch.cuda.synchronize()
        losses.append((l0, l1))
        del ts, m
        gc.collect()
        torch.cuda.empty_cache()
    assert abs(losses[0][0] - g["loss"][0]) / g["loss"][0] < LOSS_TOL and abs(losses[0][1] - g["loss"][1]) / g["loss"][1] < LOSS_TOL
    assert abs(losses[1][0] - losses[0][0]) < 1e-5 and abs(losses[1][1] - losses[0][1]) < 1e-5, losses


@pytest.mark.parametrize("math_mode", ["tf32x3"])
def test_config5_network_forward_512x640(math_mode):
    """BASELINE config 5 resolution (512x640, downsampling 2.0): forward of the pair network against the fp32 CPU oracle
    (the persistent TMA-fed kernels then run levels 0-2, 5,120 tiles per image pair at level 0)."""
    cfg = onet.FCDENSENET57
    b, h, w, seed = 1, 512, 640, 50085
    state = onet.condition_state(onet.init_state(cfg, seed=seed, perturb=True))
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    x1 = batch["boundaries"] * batch["colors_1"]
    x2 = batch["boundaries"] * batch["colors_2"]
    with torch.no_grad():
        y1_ref = onet.forward(state, x1, cfg, True, {})
        y2_ref = onet.forward(state, x2, cfg, True, {})
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    p1, p2 = model.forward_pair(x1.cuda(), x2.cuda())
    e1, e2 = rel_err(p1, y1_ref), rel_err(p2, y2_ref)
    print(f"{math_mode} forward_pair 512x640 vs fp32 oracle: {e1:.2e} {e2:.2e}")
    assert e1 < DEPTH_TOL and e2 < DEPTH_TOL
    (p1.sum() + p2.sum()).backward()
    assert bool(torch.isfinite(model.flat_grads).all())

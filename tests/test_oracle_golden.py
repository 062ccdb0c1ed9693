"""Pin the oracle (oracle/) to fixtures produced by the unmodified reference
(oracle/gen_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

import endo_b200
from oracle import geometry, losses, net, step
from conftest import load_golden, rel_err

TOL = 2e-5   # fp32 restatement vs fp32 reference: different op order only


def _geo_inputs(meta):
    b, h, w, seed, ones = [int(v) for v in meta]
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, all_ones_boundary=bool(ones), sparse_prob=0.02)
    d1, d2 = endo_b200.synthetic.jitter_depths(batch, seed=seed + 1)
    return batch, d1.clone().requires_grad_(True), d2.clone().requires_grad_(True)


@pytest.mark.parametrize("tag", ["geo_a", "geo_b"])
def test_depth_scaling(tag):
    g = load_golden(tag)
    batch, d1, _ = _geo_inputs(g["meta"])
    out, std = geometry.depth_scaling(d1, batch["sparse_depths_1"], batch["sparse_depth_masks_1"])
    assert rel_err(out, g["scale_out"]) < TOL
    assert rel_err(std, g["scale_std"]) < 1e-4
    (gd,) = torch.autograd.grad((out * torch.tensor(g["scale_gout"])).sum(), d1)
    assert rel_err(gd, g["scale_gd"]) < TOL


@pytest.mark.parametrize("tag", ["geo_a", "geo_b"])
def test_flow_from_depth(tag):
    g = load_golden(tag)
    batch, d1, _ = _geo_inputs(g["meta"])
    f = geometry.flow_from_depth(d1, batch["boundaries"], batch["translations_1_wrt_2"],
                                 batch["rotations_1_wrt_2"], batch["intrinsics"])
    assert rel_err(f, g["flow_out"]) < TOL
    (gd,) = torch.autograd.grad((f * torch.tensor(g["flow_gout"])).sum(), d1)
    assert rel_err(gd, g["flow_gd"]) < TOL


@pytest.mark.parametrize("tag", ["geo_a", "geo_b"])
def test_depth_warping(tag):
    g = load_golden(tag)
    batch, d1, d2 = _geo_inputs(g["meta"])
    wd, inter = geometry.depth_warping(d1, d2, batch["boundaries"], batch["translations_1_wrt_2"],
                                       batch["rotations_1_wrt_2"], batch["intrinsics"])
    assert rel_err(wd, g["warp_out"]) < TOL
    assert int((inter.numpy() != g["warp_inter"]).sum()) == 0          # bit-exact mask
    gd1, gd2 = torch.autograd.grad((wd * torch.tensor(g["warp_gout"])).sum(), [d1, d2])
    assert rel_err(gd1, g["warp_gd1"]) < 1e-4
    assert rel_err(gd2, g["warp_gd2"]) < TOL


@pytest.mark.parametrize("tag", ["geo_a", "geo_b"])
def test_losses(tag):
    g = load_golden(tag)
    batch, d1, d2 = _geo_inputs(g["meta"])
    bound = batch["boundaries"]
    f1 = torch.tensor(g["flow_out"]).requires_grad_(True)
    lv = losses.sparse_masked_l1_loss(batch["sparse_flows_1"] * bound, f1 * bound,
                                      batch["sparse_flow_masks_1"] * bound)
    assert rel_err(lv, g["l1_out"]) < TOL
    (gf,) = torch.autograd.grad(lv, f1)
    assert rel_err(gf, g["l1_gflow"]) < TOL

    wd = torch.tensor(g["warp_out"]).requires_grad_(True)
    inter = torch.tensor(g["warp_inter"])
    nv = losses.normalized_distance_loss(d1, wd, inter, batch["intrinsics"])
    assert rel_err(nv, g["ndl_out"]) < TOL
    gd, gw = torch.autograd.grad(nv, [d1, wd])
    assert rel_err(gd, g["ndl_gd"]) < TOL
    assert rel_err(gw, g["ndl_gw"]) < TOL

    sv = losses.scale_invariant_loss(d1, d2, bound)
    assert rel_err(sv, g["sil_out"]) < TOL
    gp, gg = torch.autograd.grad(sv, [d1, d2])
    assert rel_err(gp, g["sil_gp"]) < TOL
    assert rel_err(gg, g["sil_gg"]) < TOL


def test_param_names_match_reference_layout():
    shapes = net.param_shapes(net.FCDENSENET57)
    n_param = sum(int(np.prod(s)) for k, s in shapes.items() if not net.is_buffer(k))
    n_tensors = sum(1 for k in shapes if not net.is_buffer(k))
    n_buf = sum(int(np.prod(s)) if s else 1 for k, s in shapes.items() if net.is_buffer(k))
    assert n_param == 1374865 and n_tensors == 210 and n_buf == 21217
    assert abs(net.conv_flops_per_image(net.FCDENSENET57, 256, 320) / 1e9 - 32.196) < 0.01


def test_network_forward_backward():
    g = load_golden("net_a")
    b, h, w, seed = [int(v) for v in g["meta"]]
    cfg = net.FCDENSENET57
    state = net.init_state(cfg, seed=seed, perturb=True)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    x = batch["boundaries"] * batch["colors_1"]
    params = {k: (v if net.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in state.items()}
    new_buf = {}
    y = net.forward(params, x, cfg, True, new_buf)
    assert rel_err(y, g["y"]) < 1e-4
    (y * torch.tensor(g["gy"])).sum().backward()
    names = [k for k in params if not net.is_buffer(k)]
    l2 = np.array([params[k].grad.double().norm().item() for k in names])
    # conv biases that feed a BatchNorm have a mathematically zero gradient (pure rounding noise in
    # fp32), hence the absolute floor relative to the largest gradient norm
    assert np.all(np.abs(l2 - g["grad_l2"]) <= 2e-3 * g["grad_l2"] + 1e-6 * g["grad_l2"].max())
    for k in g:
        if k.startswith("grad::"):
            assert rel_err(params[k[6:]].grad, g[k]) < 1e-2, k   # fp32 backward noise of two fp32 codes (fp64 oracle agrees to 1e-5)
        if k.startswith("buf::"):
            assert rel_err(new_buf[k[5:]], g[k]) < 1e-5, k
    assert int(new_buf["denseBlocksDown.0.layers.0.norm.num_batches_tracked"]) == int(g["num_batches_tracked"])
    with torch.no_grad():
        y_eval = net.forward({**state, **new_buf}, x, cfg, False)
    assert rel_err(y_eval, g["y_eval"]) < 1e-4


def test_full_step():
    g = load_golden("step_a")
    b, h, w, seed = [int(v) for v in g["meta"]]
    cfg = net.FCDENSENET57
    state = net.init_state(cfg, seed=seed, perturb=False)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=0.02)
    mom = {}
    for it in range(2):
        loss, dcl, sfl, grads, new_buf, ex = step.forward_backward(state, batch, cfg, 5.0, 20.0)
        if it == 0:
            assert rel_err(ex["depth_1"], g["p1"]) < 1e-4
            # the scale is a sum of sparse_depth / predicted_depth terms; at random init abs(net) crosses
            # zero, so 1e-7 differences in the prediction are amplified (ill-conditioned, not a defect)
            assert rel_err(ex["scaled_1"], g["s1"]) < 2e-3
            assert rel_err(ex["warped_2to1"], g["w21"]) < 2e-2
            assert int((ex["inter_1"].numpy() != g["i1"]).sum()) <= 2
        # the second iteration runs on weights updated with a gradient of norm 1.3e5 clipped to 10:
        # fp32 noise of the first backward is visible there, hence the looser bound
        tol = 2e-3 if it == 0 else 3e-2
        assert abs(float(loss) - g["loss"][it]) / g["loss"][it] < tol, (it, float(loss), g["loss"][it])
        assert abs(float(dcl) - g["dcl"][it]) / g["dcl"][it] < tol
        assert abs(float(sfl) - g["sfl"][it]) / g["sfl"][it] < tol
        gn = step.clip_and_sgd(state, grads, mom, lr=1e-3)
        assert abs(float(gn) - g["gnorm"][it]) / g["gnorm"][it] < 2e-2
        state.update(new_buf)
    names = [k for k in state if not net.is_buffer(k)]
    l2 = np.array([state[k].double().norm().item() for k in names])
    assert np.all(np.abs(l2 - g["w_l2_after"]) <= 1e-3 * g["w_l2_after"] + 1e-5)   # biases start at 0


def test_library_ops_mode_matches_restatement():
    """bench.py times the oracle with LIBRARY_OPS=True (F.batch_norm / F.interpolate / F.grid_sample, the ops
    the reference itself calls); that mode must be the same function as the written-out restatement."""
    from oracle import net as onet, geometry as ogeo
    g = load_golden("step_a")
    b, h, w, seed = [int(v) for v in g["meta"]]
    state = onet.init_state(onet.FCDENSENET57, seed=seed, perturb=False)
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=0.02)
    try:
        onet.LIBRARY_OPS = ogeo.LIBRARY_OPS = True
        loss, dcl, sfl, grads, new_buf, ex = step.forward_backward(state, batch, onet.FCDENSENET57, 5.0, 20.0)
    finally:
        onet.LIBRARY_OPS = ogeo.LIBRARY_OPS = False
    assert abs(float(loss) - g["loss"][0]) / g["loss"][0] < 2e-3
    assert rel_err(ex["depth_1"], g["p1"]) < 1e-4
    assert int(new_buf["denseBlocksDown.0.layers.0.norm.num_batches_tracked"]) == 2


def _oracle_step_vs_fixture(tag, iters):
    g = load_golden(tag)
    b, h, w, seed, stride = [int(v) for v in g["meta"]]
    cfg = net.FCDENSENET57
    perturb = tag == "step_b"
    state = net.condition_state(net.init_state(cfg, seed=seed, perturb=perturb))
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=0.02 if tag == "step_b" else 0.005)
    mom = {}
    names = [k for k in state if not net.is_buffer(k)]
    for it in range(iters):
        loss, dcl, sfl, grads, new_buf, ex = step.forward_backward(state, batch, cfg, 5.0, 20.0)
        # WELL-CONDITIONED fixture: the stated 1e-4 bound on every loss scalar, both iterations (north_star)
        for name, got in (("loss", loss), ("dcl", dcl), ("sfl", sfl)):
            assert abs(float(got) - g[name][it]) / g[name][it] < 1e-4, (tag, it, name, float(got), g[name][it])
        if it == 0:
            sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
            assert rel_err(ex["depth_1"][sub], g["p1"]) < 1e-4
            assert rel_err(ex["depth_2"][sub], g["p2"]) < 1e-4
            assert rel_err(ex["scaled_1"][sub], g["s1"]) < 1e-4
            assert rel_err(ex["warped_2to1"][sub], g["w21"]) < 1e-4
            assert rel_err(ex["flow_1"][sub], g["f1"] * batch["boundaries"][sub].numpy()) < 1e-4
            assert int((ex["inter_1"][sub].numpy() != g["i1"]).sum()) == 0
            l2 = np.array([grads[k].double().norm().item() for k in names])
            assert np.all(np.abs(l2 - g["grad_l2"]) <= 2e-3 * g["grad_l2"] + 1e-5 * g["grad_l2"].max())
            for k in g:
                if k.startswith("grad::"):
                    # two fp32 CPU codes (this restatement / the reference's ATen BatchNorm): measured up to 2.7e-3 on a
                    # BatchNorm gamma gradient at bs8 256x320 (a cancelling sum over 82k samples), 1e-5 on the conv weights
                    assert rel_err(grads[k[6:]].detach(), g[k]) < 5e-3, k
        gn = step.clip_and_sgd(state, grads, mom, lr=1e-3)
        assert abs(float(gn) - g["gnorm"][it]) / g["gnorm"][it] < 1e-4, (it, float(gn), g["gnorm"][it])
        state.update(new_buf)
    if iters == 2:
        l2 = np.array([state[k].double().norm().item() for k in names])
        assert np.all(np.abs(l2 - g["w_l2_after"]) <= 1e-5 * g["w_l2_after"] + 1e-7)
        for k in g:
            if k.startswith("after::"):
                assert rel_err(state[k[7:]], g[k]) < 1e-4, k
            if k.startswith("buf::"):
                assert rel_err(state[k[5:]], g[k]) < 1e-5, k


def test_full_step_well_conditioned():
    """Two iterations of train.py:272-328 against the reference trace of the well-conditioned fixture
    (oracle.net.condition_state): 1e-4 on loss / dcl / sfl / gradient norm, 1e-4 on the updated weights."""
    _oracle_step_vs_fixture("step_b", 2)


def test_full_step_benchmark_configuration():
    """The same at the benchmarked configuration (bs8 256x320, seed 10085; BASELINE.json configs[1]); one
    iteration (an 8-image fp32 backward of FCDenseNet57 on the CPU takes ~15 s)."""
    _oracle_step_vs_fixture("step_c", 1)

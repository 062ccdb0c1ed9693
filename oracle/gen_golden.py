"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:   python -m oracle.gen_golden
Imports /root/reference/models.py and /root/reference/losses.py read-only with the two shims
of SURVEY.md section 8(c):
  * torch.Tensor.cuda -> identity          (the layers hard-code .cuda(); no GPU here)
  * torch.solve(B, A) -> (linalg.solve(A, B), None)   (models.py:392,493; removed in torch>=1.13)
and records, for seeded synthetic inputs, the reference's outputs and gradients.  The
fixtures are small (KBs): network parameters are regenerated from the seed by
`oracle.net.init_state`, inputs by `endo_b200.synthetic.make_batch`.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    sys.path.insert(0, REF)
    import models as ref_models  # noqa
    import losses as ref_losses  # noqa
    sys.path.pop(0)
    return ref_models, ref_losses


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def reference_function(name, path=REF + "/utils.py"):
    """Execute ONE function of a reference file that cannot be imported as a module here (utils.py needs plyfile /
    matplotlib): its FunctionDef is cut out of the unmodified source with `ast` and run against numpy."""
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    import random
    from pathlib import Path
    ns = {"np": np, "random": random, "Path": Path}
    try:
        import cv2
        ns["cv2"] = cv2
    except ImportError:
        pass
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def record_export():
    """utils.point_cloud_from_depth (utils.py:825-852) on a seeded 48 x 64 case, with and without the colour thresholds."""
    fn = reference_function("point_cloud_from_depth")
    rs = np.random.RandomState(77)
    h, w = 48, 64
    depth = (0.3 + 1.2 * rs.rand(h, w)).astype(np.float32)
    color = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    mask = (((yy - h / 2) ** 2 + (xx - w / 2) ** 2) < (0.45 * w) ** 2).astype(np.float32)
    k = np.array([[677.171 / 8, 0, w / 2.0], [0, 677.171 / 8, h / 2.0], [0, 0, 1]], dtype=np.float32)
    out = dict(depth=depth, color=color, mask=mask, k=k)
    out["pc_all"] = fn(depth, color, mask, k, 1)
    out["pc_ds2"] = fn(depth, color, mask, k, 2)
    out["pc_thr"] = fn(depth, color, mask, k, 1, min_threshold=60, max_threshold=200)
    np.savez_compressed(os.path.join(OUT, "export_a.npz"), **out)
    print("export_a", {k_: v.shape for k_, v in out.items()})


def record_raster():
    """utils.get_torch_training_data (utils.py:460-612) on the seeded scene above, with and without the inlier list."""
    fn = reference_function("get_torch_training_data")
    from oracle.export import raster_scene
    sc = raster_scene()
    out = dict(sc)
    args = ([sc["extr"][0], sc["extr"][1]], [sc["proj"][0], sc["proj"][1]], list(sc["pair_indexes"]), [list(p) for p in sc["pts"]],
            sc["mask"], sc["vis"])
    for tag, clean in (("clean", sc["clean"]), ("noclean", [])):
        dm, d, fm, fl = fn(*args, clean, list(sc["visible_view_indexes"]))
        out.update({f"{tag}_depth_mask": dm, f"{tag}_depth": d, f"{tag}_flow_mask": fm, f"{tag}_flow": fl})
        print("raster_a", tag, "points drawn per view", dm.sum(axis=(1, 2, 3)), "flow points", fm.sum(axis=(1, 2, 3)))
    np.savez_compressed(os.path.join(OUT, "raster_a.npz"), **out)


def record_sampler():
    """utils.generating_pos_and_increment (utils.py:410-438), the pair sampler in front of the rasteriser: seeded `random`, a sweep
    of sequence lengths / adjacent ranges / indexes incl. the short-sequence clamp (utils.py:418-419) and both edge branches."""
    import random
    fn = reference_function("generating_pos_and_increment")
    cases, out = [], []
    for n, rng in ((60, (5, 20)), (12, (5, 20)), (9, (5, 20)), (3, (1, 2)), (200, (1, 50)), (31, (10, 10))):
        views = list(range(100, 100 + 3 * n, 3))
        for idx in list(range(0, n, max(1, n // 12))) + [n - 1, n, 5 * n + 7]:
            for seed in (0, 1, 2):
                random.seed(1000 * seed + idx)
                pos, inc = fn(idx, views, list(rng))
                cases.append((n, rng[0], rng[1], idx, seed))
                out.append((pos, inc))
    np.savez_compressed(os.path.join(OUT, "sampler_a.npz"), cases=np.array(cases, dtype=np.int64), out=np.array(out, dtype=np.int64))
    print("sampler_a", len(cases), "cases")


def record_pipeline():
    """utils.get_pair_color_imgs (utils.py:441-457) on two synthetic JPEG frames written to a temporary sequence folder: the
    fixture keeps the JPEG bytes, the frames as cv2 decoded them here, and the reference's outputs for downsampling 4 / 3 / 2.5
    in both colour orders."""
    import tempfile
    import cv2
    fn = reference_function("get_pair_color_imgs")
    rs = np.random.RandomState(123)
    h, w = 216, 384
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for k, idx in enumerate((17, 23)):
            base = np.stack([128 + 100 * np.sin(xx / (23.0 + 5 * k) + c) * np.cos(yy / (31.0 - 3 * k) + 2 * c) for c in range(3)], -1)
            img = np.clip(base + rs.randn(h, w, 3) * 6.0, 0, 255).astype(np.uint8)
            cv2.imwrite(os.path.join(tmp, "{:08d}.jpg".format(idx)), img, [cv2.IMWRITE_JPEG_QUALITY, 90])
            out[f"jpeg_{k}"] = np.frombuffer(open(os.path.join(tmp, "{:08d}.jpg".format(idx)), "rb").read(), dtype=np.uint8)
            out[f"decoded_{k}"] = cv2.imread(os.path.join(tmp, "{:08d}.jpg".format(idx)))
        cases = []
        for ds, crop in ((4.0, (3, 51, 4, 92)), (3.0, (0, 72, 0, 128)), (2.5, (5, 85, 2, 150))):
            for mode in ("rgb", "bgr"):
                tag = f"ds{ds}_{mode}"
                out["out_" + tag] = fn(tmp, [17, 23], crop[0], crop[1], crop[2], crop[3], ds, False, mode)
                cases.append((ds, *crop, 1 if mode == "rgb" else 0))
        out["cases"] = np.array(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "pipeline_a.npz"), **out)
    print("pipeline_a", {k: v.shape for k, v in out.items()})


def main(only=None):
    for tag, rec in (("export_a", record_export), ("raster_a", record_raster), ("sampler_a", record_sampler),
                     ("pipeline_a", record_pipeline)):
        if only and tag in only:
            rec()
            only = [t for t in only if t != tag]
            if not only:
                return
    if only:                                    # regenerate selected step fixtures only: python -m oracle.gen_golden step_d
        ref_models, ref_losses = load_reference()
        torch.set_num_threads(8)
        table = {"step_a": (2, 64, 64, 404, False, False, 0.02, 1, 2), "step_b": (2, 64, 96, 505, True, True, 0.02, 1, 2),
                 "step_c": (8, 256, 320, 10085, False, True, 0.005, 4, 2), "step_d": (32, 256, 320, 20085, False, True, 0.005, 8, 1)}
        for tag in only:
            b, h, w, seed, perturb, cond, sp, stride, iters = table[tag]
            record_step(ref_models, ref_losses, tag, b, h, w, seed, perturb=perturb, conditioned=cond, sparse_prob=sp,
                        stride=stride, iters=iters, forward_only=(tag == "step_d"))
        return
    import endo_b200
    from oracle import net as onet
    ref_models, ref_losses = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    # ---------------------------------------------------------------- geometric layers + losses
    for tag, (b, h, w, seed, ones) in {"geo_a": (2, 64, 96, 101, False), "geo_b": (3, 32, 64, 202, True)}.items():
        batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, all_ones_boundary=ones, sparse_prob=0.02)
        d1, d2 = endo_b200.synthetic.jitter_depths(batch, seed=seed + 1)
        d1 = d1.clone().requires_grad_(True)
        d2 = d2.clone().requires_grad_(True)
        out = {}
        scale = ref_models.DepthScalingLayer(epsilon=1e-8)
        s1, std1 = scale([d1, batch["sparse_depths_1"], batch["sparse_depth_masks_1"]])
        g = torch.Generator().manual_seed(seed + 7)
        gs = torch.randn(s1.shape, generator=g)
        (gd_scale,) = torch.autograd.grad((s1 * gs).sum(), d1, retain_graph=True)
        out.update(scale_out=np32(s1), scale_std=np32(std1), scale_gout=np32(gs), scale_gd=np32(gd_scale))

        flow_layer = ref_models.FlowfromDepthLayer()
        f1 = flow_layer([d1, batch["boundaries"], batch["translations_1_wrt_2"], batch["rotations_1_wrt_2"],
                         batch["intrinsics"]])
        gf = torch.randn(f1.shape, generator=g)
        (gd_flow,) = torch.autograd.grad((f1 * gf * batch["boundaries"]).sum(), d1, retain_graph=True)
        out.update(flow_out=np32(f1), flow_gout=np32(gf * batch["boundaries"]), flow_gd=np32(gd_flow))

        warp = ref_models.DepthWarpingLayer(epsilon=1e-8)
        wd, inter = warp([d1, d2, batch["boundaries"], batch["translations_1_wrt_2"], batch["rotations_1_wrt_2"],
                          batch["intrinsics"]])
        gw = torch.randn(wd.shape, generator=g)
        gd1_w, gd2_w = torch.autograd.grad((wd * gw).sum(), [d1, d2], retain_graph=True)
        out.update(warp_out=np32(wd), warp_inter=np32(inter), warp_gout=np32(gw), warp_gd1=np32(gd1_w),
                   warp_gd2=np32(gd2_w))

        l1 = ref_losses.SparseMaskedL1Loss()
        fm = batch["sparse_flow_masks_1"] * batch["boundaries"]
        f1_leaf = f1.detach().clone().requires_grad_(True)
        lv = l1([batch["sparse_flows_1"] * batch["boundaries"], f1_leaf * batch["boundaries"], fm])
        (gf_l1,) = torch.autograd.grad(lv, f1_leaf, retain_graph=True)
        out.update(l1_out=np32(lv), l1_gflow=np32(gf_l1))

        ndl = ref_losses.NormalizedDistanceLoss(height=h, width=w)
        wd_leaf = wd.detach().clone().requires_grad_(True)      # partial derivatives of the loss layer alone
        nv = ndl([d1, wd_leaf, inter, batch["intrinsics"]])
        gd_n, gw_n = torch.autograd.grad(nv, [d1, wd_leaf], retain_graph=True)
        out.update(ndl_out=np32(nv), ndl_gd=np32(gd_n), ndl_gw=np32(gw_n))

        sil = ref_losses.ScaleInvariantLoss(epsilon=1e-8)
        sv = sil([d1, d2, batch["boundaries"]])
        gp_s, gg_s = torch.autograd.grad(sv, [d1, d2])
        out.update(sil_out=np32(sv), sil_gp=np32(gp_s), sil_gg=np32(gg_s))
        np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), meta=np.array([b, h, w, seed, int(ones)]), **out)
        print(tag, {k: v.shape for k, v in out.items()})

    # ---------------------------------------------------------------- network fwd / bwd
    b, h, w, seed = 2, 64, 96, 303
    cfg = onet.FCDENSENET57
    state = onet.init_state(cfg, seed=seed, perturb=True)
    model = ref_models.FCDenseNet57(n_classes=1)
    missing = model.load_state_dict(state, strict=True)
    model.train()
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed)
    x = (batch["boundaries"] * batch["colors_1"]).clone()
    y = model(x)
    g = torch.Generator().manual_seed(seed + 7)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out = dict(y=np32(y), gy=np32(gy))
    sd = model.state_dict()
    names = [k for k in onet.param_shapes(cfg) if not onet.is_buffer(k)]
    params = dict(model.named_parameters())
    out["grad_l2"] = np.array([params[k].grad.double().norm().item() for k in names], dtype=np.float64)
    out["grad_sum"] = np.array([params[k].grad.double().sum().item() for k in names], dtype=np.float64)
    for k in ("firstconv.weight", "firstconv.bias", "finalConv.weight", "denseBlocksDown.0.layers.3.conv.weight",
              "denseBlocksDown.0.layers.0.norm.weight", "denseBlocksDown.0.layers.0.norm.bias",
              "transDownBlocks.2.conv.weight", "transDownBlocks.2.norm.weight",
              "bottleneck.bottleneck.layers.1.conv.weight", "transUpBlocks.0.convTrans.1.weight",
              "transUpBlocks.4.convTrans.1.bias", "denseBlocksUp.4.layers.3.conv.weight",
              "denseBlocksUp.2.layers.0.norm.bias"):
        out["grad::" + k] = np32(params[k].grad)
    for k in ("denseBlocksDown.0.layers.0.norm.running_mean", "denseBlocksDown.0.layers.0.norm.running_var",
              "transDownBlocks.4.norm.running_var", "denseBlocksUp.4.layers.3.norm.running_mean",
              "bottleneck.bottleneck.layers.3.norm.running_var"):
        out["buf::" + k] = np32(sd[k])
    out["num_batches_tracked"] = np.array(int(sd["denseBlocksDown.0.layers.0.norm.num_batches_tracked"]))
    # eval-mode forward (evaluate.py path)
    model.eval()
    with torch.no_grad():
        out["y_eval"] = np32(model(x))
    np.savez_compressed(os.path.join(OUT, "net_a.npz"), meta=np.array([b, h, w, seed]), **out)
    print("net_a", y.shape, float(y.mean()))

    record_export()
    # ---------------------------------------------------------------- one full train step (train.py:272-328)
    record_step(ref_models, ref_losses, "step_a", 2, 64, 64, 404, perturb=False, conditioned=False)
    # well-conditioned variants (oracle.net.condition_state): the bounds that matter -- 1e-4 on every loss term
    record_step(ref_models, ref_losses, "step_b", 2, 64, 96, 505, perturb=True, conditioned=True)
    # ... and at the benchmarked configuration (BASELINE.json configs[1]: bs8 256x320, seed of train.py:80)
    record_step(ref_models, ref_losses, "step_c", 8, 256, 320, 10085, perturb=False, conditioned=True, sparse_prob=0.005,
                stride=4)
    # BASELINE.json configs[2]: bs32 256x320 (one iteration, scalars + a coarse depth map: the loss-parity case of the bf16 path)
    record_step(ref_models, ref_losses, "step_d", 32, 256, 320, 20085, perturb=False, conditioned=True, sparse_prob=0.005,
                stride=8, iters=1, forward_only=True)


STEP_TENSORS = ("firstconv.weight", "finalConv.weight", "denseBlocksUp.4.layers.3.conv.weight",
                "denseBlocksDown.2.layers.1.norm.weight")


def record_step(ref_models, ref_losses, tag, b, h, w, seed, perturb, conditioned, sparse_prob=0.02, stride=1, iters=2,
                forward_only=False):
    """Two iterations of train.py:272-328 on the unmodified reference modules; records the loss terms, the gradient norm,
    per-tensor gradient / weight norms and a few tensors.  `stride` subsamples the stored maps (large configurations).
    `forward_only`: train.py:272-315 under no_grad (BatchNorm in training mode), loss terms and maps only -- the bs32
    case: autograd would keep ~46 GB of activations alive for the two 32-image forwards, more than this container has."""
    if forward_only:
        torch.set_grad_enabled(False)
    import endo_b200
    from oracle import net as onet
    cfg = onet.FCDENSENET57
    names = [k for k in onet.param_shapes(cfg) if not onet.is_buffer(k)]
    state = onet.init_state(cfg, seed=seed, perturb=perturb)          # perturb=False: reference init (kaiming / zero bias / gamma 1)
    if conditioned:
        state = onet.condition_state(state)
    model = ref_models.FCDenseNet57(n_classes=1)
    model.load_state_dict(state, strict=True)
    model.train()
    batch = endo_b200.synthetic.make_batch(b, h, w, seed=seed, sparse_prob=sparse_prob)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    scale = ref_models.DepthScalingLayer(epsilon=1e-8)
    warp = ref_models.DepthWarpingLayer(epsilon=1e-8)
    flow_layer = ref_models.FlowfromDepthLayer()
    l1 = ref_losses.SparseMaskedL1Loss()
    ndl = ref_losses.NormalizedDistanceLoss(height=h, width=w)
    rec = {"loss": [], "dcl": [], "sfl": [], "gnorm": []}
    first = {}
    for it in range(iters):
        B = batch
        c1 = B["boundaries"] * B["colors_1"]
        c2 = B["boundaries"] * B["colors_2"]
        p1 = model(c1)
        p2 = model(c2)
        s1, _ = scale([p1, B["sparse_depths_1"], B["sparse_depth_masks_1"]])
        s2, _ = scale([p2, B["sparse_depths_2"], B["sparse_depth_masks_2"]])
        f1 = flow_layer([s1, B["boundaries"], B["translations_1_wrt_2"], B["rotations_1_wrt_2"], B["intrinsics"]])
        f2 = flow_layer([s2, B["boundaries"], B["translations_2_wrt_1"], B["rotations_2_wrt_1"], B["intrinsics"]])
        m1 = B["sparse_flow_masks_1"] * B["boundaries"]
        m2 = B["sparse_flow_masks_2"] * B["boundaries"]
        sfl = 20.0 * 0.5 * (l1([B["sparse_flows_1"] * B["boundaries"], f1 * B["boundaries"], m1]) +
                            l1([B["sparse_flows_2"] * B["boundaries"], f2 * B["boundaries"], m2]))
        w21, i1 = warp([s1, s2, B["boundaries"], B["translations_1_wrt_2"], B["rotations_1_wrt_2"], B["intrinsics"]])
        w12, i2 = warp([s2, s1, B["boundaries"], B["translations_2_wrt_1"], B["rotations_2_wrt_1"], B["intrinsics"]])
        dcl = 5.0 * 0.5 * (ndl([s1, w21, i1, B["intrinsics"]]) + ndl([s2, w12, i2, B["intrinsics"]]))
        loss = dcl + sfl
        if forward_only:
            rec["loss"].append(loss.item()); rec["dcl"].append(dcl.item()); rec["sfl"].append(sfl.item())
            sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
            first.update(p1=np32(p1[sub]), p2=np32(p2[sub]), s1=np32(s1[sub]), w21=np32(w21[sub]), i1=np32(i1[sub]))
            break
        opt.zero_grad()
        loss.backward()
        if it == 0 and conditioned:
            params = dict(model.named_parameters())
            first["grad_l2"] = np.array([params[k].grad.double().norm().item() for k in names], dtype=np.float64)
            for k in STEP_TENSORS:
                first["grad::" + k] = np32(params[k].grad)
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        rec["loss"].append(loss.item()); rec["dcl"].append(dcl.item()); rec["sfl"].append(sfl.item())
        rec["gnorm"].append(float(gn))
        if it == 0:
            sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
            first.update(p1=np32(p1[sub]), s1=np32(s1[sub]), w21=np32(w21[sub]), i1=np32(i1[sub]), f1=np32(f1[sub]))
            if conditioned:
                first.update(p2=np32(p2[sub]), inter_sum=np.array([float(i1.sum()), float(i2.sum())]))
    out = {k: np.array(v, dtype=np.float64) for k, v in rec.items()}
    out.update(first)
    if forward_only:
        torch.set_grad_enabled(True)
        np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), meta=np.array([b, h, w, seed, stride]), **out)
        print(tag, rec)
        return
    params = dict(model.named_parameters())
    out["w_l2_after"] = np.array([params[k].double().norm().item() for k in names], dtype=np.float64)
    for k in STEP_TENSORS:
        out["after::" + k] = np32(params[k])
    if conditioned:
        sd = model.state_dict()
        for k in ("denseBlocksDown.0.layers.0.norm.running_mean", "denseBlocksUp.4.layers.3.norm.running_var",
                  "bottleneck.bottleneck.layers.3.norm.running_var"):
            out["buf::" + k] = np32(sd[k])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), meta=np.array([b, h, w, seed] + ([stride] if conditioned else [])), **out)
    print(tag, rec)


if __name__ == "__main__":
    main(sys.argv[1:])

"""GPU debugging aid: stage-by-stage comparison of the CUDA network against the oracle.
   python tools/debug_net.py [B H W]      (prints one line per stage / per parameter gradient)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import endo_b200  # noqa: E402
from oracle import net as onet  # noqa: E402


def plan57(cfg, B, H, W, G):
    nd, g = len(cfg.down_blocks), cfg.growth_rate
    C0, Dn, U, Un = [0] * (nd + 1), [0] * (nd + 1), [0] * (nd + 1), [0] * (nd + 1)
    cur = cfg.out_chans_first_conv
    for l in range(nd):
        C0[l], Dn[l] = cur, g * cfg.down_blocks[l]
        cur += Dn[l]
    C0[nd], Dn[nd] = cur, g * cfg.bottleneck_layers
    prev = g * cfg.bottleneck_layers
    for i in range(nd):
        l = nd - 1 - i
        U[l], Un[l] = prev, g * cfg.up_blocks[i]
        prev = Un[l]
    Ctot = [U[l] + C0[l] + Dn[l] + Un[l] for l in range(nd + 1)]
    off, xoff = 0, []
    for l in range(nd + 1):
        xoff.append(off)
        off = (off + 4 * B * (H >> l) * (W >> l) * Ctot[l] + 255) // 256 * 256
    return dict(C0=C0, Dn=Dn, U=U, Un=Un, Ctot=Ctot, xoff=xoff)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 64, 96)
    math_mode = sys.argv[4] if len(sys.argv) >= 5 else "fp32"
    cfg = onet.FCDENSENET57
    state = onet.init_state(cfg, seed=303, perturb=True)
    batch = endo_b200.synthetic.make_batch(B, H, W, seed=303)
    x = batch["boundaries"] * batch["colors_1"]
    rec = {}
    p64 = {k: (v if v.dtype == torch.long else v.double()) for k, v in state.items()}
    p64 = {k: (v if onet.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in p64.items()}
    y_ref = onet.forward(p64, x.double(), cfg, True, {}, rec)
    gy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(5))
    (y_ref * gy.double()).sum().backward()

    model = endo_b200.models.FCDenseNet57(1, math=math_mode)
    model.load_state_dict(state)
    model.cuda().train()
    model._debug_keep_acts = True
    y = model(x.cuda())
    torch.cuda.synchronize()
    acts = model._debug_acts
    P = plan57(cfg, B, H, W, 1)
    nd = len(cfg.down_blocks)

    def level(l):
        n = B * (H >> l) * (W >> l) * P["Ctot"][l]
        buf = acts[P["xoff"][l]:P["xoff"][l] + 4 * n].view(torch.float32).reshape(B, H >> l, W >> l, P["Ctot"][l])
        return buf.permute(0, 3, 1, 2)

    def show(name, got, ref):
        print(f"{name:28s} rel_err {rel(got, ref):.3e}   |ref|max {float(ref.abs().max()):.3e}")

    show("firstconv", level(0)[:, P["U"][0]:P["U"][0] + P["C0"][0]], rec["firstconv"])
    for l in range(nd):
        cs = P["C0"][l] + P["Dn"][l]
        show(f"down{l} (skip)", level(l)[:, P["U"][l]:P["U"][l] + cs], rec[f"down{l}"])
        show(f"td{l}", level(l + 1)[:, P["U"][l + 1]:P["U"][l + 1] + P["C0"][l + 1]], rec[f"td{l}"])
    show("bottleneck", level(nd)[:, P["C0"][nd]:P["C0"][nd] + P["Dn"][nd]], rec["bottleneck"])
    for i in range(nd):
        l = nd - 1 - i
        cs = P["U"][l] + P["C0"][l] + P["Dn"][l]
        show(f"tu{i} (cat up,skip)", level(l)[:, :cs], rec[f"tu{i}"])
        if i != nd - 1:
            show(f"up{i} (new)", level(l)[:, cs:cs + P["Un"][l]], rec[f"up{i}"])
        else:
            show(f"up{i} (all)", level(l), rec[f"up{i}"])
    show("y", y, y_ref)

    (y * gy.cuda()).sum().backward()
    torch.cuda.synchronize()
    names = [k for k in state if not onet.is_buffer(k)]
    params = dict(model.named_parameters())
    gmax = max(float(p64[k].grad.abs().max()) for k in names)
    rows = []
    for k in names:
        g_ref = p64[k].grad
        g = params[k].grad
        err = float((g.double().cpu() - g_ref).abs().max())
        rows.append((err / max(float(g_ref.abs().max()), 1e-6 * gmax), k, float(g_ref.abs().max()), err))
    rows.sort(reverse=True)
    print("---- worst parameter gradients (err relative to the tensor's own max, floor 1e-6 of global max)")
    for r in rows[:25]:
        print(f"{r[1]:52s} rel {r[0]:.3e}  |ref|max {r[2]:.3e}  abs_err {r[3]:.3e}")
    print("median rel", float(np.median([r[0] for r in rows])))


if __name__ == "__main__":
    main()

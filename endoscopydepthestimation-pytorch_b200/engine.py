"""FC-DenseNet on the B200 engine: the reference's module tree as a parameter container, one
C-ABI call for the whole forward and one for the whole backward.

Mirrors /root/reference/models.py:19-208 (DenseLayer, DenseBlock, TransitionDown, TransitionUp,
Bottleneck, FCDenseNet, FCDenseNet57/67/103): identical sub-module names, parameter shapes and
registration order, so reference checkpoints (`utils.py:674-682`, keys optionally prefixed with
`module.` by DataParallel), `utils.init_net` (`utils.py:619-671`) and
`torch.optim.SGD(model.parameters())` (`train.py:202`) work unchanged.  What differs is execution:
`FCDenseNet.forward` does not run the sub-modules; it hands ONE flat parameter array to
`endo_net_fwd` (csrc/net.cu), which runs the NHWC / in-place-concatenation / fused BN-ReLU-conv
kernels, and autograd's backward is `endo_net_bwd`, which accumulates every parameter gradient
directly into ONE flat gradient array that `p.grad` of each nn.Parameter is a view of (so the
DDP all-reduce and the fused clip+SGD step see a single contiguous bucket).
"""
import ctypes
from collections import OrderedDict

import torch
from torch import nn

from . import _lib as L

MATH_MODES = {"fp32": 0, "tf32": 1, "bf16": 2, "tf32x3": 3, "bf16x3": 4}


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's names (never executed one by one)
# ------------------------------------------------------------------------------------------------
class DenseLayer(nn.Sequential):
    """models.py:19-28: norm (BatchNorm2d) -> relu -> conv (3x3, growth_rate outputs)."""

    def __init__(self, in_channels, growth_rate):
        super().__init__()
        self.add_module("norm", nn.BatchNorm2d(in_channels))
        self.add_module("relu", nn.ReLU(True))
        self.add_module("conv", nn.Conv2d(in_channels, growth_rate, kernel_size=3, stride=1, padding=1, bias=True))


class DenseBlock(nn.Module):
    """models.py:31-53."""

    def __init__(self, in_channels, growth_rate, n_layers, upsample=False):
        super().__init__()
        self.upsample = upsample
        self.layers = nn.ModuleList([DenseLayer(in_channels + i * growth_rate, growth_rate) for i in range(n_layers)])


class TransitionDown(nn.Sequential):
    """models.py:56-67: norm -> relu -> conv 1x1 -> maxpool 2."""

    def __init__(self, in_channels):
        super().__init__()
        self.add_module("norm", nn.BatchNorm2d(num_features=in_channels))
        self.add_module("relu", nn.ReLU(inplace=True))
        self.add_module("conv", nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0, bias=True))
        self.add_module("maxpool", nn.MaxPool2d(2))


class TransitionUp(nn.Module):
    """models.py:70-80: convTrans = Sequential(Upsample nearest x2, Conv2d 3x3)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.convTrans = nn.Sequential(nn.Upsample(mode="nearest", scale_factor=2),
                                       nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1))


class Bottleneck(nn.Sequential):
    """models.py:83-90."""

    def __init__(self, in_channels, growth_rate, n_layers):
        super().__init__()
        self.add_module("bottleneck", DenseBlock(in_channels, growth_rate, n_layers, upsample=True))


def _raise_not_executable(*_a, **_k):
    raise RuntimeError("endo_b200 sub-modules are parameter containers; call the FCDenseNet itself "
                       "(the whole network runs as one fused CUDA pipeline, there is no per-module or CPU path)")


for _cls in (DenseLayer, DenseBlock, TransitionDown, TransitionUp, Bottleneck):
    _cls.forward = _raise_not_executable


class _NetFn(torch.autograd.Function):
    """forward = endo_net_fwd, backward = endo_net_bwd.  `anchor` is a dummy leaf that keeps the node in
    the graph; parameter gradients are written straight into the module's flat gradient bucket."""

    @staticmethod
    @L.on_device
    def forward(ctx, x, anchor, net, groups):
        lib = L.lib()
        b, _, h, w = x.shape
        cfg = net._cfg
        acts_bytes = lib.endo_net_activation_bytes(ctypes.byref(cfg), b, h, w)
        if acts_bytes == 0:
            raise RuntimeError(f"FCDenseNet: unsupported input shape {tuple(x.shape)} "
                               f"(H and W must be multiples of {1 << cfg.n_down}, batch divisible by the group count)")
        acts = torch.empty(acts_bytes, dtype=torch.uint8, device=x.device)
        y = torch.empty((b, 1, h, w), dtype=torch.float32, device=x.device)
        training = 1 if net.training else 0
        L.check(lib.endo_net_fwd(ctypes.byref(cfg), x.data_ptr(), net._flat.data_ptr(), net._flat_buf.data_ptr(),
                                 y.data_ptr(), acts.data_ptr(), acts.numel(), b, h, w, groups, training,
                                 net._math, L.stream_ptr(x.device)), "net_fwd")
        ctx.net, ctx.acts, ctx.groups, ctx.training = net, acts, groups, training
        ctx.param_version = net._flat._version
        if getattr(net, "_debug_keep_acts", False):
            net._debug_acts = acts
        ctx.save_for_backward(x)
        return y

    @staticmethod
    @L.on_device
    def backward(ctx, g_y):
        net = ctx.net
        if not ctx.training:
            raise RuntimeError("FCDenseNet backward is only defined in training mode (BatchNorm batch statistics)")
        # deviations from the nn.Module autograd contract are errors, not silent differences (the parameter gradients are
        # written straight into the flat bucket, outside autograd's own accumulation)
        if ctx.acts is None:
            raise RuntimeError("FCDenseNet: second backward through the same forward (retain_graph) is not supported: "
                               "the saved activations are released after the first backward")
        if net._flat._version != ctx.param_version:
            raise RuntimeError("FCDenseNet: parameters were modified in place between forward and backward "
                               "(the backward would differentiate a different function than the forward computed)")
        (x,) = ctx.saved_tensors
        lib = L.lib()
        b, _, h, w = x.shape
        cfg = net._cfg
        g_y = L.contig(g_y)
        accumulate = net._prepare_grad_views()
        scratch = torch.empty(lib.endo_net_backward_scratch_bytes(ctypes.byref(cfg), b, h, w), dtype=torch.uint8,
                              device=x.device)
        L.check(lib.endo_net_bwd(ctypes.byref(cfg), g_y.data_ptr(), x.data_ptr(), net._flat.data_ptr(),
                                 net._flat_grad.data_ptr(), None, ctx.acts.data_ptr(), ctx.acts.numel(),
                                 scratch.data_ptr(), scratch.numel(), b, h, w, ctx.groups, accumulate, net._math,
                                 L.stream_ptr(x.device)), "net_bwd")
        net._finish_grad_views()
        ctx.acts = None
        return None, None, None, None


class FCDenseNet(nn.Module):
    """models.py:100-187.  Same constructor; `forward(x)` returns abs(finalConv(...)) as [B,1,H,W]."""

    def __init__(self, in_channels=3, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5,
                 growth_rate=16, out_chans_first_conv=48, n_classes=1, math="fp32"):
        super().__init__()
        if len(down_blocks) != len(up_blocks) or len(down_blocks) > 8:
            raise ValueError("down_blocks and up_blocks must have the same length (<= 8)")
        self.down_blocks, self.up_blocks = tuple(down_blocks), tuple(up_blocks)
        skip_counts = []
        self.add_module("firstconv", nn.Conv2d(in_channels, out_chans_first_conv, kernel_size=3, stride=1, padding=1,
                                               bias=True))
        cur = out_chans_first_conv
        self.denseBlocksDown = nn.ModuleList([])
        self.transDownBlocks = nn.ModuleList([])
        for n in down_blocks:
            self.denseBlocksDown.append(DenseBlock(cur, growth_rate, n))
            cur += growth_rate * n
            skip_counts.insert(0, cur)
            self.transDownBlocks.append(TransitionDown(cur))
        self.add_module("bottleneck", Bottleneck(cur, growth_rate, bottleneck_layers))
        prev = growth_rate * bottleneck_layers
        self.transUpBlocks = nn.ModuleList([])
        self.denseBlocksUp = nn.ModuleList([])
        for i, n in enumerate(up_blocks):
            self.transUpBlocks.append(TransitionUp(prev, prev))
            cur = prev + skip_counts[i]
            self.denseBlocksUp.append(DenseBlock(cur, growth_rate, n, upsample=(i != len(up_blocks) - 1)))
            prev = growth_rate * n
            cur += prev
        self.finalConv = nn.Conv2d(cur, n_classes, kernel_size=1, stride=1, padding=0, bias=True)

        cfg = L.NetConfig()
        cfg.in_channels, cfg.n_down = in_channels, len(down_blocks)
        for i, n in enumerate(down_blocks):
            cfg.down_layers[i] = n
        for i, n in enumerate(up_blocks):
            cfg.up_layers[i] = n
        cfg.bottleneck_layers, cfg.growth_rate = bottleneck_layers, growth_rate
        cfg.first_conv_channels, cfg.n_classes = out_chans_first_conv, n_classes
        self._cfg = cfg
        self._math = MATH_MODES[math]
        self._flat = self._flat_grad = self._flat_buf = self._anchor = None
        self._flat_ok = False
        self._params = self._bufs = self._nbt = None

    # ---------------------------------------------------------------- flat storage
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)      # .cuda()/.to()/.float() re-allocate every tensor
        self._flat_ok = False
        return out

    def _flatten(self, device):
        """(Re)pack parameters and BN running buffers into flat arrays in state_dict order and make every
        nn.Parameter / buffer a view of them.  Idempotent; triggered after .cuda()/.to()."""
        params = [p for _, p in self.named_parameters()]
        bufs, nbt = [], []
        for name, b_ in self.named_buffers():
            (nbt if name.endswith("num_batches_tracked") else bufs).append(b_)
        lib = L.lib()
        n = sum(p.numel() for p in params)
        if n != lib.endo_net_param_count(ctypes.byref(self._cfg)):
            raise RuntimeError("parameter layout mismatch between the module tree and libendo_b200 "
                               "(unsupported FCDenseNet configuration)")
        nb = sum(b_.numel() for b_ in bufs)
        assert nb == lib.endo_net_buffer_count(ctypes.byref(self._cfg))
        flat = torch.empty(n, dtype=torch.float32, device=device)
        # 4 spare floats behind the bucket: slot n carries the is-finite flag through the gradient all-reduce
        self._flat_grad_store = torch.zeros(n + 4, dtype=torch.float32, device=device)
        flat_grad = self._flat_grad_store[:n]
        flat_buf = torch.empty(max(nb, 1), dtype=torch.float32, device=device)
        off = 0
        self._grad_views = []
        with torch.no_grad():
            for p in params:
                k = p.numel()
                flat[off:off + k].copy_(p.detach().reshape(-1).to(device=device, dtype=torch.float32))
                old_grad = p.grad
                p.data = flat[off:off + k].view(p.shape)
                gv = flat_grad[off:off + k].view(p.shape)
                if old_grad is not None:
                    gv.copy_(old_grad.to(device))
                    p.grad = gv
                self._grad_views.append(gv)
                off += k
            off = 0
            for b_ in bufs:
                k = b_.numel()
                flat_buf[off:off + k].copy_(b_.detach().reshape(-1).to(device=device, dtype=torch.float32))
                b_.data = flat_buf[off:off + k].view(b_.shape)
                off += k
            for t in nbt:
                t.data = t.data.to(device)
        self._flat, self._flat_grad, self._flat_buf = flat, flat_grad, flat_buf
        self._params, self._bufs, self._nbt = params, bufs, nbt
        self._anchor = torch.zeros((), dtype=torch.float32, device=device, requires_grad=True)
        self._flat_ok = True

    def materialize(self, device=None):
        """Build the flat parameter / gradient / buffer arrays now (they are otherwise built by the first CUDA forward);
        needed before anything touches `flat_params` -- e.g. `ddp.broadcast_parameters` -- ahead of the first step."""
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("endo_b200.FCDenseNet lives on CUDA devices only (move it with .cuda() first)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._ensure_flat(device)
        return self

    def _ensure_flat(self, device):
        ok = self._flat_ok and self._flat.device == device
        if ok:   # cheap sanity check that nobody re-allocated the parameters behind our back
            last = self._params[-1]
            ok = (self._params[0].data_ptr() == self._flat.data_ptr()
                  and last.data_ptr() == self._flat.data_ptr() + 4 * (self._flat.numel() - last.numel()))
        if not ok:
            self._flatten(device)

    @property
    def flat_params(self):
        """The flat fp32 parameter array (state_dict order); every nn.Parameter is a view of it."""
        return self._flat

    @property
    def flat_grads(self):
        """The flat gradient bucket: `p.grad` of every parameter is a view of it after backward."""
        return self._flat_grad

    def _prepare_grad_views(self) -> int:
        """Decide whether this backward accumulates into the bucket (returns 1) or starts it (0)."""
        grads = [p.grad for p in self._params]
        if all(g is None for g in grads):
            self._foreign = None
            return 0                                     # kernel clears the bucket itself
        if all(g is not None and g.data_ptr() == v.data_ptr() for g, v in zip(grads, self._grad_views)):
            self._foreign = None
            return 1
        # foreign .grad tensors (user-assigned): fold them in afterwards
        self._foreign = grads
        return 0

    def _finish_grad_views(self):
        foreign = getattr(self, "_foreign", None)
        for p, v in zip(self._params, self._grad_views):
            p.grad = v
        if foreign is not None:
            with torch.no_grad():
                for v, g in zip(self._grad_views, foreign):
                    if g is not None and g.data_ptr() != v.data_ptr():
                        v.add_(g.to(v.device))
            self._foreign = None

    # ---------------------------------------------------------------- forward
    def _run(self, x, groups):
        if not x.is_cuda:
            raise RuntimeError("endo_b200.FCDenseNet runs on CUDA tensors only (there is no CPU fallback)")
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != self._cfg.in_channels:
            raise RuntimeError(f"expected a float32 [B,{self._cfg.in_channels},H,W] input, got {tuple(x.shape)} {x.dtype}")
        self._ensure_flat(x.device)
        if self.training and torch.is_grad_enabled() and not all(p.requires_grad for p in self._params):
            raise RuntimeError("endo_b200.FCDenseNet computes the gradient of EVERY parameter in one fused backward; "
                               "frozen parameters (requires_grad=False) are not supported -- run under torch.no_grad() "
                               "or leave them out of the optimiser instead")
        x = L.contig(x)
        if torch.is_grad_enabled() and self.training:
            y = _NetFn.apply(x, self._anchor, self, groups)
        else:
            with torch.no_grad():
                y = _NetFn.apply(x, self._anchor, self, groups)
        if self.training and self._nbt:
            with torch.no_grad():
                torch._foreach_add_(self._nbt, groups)   # num_batches_tracked += 1 per BN-train forward
        return y

    def forward(self, x):
        return self._run(x, 1)

    def forward_pair(self, x1, x2):
        """`net(x1), net(x2)` of train.py:276-277 as ONE launch sequence: the two batches are stacked and
        run with two independent BatchNorm statistic groups (running buffers receive both updates, in order)."""
        b = x1.shape[0]
        y = self._run(torch.cat([x1, x2], dim=0), 2)
        return y[:b], y[b:]


def FCDenseNet57(n_classes, **kw):
    """models.py:190-194."""
    return FCDenseNet(in_channels=3, down_blocks=(4, 4, 4, 4, 4), up_blocks=(4, 4, 4, 4, 4), bottleneck_layers=4,
                      growth_rate=12, out_chans_first_conv=48, n_classes=n_classes, **kw)


def FCDenseNet67(n_classes, **kw):
    """models.py:197-201."""
    return FCDenseNet(in_channels=3, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5,
                      growth_rate=16, out_chans_first_conv=48, n_classes=n_classes, **kw)


def FCDenseNet103(n_classes, **kw):
    """models.py:204-208."""
    return FCDenseNet(in_channels=3, down_blocks=(4, 5, 7, 10, 12), up_blocks=(12, 10, 7, 5, 4),
                      bottleneck_layers=15, growth_rate=16, out_chans_first_conv=48, n_classes=n_classes, **kw)


def kaiming_init_(model: nn.Module, seed=None):
    """Restatement of `utils.init_net(..., type="kaiming", mode="fan_in", activation_mode="relu",
    distribution="normal")` (utils.py:619-671): Kaiming-normal conv weights, zero biases, BN gamma 1."""
    if seed is not None:
        torch.manual_seed(seed)
    for module in model.modules():
        if hasattr(module, "weight") and module.weight is not None:
            if "BatchNorm" not in module.__class__.__name__:
                torch.nn.init.kaiming_normal_(module.weight, mode="fan_in", nonlinearity="relu")
            else:
                torch.nn.init.constant_(module.weight, 1)
        if hasattr(module, "bias") and module.bias is not None:
            torch.nn.init.constant_(module.bias, 0)
    return model

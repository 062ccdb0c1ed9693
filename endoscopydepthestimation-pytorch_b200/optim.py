"""Optimiser tail of the reference step (`/root/reference/train.py:202, 327-328`) as ONE fused pass over
the flat parameter / gradient bucket: `clip_grad_norm_(params, 10.0)` + `SGD(momentum=0.9).step()`."""
import torch

from . import _lib as L


class FusedClipSGD:
    """clip_grad_norm_ + SGD(momentum, no weight decay, no nesterov) on `net.flat_params` / `net.flat_grads`.

    Semantics follow torch (first step initialises the momentum buffer with the clipped gradient;
    gradients are scaled in place by min(1, max_norm / (norm + 1e-6)))."""

    def __init__(self, net, lr=1.0e-3, momentum=0.9, max_norm=10.0):
        self.net, self.lr, self.momentum, self.max_norm = net, float(lr), float(momentum), float(max_norm)
        self.buf = None
        self.grad_norm = None
        self.steps = 0

    def step_graph(self, lr_dev, finite_flag=None):
        """The same update with the learning rate read from the device scalar `lr_dev` and no step-dependent launch
        parameter (capturable in a CUDA graph; the momentum buffer starts at zero, which reproduces torch's first step)."""
        net = self.net
        flat, grad = net.flat_params, net.flat_grads
        if flat is None:
            raise RuntimeError("FusedClipSGD.step_graph() before the first forward/backward of the network")
        L.require_cuda(flat, grad, lr_dev)
        if self.buf is None or self.buf.numel() != flat.numel() or self.buf.device != flat.device:
            self.buf = torch.zeros_like(flat)
            self.grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.device)
            self.steps = 0
        lib = L.lib()
        n = flat.numel()
        ws = L.workspace(flat.device, lib.endo_sgd_workspace_bytes(n))
        with torch.cuda.device(flat.device):
            L.check(lib.endo_sgd_clip_step_dev(flat.data_ptr(), grad.data_ptr(), self.buf.data_ptr(), n, lr_dev.data_ptr(),
                                               self.momentum, self.max_norm, L.ptr(finite_flag), self.grad_norm.data_ptr(),
                                               ws.data_ptr(), ws.numel(), L.stream_ptr(flat.device)), "sgd_clip_step_dev")
        self.steps += 1
        return self.grad_norm

    def step(self, finite_flag=None, lr=None):
        net = self.net
        flat, grad = net.flat_params, net.flat_grads
        if flat is None:
            raise RuntimeError("FusedClipSGD.step() before the first forward/backward of the network")
        L.require_cuda(flat, grad)
        if self.buf is None or self.buf.data_ptr() == 0 or self.buf.numel() != flat.numel() or self.buf.device != flat.device:
            self.buf = torch.zeros_like(flat)
            self.grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.device)
            self.steps = 0
        lib = L.lib()
        n = flat.numel()
        ws = L.workspace(flat.device, lib.endo_sgd_workspace_bytes(n))
        L.check(lib.endo_sgd_clip_step(flat.data_ptr(), grad.data_ptr(), self.buf.data_ptr(), n,
                                       float(self.lr if lr is None else lr), self.momentum, self.max_norm,
                                       1 if self.steps == 0 else 0, L.ptr(finite_flag), self.grad_norm.data_ptr(),
                                       ws.data_ptr(), ws.numel(), L.stream_ptr(flat.device)), "sgd_clip_step")
        self.steps += 1
        return self.grad_norm

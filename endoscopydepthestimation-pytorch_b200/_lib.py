"""ctypes binding of libendo_b200.so (the C ABI in include/endo_b200.h).

The library is the product: there is no fallback.  `lib()` raises if the shared object is
missing (run `python -m endo_b200.build` / `__graft_entry__.build()`), and every wrapper
raises on a non-CUDA tensor or a non-zero status code.
"""
import ctypes
import os
import threading
from ctypes import c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libendo_b200.so")
WS_HEADER_BYTES = 256

_lib = None
_lock = threading.Lock()


class NetConfig(ctypes.Structure):
    """endo_net_config (include/endo_b200.h) == FCDenseNet.__init__ arguments (models.py:101-103)."""
    _fields_ = [("in_channels", c_int), ("n_down", c_int), ("down_layers", c_int * 8), ("up_layers", c_int * 8),
                ("bottleneck_layers", c_int), ("growth_rate", c_int), ("first_conv_channels", c_int),
                ("n_classes", c_int)]


_P = c_void_p
_SIGNATURES = {
    "endo_version": (c_int, []),
    "endo_strerror": (ctypes.c_char_p, [c_int]),
    "endo_launch_count": (c_ulonglong, []),
    "endo_prof_enable": (None, [c_int]),
    "endo_prof_categories": (c_int, []),
    "endo_prof_category_name": (ctypes.c_char_p, [c_int]),
    "endo_prof_collect": (c_int, [_P, _P]),
    "endo_depth_scale_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "endo_depth_scale_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, _P]),
    "endo_depth_scale_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, _P]),
    "endo_flow_from_depth_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "endo_flow_from_depth_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "endo_depth_warp_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "endo_depth_warp_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "endo_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "endo_sparse_l1_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, _P]),
    "endo_sparse_l1_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "endo_norm_dist_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, _P]),
    "endo_norm_dist_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "endo_scale_inv_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, _P]),
    "endo_scale_inv_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "endo_net_param_count": (c_longlong, [ctypes.POINTER(NetConfig)]),
    "endo_net_buffer_count": (c_longlong, [ctypes.POINTER(NetConfig)]),
    "endo_net_activation_bytes": (c_size_t, [ctypes.POINTER(NetConfig), c_int, c_int, c_int]),
    "endo_net_backward_scratch_bytes": (c_size_t, [ctypes.POINTER(NetConfig), c_int, c_int, c_int]),
    "endo_net_fwd": (c_int, [ctypes.POINTER(NetConfig), _P, _P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int,
                             c_int, c_int, _P]),
    "endo_net_bwd": (c_int, [ctypes.POINTER(NetConfig), _P, _P, _P, _P, _P, _P, c_size_t, _P, c_size_t, c_int,
                             c_int, c_int, c_int, c_int, c_int, _P]),
    "endo_debug_trace_read": (c_int, [_P, c_int]),
    "endo_tc_probe": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "endo_point_cloud_workspace_bytes": (c_size_t, [c_int, c_int]),
    "endo_point_cloud_from_depth": (c_int, [_P, _P, _P, c_float, c_float, c_float, c_float, c_int, c_int, c_int, c_int, c_float,
                                            c_float, _P, _P, _P, c_size_t, _P]),
    "endo_rasterize_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "endo_rasterize_pair": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "endo_resize_crop_u8": (c_int, [_P, c_int, c_int, c_double, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "endo_tma_probe": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "endo_sgd_workspace_bytes": (c_size_t, [c_longlong]),
    "endo_sgd_clip_step": (c_int, [_P, _P, _P, c_longlong, c_float, c_float, c_float, c_int, _P, _P, _P, c_size_t,
                                   _P]),
    "endo_sgd_clip_step_dev": (c_int, [_P, _P, _P, c_longlong, _P, c_float, c_float, _P, _P, _P, c_size_t, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded shared library (loads it on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                        "endo_b200 has no CPU or PyTorch fallback.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(code: int, what: str = ""):
    if code != 0:
        raise RuntimeError(f"libendo_b200 {what} failed: {lib().endo_strerror(code).decode()} (code {code})")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("endo_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))
        if t.dtype != torch.float32:
            raise RuntimeError("endo_b200 expects float32 tensors, got " + str(t.dtype))


def on_device(fn):
    """Run a Function.forward / backward with the CUDA device of its first tensor argument current: the library
    configures per-device kernel attributes lazily for the CURRENT device and borrows its side stream from it
    (include/endo_b200.h), so a model on cuda:1 must not be driven while cuda:0 is current."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            dev = next((t.device for t in getattr(ctx, "saved_tensors", ()) if t.is_cuda), None)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)
    return wrapper


def same_shape(what, ref, *others):
    """Raise (instead of reading out of bounds) when an operand does not have the reference tensor's shape; the
    reference's eager ops would broadcast or reject such inputs."""
    for name, t, shape in others:
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{what}: `{name}` must have shape {tuple(shape)}, got {tuple(t.shape)} "
                               f"(reference tensor: {tuple(ref.shape)})")


def contig(t):
    return t if t.is_contiguous() else t.contiguous()


_ws_cache = {}


def workspace(device, nbytes: int) -> torch.Tensor:
    """Per (device, stream) scratch with a zeroed, self-resetting counter header (see endo_b200.h).

    While the current stream is being CAPTURED into a CUDA graph the scratch is a fresh allocation from the graph's
    private memory pool (its zero-fill is captured with it and replayed): a cached tensor would belong to the pool of
    whichever graph was captured first and be shared by every later graph of the process."""
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros(max(nbytes, 1 << 12), dtype=torch.uint8, device=device)
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def launch_count() -> int:
    return int(lib().endo_launch_count())


class profile:
    """Context manager around endo_prof_*: per-category device milliseconds and launch-site counts."""

    def __enter__(self):
        lib().endo_prof_collect(None, None)      # drop stale records
        lib().endo_prof_enable(1)
        self.ms, self.counts = {}, {}
        return self

    def __exit__(self, *exc):
        l = lib()
        l.endo_prof_enable(0)
        n = l.endo_prof_categories()
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_ulonglong * n)()
        check(l.endo_prof_collect(ms, cnt), "prof_collect")
        for i in range(n):
            name = l.endo_prof_category_name(i).decode()
            self.ms[name], self.counts[name] = float(ms[i]), int(cnt[i])
        return False

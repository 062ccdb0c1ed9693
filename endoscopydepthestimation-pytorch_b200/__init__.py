"""endo-depth-b200: B200-native (sm_100a) implementation of the training hot path of
lppllppl920/EndoscopyDepthEstimation-Pytorch behind the reference's own nn.Module API.

Import as `endo_b200` (shim at the repository root).  `endo_b200.models` and
`endo_b200.losses` mirror `/root/reference/models.py` and `/root/reference/losses.py` for the
classes on the hot path; the arithmetic runs in `csrc/` (hand-written CUDA behind the C ABI
declared in `include/endo_b200.h`).  There is no CPU fallback.
"""
from . import synthetic  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # heavy submodules are imported on first use so that `import endo_b200` stays cheap
    if name in ("models", "losses", "functional", "optim", "ddp", "_lib", "build", "engine", "train_step", "utils"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)

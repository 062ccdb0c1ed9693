"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host-side plan of
the network agrees with the reference layout, the module tree mirrors the reference's state_dict,
and the product has no CPU path."""
import ctypes
import os
import re

import pytest
import torch

import endo_b200
from endo_b200 import _lib
from oracle import net as onet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "endo_b200.h")).read()
    declared = set(re.findall(r"\b(endo_[a-z0-9_]+)\s*\(", header))
    declared -= {"endo_stream_t"}
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/endo_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.endo_version() >= 100
    assert lib.endo_strerror(0) == b"ok" and b"workspace" in lib.endo_strerror(3)


def _cfg(net):
    return ctypes.byref(net._cfg)


@pytest.mark.parametrize("factory,ocfg", [(lambda: endo_b200.models.FCDenseNet57(1), onet.FCDENSENET57),
                                          (lambda: endo_b200.models.FCDenseNet67(1), onet.FCDENSENET67),
                                          (lambda: endo_b200.models.FCDenseNet103(1), onet.FCDENSENET103)])
def test_module_tree_matches_reference_state_dict(factory, ocfg):
    model = factory()
    shapes = onet.param_shapes(ocfg)
    sd = model.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    lib = _lib.lib()
    n_param = sum(v.numel() for k, v in sd.items() if not onet.is_buffer(k))
    n_buf = sum(v.numel() for k, v in sd.items() if k.endswith(("running_mean", "running_var")))
    assert lib.endo_net_param_count(_cfg(model)) == n_param
    assert lib.endo_net_buffer_count(_cfg(model)) == n_buf
    # reference checkpoints load unchanged (also with DataParallel's "module." prefix, utils.py:676)
    state = onet.init_state(ocfg, seed=1, perturb=True)
    model.load_state_dict(state, strict=True)
    prefixed = {"module." + k: v for k, v in state.items()}
    holder = torch.nn.Module()
    holder.module = model
    holder.load_state_dict(prefixed, strict=True)


def test_shape_validation_and_sizes():
    model = endo_b200.models.FCDenseNet57(1)
    lib = _lib.lib()
    assert lib.endo_net_activation_bytes(_cfg(model), 8, 256, 320) > 0
    assert lib.endo_net_activation_bytes(_cfg(model), 8, 250, 320) == 0      # H not a multiple of 32
    assert lib.endo_net_backward_scratch_bytes(_cfg(model), 2, 64, 96) > 0
    per_img = lib.endo_net_activation_bytes(_cfg(model), 1, 256, 320)
    assert 80e6 < per_img < 200e6         # write-once activations: ~0.9e8 B/image vs 7.2e8 layer-by-layer
    # backward scratch of the benchmark shape (16 images of 256x320): gradient buffers of the six levels (channel totals 192 /
    # 240 / 288 / 336 / 384 / 336: [up 48 | in | down-new 48 | up-new 48], the bottleneck has no up part) + the bf16 operand planes of the weight-gradient GEMMs: two sets of
    # activation planes for the widest full-resolution layer (Cin 180 -> 184) and of 48-channel gradient planes, plus one routed-
    # gradient and one activation plane set per TransitionDown (96 / 144 / 192 / 240 / 288 channels); small tables on top
    px = [16 * (256 >> l) * (320 >> l) for l in range(6)]
    grads = 4 * sum(p * c for p, c in zip(px, (192, 240, 288, 336, 384, 336)))
    planes = 2 * (2 * px[0] * 184 + 2 * px[0] * 48) + sum(2 * 2 * p * c for p, c in zip(px, (96, 144, 192, 240, 288)))
    total = lib.endo_net_backward_scratch_bytes(_cfg(model), 16, 256, 320)
    assert grads + planes < total < grads + planes + 64e6, (total, grads, planes)


def test_no_cpu_fallback():
    model = endo_b200.models.FCDenseNet57(1)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        endo_b200.models.DepthWarpingLayer()([torch.zeros(1, 1, 8, 8)] * 3 + [torch.zeros(1, 3, 1), torch.eye(3)[None],
                                                                               torch.eye(3)[None]])
    with pytest.raises(RuntimeError):
        endo_b200.losses.SparseMaskedL1Loss()([torch.zeros(1, 2, 8, 8), torch.zeros(1, 2, 8, 8), torch.zeros(1, 1, 8, 8)])
    with pytest.raises(RuntimeError):
        model.denseBlocksDown[0](torch.zeros(1, 48, 8, 8))     # sub-modules are containers, not an eager path


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: the product package and the C sources must not reference it."""
    pkg = os.path.join(ROOT, "endoscopydepthestimation-pytorch_b200")
    pat = re.compile(r"^\s*(from|import)\s+\.*oracle|oracle\.|/oracle/|\"oracle", re.M)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not pat.search(src), f"{os.path.join(root, f)} references the oracle"


def test_synthetic_batch_shapes():
    b = endo_b200.synthetic.make_batch(2, 64, 96, seed=3)
    assert b["colors_1"].shape == (2, 3, 64, 96) and b["sparse_flows_1"].shape == (2, 2, 64, 96)
    assert b["translations_1_wrt_2"].shape == (2, 3, 1) and b["intrinsics"].shape == (2, 3, 3)
    assert set(b["boundaries"].unique().tolist()) <= {0.0, 1.0}
    r = b["rotations_1_wrt_2"]
    assert torch.allclose(r @ r.transpose(1, 2), torch.eye(3).expand(2, 3, 3), atol=1e-5)
    assert torch.allclose(b["rotations_2_wrt_1"] @ b["translations_1_wrt_2"], -b["translations_2_wrt_1"], atol=1e-6)

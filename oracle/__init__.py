"""CPU oracle for the endo-depth hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch CPU tensors, fp32 or fp64, autograd for the
gradients) of the algorithm the reference runs on its training hot path
(`/root/reference/train.py:272-328`): FCDenseNet-57, DepthScalingLayer, FlowfromDepthLayer,
DepthWarpingLayer, SparseMaskedL1Loss, NormalizedDistanceLoss, ScaleInvariantLoss, the
loss assembly, gradient clipping and the SGD-momentum update.  Every function cites the
reference file:line it restates.

Rules (enforced by tests/test_host.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import anything from here, and only as the checker / the reported CPU baseline;
  * the product package (`endoscopydepthestimation-pytorch_b200/`) never imports it and has no
    CPU fallback: it raises if the CUDA library is missing.

Pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c), and its arithmetic lives in an unpinned PyTorch.  The oracle is therefore
pinned against outputs of the reference itself: `oracle/gen_golden.py` imports the unmodified
`/root/reference/models.py` and `/root/reference/losses.py` in the build container (with the two
shims of SURVEY.md §8c) on seeded inputs and commits the results under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those fixtures.
"""
from . import net, geometry, losses, step, export, pipeline  # noqa: F401

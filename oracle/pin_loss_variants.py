"""TEST INFRASTRUCTURE — container-only: pins NTXentMultiplePositivesV2 / V3 of 3dinfomax_b200/losses.py on the
reference's own classes (commons/losses.py:598-689), cut out of the source file by position (the module imports dgl)
and executed unmodified.  The package classes are evaluated with the CPU oracle's single-positive NTXent standing in
for the CUDA kernels (``cpu_stand_in``), so that the composition around the kernels is what is checked here; the GPU
case then checks the real thing against the vectors this script writes to tests/golden/loss_variants.npz.

    python -m oracle.pin_loss_variants
"""
import contextlib
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("I3D_REFERENCE_ROOT", "/root/reference")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASE = dict(seed=21, B=40, C=3, D=64, tau=0.1)
REG = dict(variance_reg=0.3)          # (cov_loss / uniformity_loss of a [B, C, D] tensor raise in the reference itself)


def reference_classes():
    from torch.nn.modules.loss import _Loss
    src = open(os.path.join(REF, "commons", "losses.py")).read()
    ns = {"torch": torch, "Tensor": torch.Tensor, "_Loss": _Loss}
    a, b = src.index("def uniformity_loss"), src.index("class NTXentShuffled")
    exec(compile(src[a:b], "commons/losses.py[946-964]", "exec"), ns)
    a, b = src.index("class NTXentMultiplePositivesV2"), src.index("class NTXentMultiplePositivesSeparate2D")
    exec(compile(src[a:b], "commons/losses.py[598-689]", "exec"), ns)
    return ns["NTXentMultiplePositivesV2"], ns["NTXentMultiplePositivesV3"]


def inputs():
    g = torch.Generator().manual_seed(CASE["seed"])
    B, C, D = CASE["B"], CASE["C"], CASE["D"]
    z1 = torch.randn(B, D, generator=g) * 0.8
    z2 = torch.randn(B * C, D, generator=g) * 0.5 + 0.2 * z1.repeat_interleave(C, 0)
    return z1, z2


@contextlib.contextmanager
def cpu_stand_in():
    """ops.ntxent -> the CPU oracle (single positive per row, no epsilon: commons/losses.py:225-246 with C = 1)"""
    from oracle import oracle as O
    ops = importlib.import_module("3dinfomax_b200.ops")
    saved = ops.ntxent

    def ntxent(z1, z2, conformers, tau, norm, eps, row_offset=0, total_rows=None):
        assert conformers == 1 and eps == 0.0 and not row_offset and total_rows is None
        return O.ntxent_multiple_positives(z1, z2, tau=tau, norm=norm)

    ops.ntxent = ntxent
    try:
        yield
    finally:
        ops.ntxent = saved


def main():
    V2, V3 = reference_classes()
    L = importlib.import_module("3dinfomax_b200.losses")
    z1, z2 = inputs()
    out = {}
    for tag, Ref, Mine, kw in (("v2", V2, L.NTXentMultiplePositivesV2, {}), ("v3", V3, L.NTXentMultiplePositivesV3, {}),
                               ("v3_reg", V3, L.NTXentMultiplePositivesV3, REG)):
        a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        ref = Ref(tau=CASE["tau"], **kw)(a, b)
        ref.backward()
        a2, b2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        with cpu_stand_in():
            mine = Mine(tau=CASE["tau"], **kw)(a2, b2)
        mine.backward()
        assert abs(mine.item() - ref.item()) <= 2e-6 * abs(ref.item()), (tag, mine.item(), ref.item())
        assert (a2.grad - a.grad).abs().max() <= 2e-5 * a.grad.abs().max(), tag
        assert (b2.grad - b.grad).abs().max() <= 2e-5 * b.grad.abs().max(), tag
        out.update({tag: ref.detach().numpy(), tag + "_dz1": a.grad.numpy(), tag + "_dz2": b.grad.numpy()})
        print("pinned %-6s loss %.6f — package composition == reference class (value and gradients)" % (tag, ref.item()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss_variants.npz"), **out)


if __name__ == "__main__":
    main()

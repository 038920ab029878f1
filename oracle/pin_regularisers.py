"""TEST INFRASTRUCTURE — container-only: pins the NTXent regularisers of 3dinfomax_b200/losses.py on the reference's own
functions.

``uniformity_loss``, ``cov_loss`` and ``std_loss`` are cut out of /root/reference/commons/losses.py by source position
(the module itself imports dgl) and executed unmodified; the regulariser terms are then assembled exactly as
commons/losses.py:157-162 (NTXent) and :250-258 (NTXentMultiplePositives) do.  Values and the gradients with respect to
both embedding matrices go to tests/golden/regularisers.npz; the script asserts that the package's terms agree.

    python -m oracle.pin_regularisers
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("I3D_REFERENCE_ROOT", "/root/reference")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WEIGHTS = dict(variance_reg=0.3, covariance_reg=0.2, uniformity_reg=0.1)
WEIGHTS_MP = dict(variance_reg=0.3, conformer_variance_reg=0.4)
CASE = dict(seed=11, B=48, D=32, C=3)


def reference_functions():
    src = open(os.path.join(REF, "commons", "losses.py")).read()
    start, end = src.index("def uniformity_loss"), src.index("class NTXentShuffled")
    ns = {"torch": torch, "Tensor": torch.Tensor}
    exec(compile(src[start:end], "commons/losses.py[946-964]", "exec"), ns)
    return ns["std_loss"], ns["cov_loss"], ns["uniformity_loss"]


def inputs():
    g = torch.Generator().manual_seed(CASE["seed"])
    B, D, C = CASE["B"], CASE["D"], CASE["C"]
    # scaled so that some per-dimension standard deviations fall below 1 (hinge active) and some above
    scale = torch.linspace(0.3, 1.6, D)
    z1 = torch.randn(B, D, generator=g) * scale
    z2 = torch.randn(B, D, generator=g) * scale
    z2c = (z2[:, None, :] + 0.5 * torch.randn(B, C, D, generator=g) * scale).reshape(B * C, D)
    return z1, z2, z2c


def main():
    std_loss, cov_loss, uniformity_loss = reference_functions()
    L = importlib.import_module("3dinfomax_b200.losses")
    z1, z2, z2c = inputs()
    out = {}
    # ---- NTXent (commons/losses.py:157-162)
    a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    ref = (WEIGHTS["variance_reg"] * (std_loss(a) + std_loss(b)) + WEIGHTS["covariance_reg"] * (cov_loss(a) + cov_loss(b))
           + WEIGHTS["uniformity_reg"] * uniformity_loss(a, b))
    ref.backward()
    a2, b2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    mine = L.NTXent(tau=0.1, **WEIGHTS).regularisers(a2, b2, 1)
    mine.backward()
    assert abs(mine.item() - ref.item()) <= 1e-6 * abs(ref.item()), (mine.item(), ref.item())
    assert (a2.grad - a.grad).abs().max() <= 1e-6 * a.grad.abs().max()
    assert (b2.grad - b.grad).abs().max() <= 1e-6 * b.grad.abs().max()
    out.update(ntxent=ref.detach().numpy(), ntxent_dz1=a.grad.numpy(), ntxent_dz2=b.grad.numpy())
    # ---- NTXentMultiplePositives (commons/losses.py:250-258)
    B, D = z1.shape
    a, b = z1.clone().requires_grad_(True), z2c.clone().requires_grad_(True)
    bv = b.view(B, -1, D)
    ref = WEIGHTS_MP["variance_reg"] * (std_loss(a) + std_loss(bv))
    ref = ref + WEIGHTS_MP["conformer_variance_reg"] * torch.mean(torch.relu(1 - torch.sqrt(bv.var(dim=1) + 1e-04)))
    ref.backward()
    a2, b2 = z1.clone().requires_grad_(True), z2c.clone().requires_grad_(True)
    mine = L.NTXentMultiplePositives(tau=0.1, **WEIGHTS_MP).regularisers(a2, b2, CASE["C"])
    mine.backward()
    assert abs(mine.item() - ref.item()) <= 1e-6 * abs(ref.item()), (mine.item(), ref.item())
    assert (a2.grad - a.grad).abs().max() <= 1e-6 * a.grad.abs().max()
    assert (b2.grad - b.grad).abs().max() <= 1e-6 * b.grad.abs().max()
    out.update(mp=ref.detach().numpy(), mp_dz1=a.grad.numpy(), mp_dz2=b.grad.numpy())
    path = os.path.join(ROOT, "tests", "golden", "regularisers.npz")
    np.savez_compressed(path, **out)
    print("pinned regularisers: NTXent %.6f, MultiplePositives %.6f — package == reference (values and gradients)"
          % (float(out["ntxent"]), float(out["mp"])))


if __name__ == "__main__":
    main()

"""``NTXent`` / ``NTXentMultiplePositives`` — drop-ins for ``loss_func`` (commons/losses.py:126-163, 206-258).

Same constructor arguments and call signature (``loss(z1, z2, **kwargs)``).  The similarity GEMM, the
exp / row-sum / log rows and the whole backward run in lib3dinfomax_b200 kernels.  Under data parallelism the
3-D embeddings are all-gathered first (3dinfomax_b200/dist.py) and ``row_offset`` / ``total_rows`` place the local
rows inside the global negative set.

The optional regularisers (variance / covariance / uniformity, commons/losses.py:157-162, 946-964) have weight 0 in
every target config; non-zero weights are rejected instead of silently ignored.
"""
from torch import nn

from . import ops


class _NTXentBase(nn.Module):
    def __init__(self, norm=True, tau=0.5, uniformity_reg=0, variance_reg=0, covariance_reg=0,
                 conformer_variance_reg=0):
        super().__init__()
        if uniformity_reg or variance_reg or covariance_reg or conformer_variance_reg:
            raise NotImplementedError("NTXent regularisers are 0 in all target configs and have no kernel")
        self.norm, self.tau = norm, tau
        self.uniformity_reg, self.variance_reg, self.covariance_reg = uniformity_reg, variance_reg, covariance_reg
        self.conformer_variance_reg = conformer_variance_reg


class NTXent(_NTXentBase):
    """-mean_i log( P_ii / (sum_j P_ij - P_ii) ),  P = exp(cos_eps(z1_i, z2_j) / tau),  eps = 1e-8 in the denominator."""

    def __init__(self, norm=True, tau=0.5, uniformity_reg=0, variance_reg=0, covariance_reg=0):
        super().__init__(norm, tau, uniformity_reg, variance_reg, covariance_reg)

    def forward(self, z1, z2, row_offset=0, total_rows=None, **kwargs):
        return ops.ntxent(z1, z2, 1, self.tau, self.norm, 1e-8, row_offset, total_rows)


class NTXentMultiplePositives(_NTXentBase):
    """z2 is [B*C, dim], molecule-major; positives of row i are its C conformers; no epsilon in the cosine."""

    def forward(self, z1, z2, row_offset=0, total_rows=None, conformers=None, **kwargs):
        if conformers is None:
            rows = z1.shape[0] if total_rows is None else int(total_rows)   # z2 may be the all-gathered column set
            if z2.shape[0] % rows != 0:
                raise ValueError("z2 rows must be a multiple of the number of molecules")
            conformers = z2.shape[0] // rows
        return ops.ntxent(z1, z2, conformers, self.tau, self.norm, 0.0, row_offset, total_rows)

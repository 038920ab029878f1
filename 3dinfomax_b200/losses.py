"""``NTXent`` / ``NTXentMultiplePositives`` — drop-ins for ``loss_func`` (commons/losses.py:126-163, 206-258).

Same constructor arguments and call signature (``loss(z1, z2, **kwargs)``).  The similarity GEMM, the
exp / row-sum / log rows and the whole backward run in lib3dinfomax_b200 kernels.  Under data parallelism the
3-D embeddings are all-gathered first (3dinfomax_b200/dist.py) and ``row_offset`` / ``total_rows`` place the local
rows inside the global negative set.

The optional regularisers (variance / covariance / uniformity / conformer variance, commons/losses.py:157-162,
250-258, 946-964) have weight 0 in every target config.  With a non-zero weight they are added the way the reference
adds them, as plain tensor arithmetic on the [B, 256] embeddings (a few dozen kilobytes: no kernel of their own);
oracle/pin_regularisers.py pins values and gradients on the reference's own functions.  They see the LOCAL embeddings
only, so they are rejected under data parallelism (the reference has no data-parallel semantics to match).
"""
import torch
from torch import nn

from . import ops


def std_loss(x):
    """hinge on the per-dimension standard deviation over the batch: mean(relu(1 - sqrt(var_0(x) + 1e-4)))
    (commons/losses.py:961-963; unbiased variance over dim 0, so a [B, C, D] input gives C*D hinges)"""
    return torch.relu(1 - torch.sqrt(x.var(dim=0) + 1e-04)).mean()


def cov_loss(x):
    """sum of squared off-diagonal entries of the batch covariance / D (commons/losses.py:953-958)"""
    B, D = x.shape
    xc = x - x.mean(dim=0)
    cov = (xc.t() @ xc) / (B - 1)
    off = cov - torch.diag(torch.diagonal(cov))
    return off.pow(2).sum() / D


def uniformity_loss(x1, x2, t=2):
    """mean over both views of log mean_{i<j} exp(-t ||x_i - x_j||^2) (commons/losses.py:946-951)"""
    def one(x):
        return torch.pdist(x, p=2).pow(2).mul(-t).exp().mean().log()
    return (one(x1) + one(x2)) / 2


class _NTXentBase(nn.Module):
    def __init__(self, norm=True, tau=0.5, uniformity_reg=0, variance_reg=0, covariance_reg=0,
                 conformer_variance_reg=0):
        super().__init__()
        self.norm, self.tau = norm, tau
        self.uniformity_reg, self.variance_reg, self.covariance_reg = uniformity_reg, variance_reg, covariance_reg
        self.conformer_variance_reg = conformer_variance_reg

    def has_regularisers(self):
        return bool(self.uniformity_reg > 0 or self.variance_reg > 0 or self.covariance_reg > 0
                    or self.conformer_variance_reg > 0)

    def regularisers(self, z1, z2, conformers=1):
        """the terms commons/losses.py:157-162 (NTXent) / 250-258 (NTXentMultiplePositives) add to the loss; z2 is
        [B*C, D] molecule-major and is viewed as [B, C, D] when C > 1, exactly like the reference does"""
        z2v = z2.view(z1.shape[0], -1, z1.shape[1]) if conformers > 1 else z2
        reg = z1.new_zeros(())
        if self.variance_reg > 0:
            reg = reg + self.variance_reg * (std_loss(z1) + std_loss(z2v))
        if self.conformer_variance_reg > 0 and conformers > 1:
            reg = reg + self.conformer_variance_reg * torch.relu(1 - torch.sqrt(z2v.var(dim=1) + 1e-04)).mean()
        if self.covariance_reg > 0:
            reg = reg + self.covariance_reg * (cov_loss(z1) + cov_loss(z2v))
        if self.uniformity_reg > 0:
            reg = reg + self.uniformity_reg * uniformity_loss(z1, z2v)
        return reg

    def _with_regularisers(self, loss, z1, z2, conformers, row_offset, total_rows):
        if not self.has_regularisers():
            return loss
        if row_offset or (total_rows is not None and int(total_rows) != z1.shape[0]):
            raise NotImplementedError("NTXent regularisers are defined on one process's embeddings (data parallel: "
                                      "set the weights to 0)")
        return loss + self.regularisers(z1, z2, conformers)


class NTXent(_NTXentBase):
    """-mean_i log( P_ii / (sum_j P_ij - P_ii) ),  P = exp(cos_eps(z1_i, z2_j) / tau),  eps = 1e-8 in the denominator."""

    def __init__(self, norm=True, tau=0.5, uniformity_reg=0, variance_reg=0, covariance_reg=0):
        super().__init__(norm, tau, uniformity_reg, variance_reg, covariance_reg)

    def forward(self, z1, z2, row_offset=0, total_rows=None, **kwargs):
        loss = ops.ntxent(z1, z2, 1, self.tau, self.norm, 1e-8, row_offset, total_rows)
        return self._with_regularisers(loss, z1, z2, 1, row_offset, total_rows)


class NTXentMultiplePositives(_NTXentBase):
    """z2 is [B*C, dim], molecule-major; positives of row i are its C conformers; no epsilon in the cosine."""

    def forward(self, z1, z2, row_offset=0, total_rows=None, conformers=None, **kwargs):
        if conformers is None:
            rows = z1.shape[0] if total_rows is None else int(total_rows)   # z2 may be the all-gathered column set
            if z2.shape[0] % rows != 0:
                raise ValueError("z2 rows must be a multiple of the number of molecules")
            conformers = z2.shape[0] // rows
        loss = ops.ntxent(z1, z2, conformers, self.tau, self.norm, 0.0, row_offset, total_rows)
        return self._with_regularisers(loss, z1, z2, conformers, row_offset, total_rows)


class NTXentMultiplePositivesV3(_NTXentBase):
    """commons/losses.py:646-689 — every conformer slot u is its own NTXent term (no epsilon in the cosine):
    l_iu = -log(p_iiu / (sum_j p_iju - p_iiu)), mean over molecules and slots = the mean over u of the single-positive
    loss of z1 against conformer u of every molecule.  One pass of the NTXent kernels per conformer slot."""

    def __init__(self, norm=True, tau=0.5, uniformity_reg=0, variance_reg=0, covariance_reg=0):
        super().__init__(norm, tau, uniformity_reg, variance_reg, covariance_reg)

    def forward(self, z1, z2, **kwargs):
        B, D = z1.shape
        z2v = z2.view(B, -1, D)
        C = z2v.shape[1]
        loss = sum(ops.ntxent(z1, z2v[:, u, :].contiguous(), 1, self.tau, self.norm, 0.0) for u in range(C)) / C
        return self._with_regularisers(loss, z1, z2, C, 0, None)


class NTXentMultiplePositivesV2(_NTXentBase):
    """commons/losses.py:598-643 — positives: all conformers of the molecule, negatives: the FIRST conformer of the other
    molecules: l_i = -log(sum_u p_iiu / sum_{j != i} p_ij0).  With the single-positive loss of z1 against conformer 0,
    NTXent_i = -log(p_ii0 / sum_{j != i} p_ij0), this is NTXent_i - (log sum_u p_iiu - log p_ii0): the NTXent kernels
    on conformer 0 plus a [B, C] row-wise correction (plain tensor arithmetic on B*C cosines)."""

    def __init__(self, norm=True, tau=0.5, uniformity_reg=0, variance_reg=0, covariance_reg=0):
        super().__init__(norm, tau, uniformity_reg, variance_reg, covariance_reg)

    def forward(self, z1, z2, **kwargs):
        B, D = z1.shape
        z2v = z2.view(B, -1, D)
        C = z2v.shape[1]
        base = ops.ntxent(z1, z2v[:, 0, :].contiguous(), 1, self.tau, self.norm, 0.0)
        pos = (z1[:, None, :] * z2v).sum(dim=2)
        if self.norm:
            pos = pos / (z1.norm(dim=1)[:, None] * z2v.norm(dim=2))
        pos = pos / self.tau
        loss = base - (torch.logsumexp(pos, dim=1) - pos[:, 0]).mean()
        return self._with_regularisers(loss, z1, z2, C, 0, None)

"""FCLayer / MLP with the reference's parameter names, computing through the fused ``ops.fc`` operator.

State-dict compatibility with models/base_layers.py is the contract: every FC layer owns
``linear.{weight,bias}`` and, when normalised, ``batch_norm.{weight,bias,running_mean,running_var,
num_batches_tracked}``; an MLP owns ``fully_connected.{i}``.  Semantics followed: Linear -> activation ->
BatchNorm1d (models/base_layers.py:100-111), xavier_uniform with gain 1/in_dim and zero bias (:93-98), layer
stacking of MLP (:119-142).  Unlike the reference modules these take a list of K-segments (``ops.Seg``), which is
how concatenation / gather / degree scaling get folded into the GEMM.
"""
import math

import torch
from torch import nn

from . import ops
from .kernels import ACT

SUPPORTED_ACTIVATIONS = ("relu", "silu", "none", "leakyrelu")


def activation_code(name):
    key = "none" if name is None else str(name).lower()
    known = {"relu", "sigmoid", "tanh", "elu", "selu", "glu", "leakyrelu", "softplus", "silu", "none"}
    assert key in known, "Unhandled activation function"          # same failure mode as base_layers.py:16
    if key not in SUPPORTED_ACTIVATIONS:
        raise NotImplementedError("activation %r has no sm_100a kernel (target configs use relu / SiLU / none)" % name)
    return ACT[key]


class _Linear(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_dim, in_dim))
        self.bias = nn.Parameter(torch.zeros(out_dim))
        # lets FusedAdam hand this weight a view of its flat gradient buffer that ops._FC accumulates into directly
        self.weight._i3d_direct_grad = True


class _BatchNorm(nn.Module):
    def __init__(self, dim, momentum, eps=1e-5):
        super().__init__()
        self.momentum, self.eps = momentum, eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.register_buffer("running_mean", torch.zeros(dim))
        self.register_buffer("running_var", torch.ones(dim))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class FCLayer(nn.Module):
    def __init__(self, in_dim, out_dim, activation="relu", dropout=0.0, batch_norm=False, batch_norm_momentum=0.1):
        super().__init__()
        if dropout:
            raise NotImplementedError("dropout > 0 is not used by the target configs and has no kernel")
        self.in_dim, self.out_dim = in_dim, out_dim
        self.act = activation_code(activation)
        self.linear = _Linear(in_dim, out_dim)
        self.batch_norm = _BatchNorm(out_dim, batch_norm_momentum) if batch_norm else None
        self.reset_parameters()

    def reset_parameters(self):
        # nn.init.xavier_uniform_(weight, gain=1/in_dim); bias = 0   (models/base_layers.py:93-98)
        bound = (1.0 / self.in_dim) * math.sqrt(6.0 / (self.in_dim + self.out_dim))
        with torch.no_grad():
            self.linear.weight.uniform_(-bound, bound)
            self.linear.bias.zero_()

    def forward(self, segs, residual=None, valid=None):
        if torch.is_tensor(segs):
            segs = [ops.Seg(segs)]
        bn = None
        if self.batch_norm is not None:
            b = self.batch_norm
            bn = (b.weight, b.bias, b.running_mean, b.running_var, b.num_batches_tracked, b.momentum, b.eps)
        return ops.fc(segs, self.linear.weight, self.linear.bias, self.act, bn, self.training, residual, valid)

    def _bn_tuple(self):
        if self.batch_norm is None:
            return None
        b = self.batch_norm
        return (b.weight, b.bias, b.running_mean, b.running_var, b.num_batches_tracked, b.momentum, b.eps)

    def forward_edge_factored(self, g, h, table, valid=None, combo=None):
        """This layer applied to cat[h[src], h[dst], e] in factored form (ops._FCEdgeFactored); ``table``: this layer's
        bond-feature table from ops.bond_tables."""
        return ops.fc_edge_factored(g, h, table, self.linear.weight, self.linear.bias, self.act, self._bn_tuple(),
                                    self.training, valid, combo)

    def forward_merged(self, plan, h, agg, residual=None, valid=None):
        """This layer applied to cat[h, agg, agg*amp, agg*att] through the degree-merged weights (ops._FCPostMerged)."""
        from .kernels import MergedPosttransWeights
        W = self.linear.weight
        # one persistent scratch per (bucket count, device): a captured CUDA graph keeps pointing at the one it was
        # recorded with, so scratch is never re-allocated or shared between different bucket counts
        cache = self.__dict__.setdefault("_merged", {})
        key = (plan.n_buckets, h.shape[1], str(W.device))
        m = cache.get(key)
        if m is None:
            m = cache[key] = MergedPosttransWeights(self.out_dim, h.shape[1], plan.n_buckets, W.device)
        bn = None
        if self.batch_norm is not None:
            b = self.batch_norm
            bn = (b.weight, b.bias, b.running_mean, b.running_var, b.num_batches_tracked, b.momentum, b.eps)
        return ops.fc_post_merged(plan, m, h, agg, W, self.linear.bias, self.act, bn, self.training, residual, valid)


class MLP(nn.Module):
    def __init__(self, in_dim, out_dim, layers, hidden_size=None, mid_activation="relu", last_activation="none",
                 dropout=0.0, mid_batch_norm=False, last_batch_norm=False, batch_norm_momentum=0.1):
        super().__init__()
        self.in_dim, self.hidden_size, self.out_dim = in_dim, hidden_size, out_dim
        fcs = []
        if layers <= 1:
            fcs.append(FCLayer(in_dim, out_dim, last_activation, dropout, last_batch_norm, batch_norm_momentum))
        else:
            fcs.append(FCLayer(in_dim, hidden_size, mid_activation, dropout, mid_batch_norm, batch_norm_momentum))
            for _ in range(layers - 2):
                fcs.append(FCLayer(hidden_size, hidden_size, mid_activation, dropout, mid_batch_norm,
                                   batch_norm_momentum))
            fcs.append(FCLayer(hidden_size, out_dim, last_activation, dropout, last_batch_norm, batch_norm_momentum))
        self.fully_connected = nn.ModuleList(fcs)

    def forward(self, segs, residual=None, valid=None):
        """valid: device int32 scalar — valid leading rows when the batch is padded to a shape bucket (ops.FCConfig)"""
        x = segs
        last = len(self.fully_connected) - 1
        for i, fcl in enumerate(self.fully_connected):
            x = fcl(x, residual if i == last else None, valid)
        return x

"""Contrastive metrics of the pre-training configs, same class names and call signature as trainer/metrics.py.

``PositiveSimilarity``, ``NegativeSimilarity``, ``TruePositiveRate``, ``TrueNegativeRate``, ``ContrastiveAccuracy``
(trainer/metrics.py:240-334,444-463) are what configs_clean/pre-train_QM9.yml:15-20 logs.  The reference evaluates each
one separately — five ``einsum('ik,jk->ij')`` over the same embeddings plus [B,B] masks — here the five share ONE
similarity GEMM and ONE pass over it (``i3d_contrastive_metrics``): the first metric called on a pair of embedding
tensor OBJECTS with a given threshold computes all five, the others read the cached device vector (no host sync anywhere;
the threshold-free similarities default to 0.5, so a config with threshold 0.5009 costs two fused evaluations).

``DimensionCovariance``, ``BatchVariance``, ``Alignment``, ``Uniformity`` (trainer/metrics.py:161-176,212-230) — the other
four entries of that metric list — share one ``i3d_embedding_metrics`` evaluation the same way.

Only the global-vs-global case (``pos_mask is None``) that the target configs use has a kernel; a ``pos_mask`` raises.
"""
import torch
from torch import nn

from . import kernels as K
from . import lib as _lib

_ORDER = ("positive_similarity", "negative_similarity", "true_positive_rate", "true_negative_rate",
          "contrastive_accuracy")


def contrastive_metrics(x1, x2, threshold=0.5):
    """device fp32 [5]: (positive_similarity, negative_similarity, true_positive_rate, true_negative_rate,
    contrastive_accuracy) of embeddings x1 [B,D], x2 [>=B,D] (extra noisy rows of x2 are dropped as in the reference)."""
    if not x1.is_cuda:
        raise RuntimeError("contrastive_metrics needs CUDA tensors: the 3dinfomax_b200 path has no CPU fallback")
    x1 = x1.detach().float().contiguous()
    x2 = x2.detach().float()
    B, D = x1.shape
    if x2.shape[0] != B:
        x2 = x2[:B]                                            # trainer/metrics.py:243-244
    x2 = x2.contiguous()
    n1, n2 = K.row_norms(x1), K.row_norms(x2)
    dot = torch.empty(B, B, dtype=torch.float32, device=x1.device)
    K.gemm(K.NT, B, B, [{"A": x1, "B": x2, "K": D}], dot)
    part = torch.empty(B, 4, dtype=torch.float32, device=x1.device)
    out = torch.empty(5, dtype=torch.float32, device=x1.device)
    _lib.check(_lib.load().i3d_contrastive_metrics(dot.data_ptr(), B, n1.data_ptr(), n2.data_ptr(), float(threshold),
                                                   part.data_ptr(), out.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream), "i3d_contrastive_metrics")
    return out


class _Shared:
    """One fused evaluation per (x1, x2, threshold).  The cache is keyed on the tensor OBJECTS (held, compared with
    ``is``) and their autograd versions — never on raw storage pointers: the reference trainer hands every batch's fresh
    predictions / targets to the metrics (trainer/trainer.py evaluate_metrics), those have equal shapes, version 0 and
    usually the SAME address (caching allocator), so a pointer key would serve batch k's metrics for batch k+1."""
    x1 = x2 = None
    versions = None
    threshold = None
    value = None

    @classmethod
    def get(cls, x1, x2, threshold):
        hit = (cls.x1 is x1 and cls.x2 is x2 and cls.versions == (x1._version, x2._version)
               and cls.threshold == float(threshold))
        if not hit:
            cls.value = contrastive_metrics(x1, x2, threshold)
            cls.x1, cls.x2, cls.versions, cls.threshold = x1, x2, (x1._version, x2._version), float(threshold)
        return cls.value

    @classmethod
    def clear(cls):
        cls.x1 = cls.x2 = cls.versions = cls.threshold = cls.value = None


def embedding_metrics(x1, x2, alpha=2.0, t=2.0):
    """device fp32 [4]: (dimension_covariance, batch_variance, alignment, uniformity) — trainer/metrics.py:161-176,212-230"""
    if not x1.is_cuda:
        raise RuntimeError("embedding_metrics needs CUDA tensors: the 3dinfomax_b200 path has no CPU fallback")
    x1 = x1.detach().float().contiguous()
    x2 = x2.detach().float().contiguous()
    B1, D = x1.shape
    if x2.shape[0] < B1 or x2.shape[1] != D:
        raise ValueError("x2 must have at least as many rows as x1 and the same width")
    ws = torch.empty(8 + 2 * D, dtype=torch.float64, device=x1.device)
    out = torch.empty(4, dtype=torch.float32, device=x1.device)
    _lib.check(_lib.load().i3d_embedding_metrics(x1.data_ptr(), B1, x2.data_ptr(), x2.shape[0], D, float(alpha), float(t),
                                                 ws.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "i3d_embedding_metrics")
    return out


class _Shared4:
    """one fused evaluation of the four embedding metrics per (x1, x2, alpha) — keyed like _Shared on tensor objects"""
    x1 = x2 = None
    key = None
    value = None

    @classmethod
    def get(cls, x1, x2, alpha):
        key = (x1._version, x2._version, float(alpha))
        if not (cls.x1 is x1 and cls.x2 is x2 and cls.key == key):
            cls.value = embedding_metrics(x1, x2, alpha)
            cls.x1, cls.x2, cls.key = x1, x2, key
        return cls.value


class _Metric4(nn.Module):
    index = 0
    alpha = 2.0

    def forward(self, x1, x2, pos_mask=None):
        return _Shared4.get(x1, x2, self.alpha)[self.index]


class DimensionCovariance(_Metric4):
    index = 0


class BatchVariance(_Metric4):
    index = 1


class Alignment(_Metric4):
    index = 2

    def __init__(self, alpha=2):
        super().__init__()
        self.alpha = float(alpha)


class Uniformity(_Metric4):
    """(the reference stores ``t`` but evaluates uniformity_loss with its default t = 2, trainer/metrics.py:223-229)"""
    index = 3

    def __init__(self, t=2):
        super().__init__()
        self.t = t


class _Metric(nn.Module):
    index = 0

    def __init__(self, threshold=0.5):
        super().__init__()
        self.threshold = threshold

    def forward(self, x1, x2, pos_mask=None):
        if pos_mask is not None:
            raise NotImplementedError("local-vs-global metrics (pos_mask) are not used by the target configs")
        return _Shared.get(x1, x2, self.threshold)[self.index]


class PositiveSimilarity(_Metric):
    index = 0

    def __init__(self):
        super().__init__(0.5)


class NegativeSimilarity(_Metric):
    index = 1

    def __init__(self):
        super().__init__(0.5)


class TruePositiveRate(_Metric):
    index = 2


class TrueNegativeRate(_Metric):
    index = 3


class ContrastiveAccuracy(_Metric):
    index = 4

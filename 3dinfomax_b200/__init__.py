"""3dinfomax_b200 — B200-native hot path of 3DInfomax pre-training (PNA + Net3D + NTXent).

The directory name starts with a digit, so import it with ``importlib.import_module("3dinfomax_b200")``.
Public surface = the reference's plugin names (train.py:167-208 looks classes up by name):

    PNA, Net3D                              model_type / model3d_type      (models/pna.py, models/net3d.py)
    PNAOriginal                             model_type (tower PNA)         (models/pna_original.py)
    NTXent, NTXentMultiplePositives[V2|V3]  loss_func                      (commons/losses.py)
    SelfSupervisedTrainer                   trainer: 'contrastive'         (trainer/self_supervised_trainer.py)

Everything computes through lib3dinfomax_b200.so (hand-written sm_100a kernels, C ABI in include/i3d.h).
There is no CPU fallback: constructing modules works anywhere, running them needs a CUDA device.
"""
from .collate import PackedMoleculeStore  # noqa: F401
from .inference import Fingerprinter, fold_batch_norm  # noqa: F401
from .graph import GraphBatch, GraphStructure, annotate_max_in_degree, batch_from_numpy, graph_structure  # noqa: F401
from .losses import NTXent, NTXentMultiplePositives, NTXentMultiplePositivesV2, NTXentMultiplePositivesV3  # noqa: F401
from .metrics import (Alignment, BatchVariance, ContrastiveAccuracy, DimensionCovariance,  # noqa: F401
                      NegativeSimilarity, PositiveSimilarity, TrueNegativeRate, TruePositiveRate, Uniformity,
                      contrastive_metrics, embedding_metrics)
from .net3d import Net3D  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .pna import PNA  # noqa: F401
from .pna_original import PNAOriginal  # noqa: F401
from .trainer import BucketedStep, BucketLadder, CapturedStep, SelfSupervisedTrainer, Trainer  # noqa: F401
from . import lib, synthetic  # noqa: F401

__all__ = ["PNA", "PNAOriginal", "Net3D", "NTXent", "NTXentMultiplePositives", "NTXentMultiplePositivesV2", "NTXentMultiplePositivesV3", "PositiveSimilarity", "NegativeSimilarity",
           "TruePositiveRate", "TrueNegativeRate", "ContrastiveAccuracy", "contrastive_metrics", "DimensionCovariance", "BatchVariance", "Alignment", "Uniformity", "embedding_metrics", "SelfSupervisedTrainer", "Trainer", "CapturedStep", "BucketedStep", "BucketLadder", "FusedAdam",
           "GraphBatch", "GraphStructure", "PackedMoleculeStore", "Fingerprinter", "fold_batch_norm", "annotate_max_in_degree", "batch_from_numpy",
           "graph_structure", "lib", "synthetic"]

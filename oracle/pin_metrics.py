"""TEST INFRASTRUCTURE — container-only: pins oracle.contrastive_metrics on the reference's own metric classes.

Loads trainer/metrics.py UNMODIFIED from /root/reference by file path (stubbing the imports that are not installable:
ogb evaluators, and the dataset / loss modules it only uses for type hints and unrelated metrics), evaluates
PositiveSimilarity, NegativeSimilarity, TruePositiveRate, TrueNegativeRate, ContrastiveAccuracy (threshold 0.5009 as in
train.py) on seeded embeddings and asserts equality with the oracle.  Writes tests/golden/metrics.npz.

    python -m oracle.pin_metrics
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("I3D_REFERENCE_ROOT", "/root/reference")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
THRESHOLD = 0.5009          # train.py: ContrastiveAccuracy(threshold=0.5009) etc.


def load_reference_metrics():
    stubs = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        stubs[name] = m
        return m

    mod("ogb"), mod("ogb.graphproppred", Evaluator=object), mod("ogb.lsc", PCQM4MEvaluator=object)
    c = mod("commons")
    c.__path__ = []
    # cov_loss / uniformity_loss: the reference's own functions, cut out of commons/losses.py by source line (the module
    # itself imports dgl); plain torch, executed unmodified
    src = open(os.path.join(REF, "commons", "losses.py")).read()
    start, end = src.index("def uniformity_loss"), src.index("class NTXentShuffled")
    ns = {"torch": torch, "Tensor": torch.Tensor}
    exec(compile(src[start:end], "commons/losses.py[946-966]", "exec"), ns)
    mod("commons.losses", cov_loss=ns["cov_loss"], uniformity_loss=ns["uniformity_loss"])
    d = mod("datasets")
    d.__path__ = []
    mod("datasets.geom_drugs_dataset", GEOMDrugs=object)
    mod("datasets.qm9_dataset", QM9Dataset=object)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_ref_metrics", os.path.join(REF, "trainer", "metrics.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m


def embeddings(seed, B, D, noisy_rows=0):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, D, generator=g)
    x1 = base + 0.7 * torch.randn(B, D, generator=g)
    x2 = torch.cat([base + 0.7 * torch.randn(B, D, generator=g), torch.randn(noisy_rows, D, generator=g)])
    return x1, x2


CASES = {"b64": (5, 64, 32, 0), "b300_noisy": (6, 300, 256, 300), "b512": (7, 512, 256, 0)}


def main():
    from oracle import oracle as O
    M = load_reference_metrics()
    out = {}
    for name, (seed, B, D, noisy) in CASES.items():
        x1, x2 = embeddings(seed, B, D, noisy)
        ref = torch.stack([M.PositiveSimilarity()(x1, x2), M.NegativeSimilarity()(x1, x2),
                           M.TruePositiveRate(threshold=THRESHOLD)(x1, x2).float(),
                           M.TrueNegativeRate(threshold=THRESHOLD)(x1, x2).float(),
                           M.ContrastiveAccuracy(threshold=THRESHOLD)(x1, x2).float()])
        mine = O.contrastive_metrics(x1, x2, THRESHOLD)
        assert torch.equal(ref, mine), (name, ref, mine)
        out[name + "/ref"] = ref.numpy()
        out[name + "/cfg"] = np.array([seed, B, D, noisy])
        print("pinned metrics %s: %s — oracle == reference (bit exact)" % (name, [round(float(v), 6) for v in ref]))
        # the other four logged metrics (uniformity, alignment, batch_variance, dimension_covariance)
        ref4 = torch.stack([M.DimensionCovariance()(x1, x2), M.BatchVariance()(x1, x2), M.Alignment(alpha=2)(x1, x2),
                            M.Uniformity(t=2)(x1, x2)])
        mine4 = O.embedding_metrics(x1, x2, 2)
        assert torch.equal(ref4, mine4), (name, ref4, mine4)
        out[name + "/ref4"] = ref4.numpy()
        print("pinned embedding metrics %s: %s — oracle == reference (bit exact)" % (name, [round(float(v), 6) for v in ref4]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), threshold=THRESHOLD, **out)


if __name__ == "__main__":
    main()

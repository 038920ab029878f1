"""-m gpu: the CUDA path, called through the C ABI, against the CPU oracle / golden vectors (cases in gpu_cases.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    if not torch.cuda.is_available():
        return []
    import gpu_cases
    return gpu_cases.ALL_CASES


def _names():
    import ast
    import os
    src = open(os.path.join(os.path.dirname(__file__), "gpu_cases.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "ALL_CASES":
            return [e.id for e in node.value.elts]
    return []


@pytest.mark.parametrize("case_name", _names())
def test_case(case_name):
    import gpu_cases
    results = getattr(gpu_cases, case_name)()
    assert results, "case produced no checks"
    failures = ["%s: err %.3e > tol %.1e" % (l, e, t) for l, e, t in results if not (e <= t)]
    assert not failures, "\n".join(failures)


def test_native_library_is_the_compute_path():
    """the product must fail loudly without its extension and must have launched its own kernels"""
    import importlib
    lib = importlib.import_module("3dinfomax_b200.lib")
    K = importlib.import_module("3dinfomax_b200.kernels")
    n0 = lib.launch_count()
    K.add(torch.ones(8, device="cuda"), torch.ones(8, device="cuda"))
    assert lib.launch_count() == n0 + 1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.add(torch.ones(8), torch.ones(8))


def test_aggregate_tma_staged_variant():
    """I3D_AGG_FWD=tma: the bulk-copy (TMA) staged forward aggregation kernel — the opt-in variant of
    i3d_pna_aggregate_fwd — through the same parity cases as the default kernel (kernel level + whole models against
    the reference's golden vectors).  The variant is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, I3D_AGG_FWD="tma")
    r = subprocess.run([sys.executable, os.path.join(here, "gpu_diag.py"), "case_aggregate", "case_golden"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "TOTAL FAILURES: 0" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("== case_") == 2, r.stdout[-2000:]


def test_opt_in_kernel_variants():
    """The measured-and-kept-off variants stay correct: I3D_TN_CLUSTER=4 (split-K reduction of the weight-gradient GEMM
    through a thread-block cluster) and I3D_BN_BWD=fused (BatchNorm backward as one launch with a grid barrier), through
    the GEMM cases and whole training steps (eager, captured, shape-bucketed).  Both switches are read once per process."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, I3D_TN_CLUSTER="4", I3D_BN_BWD="fused")
    cases = ["case_gemm_tc", "case_train_steps", "case_train_steps_captured", "case_bucketed_step"]
    r = subprocess.run([sys.executable, os.path.join(here, "gpu_diag.py")] + cases, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "TOTAL FAILURES: 0" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("== case_") == len(cases), r.stdout[-2000:]

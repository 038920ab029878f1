"""Runs every parity case and prints every number (does not stop at the first failure).
usage (GPU box):  python tests/gpu_diag.py [case_name ...] > gpurun_out/diag.txt"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    import gpu_cases
    print("device:", torch.cuda.get_device_name(0), "| torch", torch.__version__, flush=True)
    want = set(sys.argv[1:])
    n_fail = 0
    for case in gpu_cases.ALL_CASES:
        if want and case.__name__ not in want:
            continue
        t0 = time.time()
        try:
            res = case()
            torch.cuda.synchronize()
        except Exception:
            n_fail += 1
            print("CASE %s RAISED\n%s" % (case.__name__, traceback.format_exc()), flush=True)
            continue
        print("== %s (%.1fs)" % (case.__name__, time.time() - t0), flush=True)
        for label, err, tol in res:
            ok = err <= tol
            n_fail += 0 if ok else 1
            print("  %-4s %-70s err %.3e  tol %.1e" % ("ok" if ok else "FAIL", label, err, tol), flush=True)
    print("TOTAL FAILURES:", n_fail, flush=True)
    return 1 if n_fail else 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Torch-free runner of the checks of tests/test_gpu_zz_bases.py (ctypes + numpy only: starts in a few seconds)."""
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import test_gpu_zz_bases as T  # noqa: E402  (pytest is importable without torch)

ok = True
for case in T.CASES:
    try:
        errs = T.check_case(case)
        good = max(errs.values()) < T.TOL
    except Exception as e:  # noqa: BLE001
        errs, good = repr(e), False
    ok &= good
    print("PASS" if good else "FAIL", case, errs, flush=True)
print("ALL PASS" if ok else "SOME FAILED", f"{time.time() - t0:.1f} s")

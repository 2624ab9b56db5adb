import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
from problems import fem_problem, mt_vector
from oracle import Oracle
from pdelab_b200.capi import GridOperator
for case in [dict(cells=(5,4,3), degree=1, a="full", with_b=True, with_c=True),
             dict(cells=(5,4,3), degree=1, a="scalar", bc="mixed"),
             dict(cells=(5,4,3), degree=1, a="scalar", bc="mixed", with_b=True),
             dict(cells=(5,4,3), degree=1, a="full", bc="mixed"),
             dict(cells=(5,4,3), degree=1, a="scalar", bc="mixed", with_c=True),
             dict(cells=(5,4,3), degree=1, a="full", bc="mixed", with_b=True, with_c=True)]:
    spec = fem_problem(**case)
    x = mt_vector(spec.num_dofs)
    a = GridOperator(spec).residual(x, np.zeros_like(x)); b = Oracle(spec).residual(x)
    d = np.abs(a-b); print(case, d.max()/np.abs(b).max(), np.flatnonzero(d > 1e-10*np.abs(b).max()))
    a = GridOperator(spec).jacobian_apply(x, np.zeros_like(x)); b = Oracle(spec).jacobian_apply(x)
    d = np.abs(a-b); print("   apply", d.max()/np.abs(b).max(), np.flatnonzero(d > 1e-10*np.abs(b).max()))

#!/usr/bin/env python
"""One assembled-Jacobian SpMV of cfg4 (Q2 160^3) for ncu (tuning helper)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
from bench_configs import rand
C = int(sys.argv[1]) if len(sys.argv) > 1 else 160
cells = (C, C, C)
nc = C ** 3
kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=2, a_mode=abi.A_SCALAR, A=kappa)
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
n = spec.num_dofs
nr, nnz = go.pattern_size()
x = rand(n, 2)
vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
go.jacobian(x, vals, fresh=True)
y = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(2):
    go.csr_mv(vals, x, y)
torch.cuda.synchronize()

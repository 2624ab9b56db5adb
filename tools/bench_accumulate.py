#!/usr/bin/env python
"""Times the three vector entry points of the DG k=2 128^3 operator: apply (y = J x), jacobian_apply
(y += J x) and residual (r += J x + R(0)); tuning helper for the accumulate epilogue."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
from bench_configs import timeit, rand

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cells = (C, C, C)
nc = C ** 3
kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
spec = abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=k, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa,
                       f=rand(nc * (k + 1) ** 3, 1))
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
n = spec.num_dofs
x, y = rand(n, 2), torch.zeros(n, dtype=torch.float64, device="cuda")
for name, fn in (("apply", lambda: go.apply(x, y)), ("jacobian_apply", lambda: go.jacobian_apply(x, y)),
                 ("residual", lambda: go.residual(x, y))):
    ms = timeit(fn, 20)
    print(f"k={k} {C}^3 {name}: {ms:.4f} ms  kernel={go.last_kernel()}")

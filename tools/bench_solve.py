#!/usr/bin/env python
"""Times the device-resident Krylov solvers (csrc/krylov.cu) on the headline operator: DG k=2 SIPG on
C^3 cells, matrix-free CG / BiCGSTAB to a fixed number of iterations, vectors resident on the device.
Reports ms per iteration and the share of the operator applications.  One JSON object per line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
from bench_configs import timeit, rand

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
nc = C ** 3
kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
spec = abi.ProblemSpec((C, C, C), space=abi.SPACE_QKDG, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa)
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
n = spec.num_dofs
b = rand(n, 2)
z = torch.zeros_like(b)
y = torch.empty_like(b)
ms_apply = timeit(lambda: go.apply(b, y), 50)
ms_bj = timeit(lambda: go.block_jacobi_apply(b, y), 50)
print(json.dumps(dict(op="block_jacobi_apply", cells=[C] * 3, dofs=n, ms=round(ms_bj, 4), ms_apply=round(ms_apply, 4),
                      note="exact D^-1 r by fast diagonalisation; 16 B/DOF vectors + 36 doubles per cell",
                      algorithmic_GBs=round((16.0 * n + 36 * 8.0 * nc) / (ms_bj * 1e-3) / 1e9, 1))))
for name, solver, applies, precond in (("CG", abi.SOLVER_CG, 1, abi.PRECOND_NONE),
                                       ("BiCGSTAB", abi.SOLVER_BICGSTAB, 2, abi.PRECOND_NONE),
                                       ("CG+BlockJacobi", abi.SOLVER_CG, 1, abi.PRECOND_BLOCK_JACOBI)):
    for _ in range(2):  # first call allocates the work vectors
        z.zero_()
        r = b.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = go.solve(z, r, 1e-30, solver=solver, maxiter=iters, precond=precond)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ms_it = dt * 1e3 / res["iterations"]
    vec_bytes = 8.0 * n
    print(json.dumps(dict(solver=name, cells=[C] * 3, dofs=n, iterations=res["iterations"], ms_total=round(dt * 1e3, 3),
                          ms_per_iteration=round(ms_it, 4), ms_apply=round(ms_apply, 4),
                          apply_share=round(applies * ms_apply / ms_it, 3),
                          dof_iterations_per_s=n / (ms_it * 1e-3), reduction=res["reduction"],
                          vector_passes_equiv_GBs=round((ms_it - applies * ms_apply) and
                                                        vec_bytes / ((ms_it - applies * ms_apply) * 1e-3) / 1e9, 1))))

# time to solution: reduce the defect by 1e-8 with and without the block-Jacobi preconditioner
for name, precond in (("CG", abi.PRECOND_NONE), ("CG+BlockJacobi", abi.PRECOND_BLOCK_JACOBI)):
    z.zero_()
    r = b.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = go.solve(z, r, 1e-8, solver=abi.SOLVER_CG, maxiter=20000, precond=precond)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(solve="reduction 1e-8", solver=name, cells=[C] * 3, dofs=n, converged=res["converged"],
                          iterations=res["iterations"], seconds=round(dt, 4), reduction=res["reduction"])))

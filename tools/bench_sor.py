#!/usr/bin/env python
"""Times the matrix-free block preconditioners on the headline operator (DG k=2 SIPG, C^3 cells): one block-Jacobi
application, one block SOR sweep (hyperplane wavefronts), and the time to reduce the defect by 1e-8 with
BiCGSTAB + block Jacobi / block SOR and CG + block Jacobi / symmetric block SOR.  One JSON object per line."""
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

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nc = C ** 3
kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
spec = abi.ProblemSpec((C, C, C), space=abi.SPACE_QKDG, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa)
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
n = spec.num_dofs
b = rand(n, 2)
y = torch.empty_like(b)
print(json.dumps(dict(op="apply / block_jacobi_apply / block_sor_apply (one forward sweep)", cells=[C] * 3, dofs=n,
                      ms_apply=round(timeit(lambda: go.apply(b, y), 30), 4),
                      ms_block_jacobi=round(timeit(lambda: go.block_jacobi_apply(b, y), 30), 4),
                      ms_block_sor_sweep=round(timeit(lambda: go.block_sor_apply(b, y), 5), 4),
                      wavefronts=3 * C - 2)), flush=True)
for name, solver, precond in (("BiCGSTAB+BlockJacobi", abi.SOLVER_BICGSTAB, abi.PRECOND_BLOCK_JACOBI),
                              ("BiCGSTAB+BlockSOR", abi.SOLVER_BICGSTAB, abi.PRECOND_BLOCK_SOR),
                              ("CG+BlockJacobi", abi.SOLVER_CG, abi.PRECOND_BLOCK_JACOBI),
                              ("CG+BlockSSOR", abi.SOLVER_CG, abi.PRECOND_BLOCK_SSOR)):
    z = torch.zeros_like(b)
    r = b.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = go.solve(z, r, 1e-8, solver=solver, maxiter=20000, precond=precond)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(solve=name, cells=[C] * 3, dofs=n, reduction_target=1e-8, converged=res["converged"],
                          iterations=res["iterations"], seconds=round(dt, 4),
                          ms_per_iteration=round(dt * 1e3 / max(res["iterations"], 1), 4))), flush=True)

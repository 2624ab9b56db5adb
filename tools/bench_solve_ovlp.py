#!/usr/bin/env python
"""Weak-scaling bench of the overlapping Krylov solvers (pdb200_solve_ovlp): DG k=2 SIPG, C^3 owned cells per GPU,
matrix-free CG / CG + block Jacobi / BiCGSTAB for a fixed number of iterations; every vector, the halo exchange
(NVLink mailboxes) and the global sums (peer mailboxes) stay on the devices.  Launch like bench.py:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/bench_solve_ovlp.py [C] [iterations]
Rank 0 prints one JSON line per solver: ms per iteration = max over ranks of the device time of the solve call."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch
import torch.distributed as dist
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
from pdelab_b200.partition import OverlappingPartition, OverlappingSolverBackend, exchange_cell_field

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
part = OverlappingPartition.weak((C, C, C), world, rank, overlap=1)
ncl = int(np.prod(part.local_cells))
g = torch.Generator(device=dev).manual_seed(42 + rank)
kappa = 10.0 ** (2.0 * torch.rand(ncl, dtype=torch.float64, device=dev, generator=g) - 1.0)
if world > 1:
    exchange_cell_field(kappa.view(part.local_cells[::-1]), part, dist)
spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QKDG, degree=2, lower=part.local_lower, upper=part.local_upper,
                       alpha=3.0, a_mode=abi.A_SCALAR, A=kappa, side_kind=part.side_kind, device=local_rank)
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
n = spec.num_dofs
owned = int(np.prod(part.owned_cells)) * 27
b = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
z = torch.zeros_like(b)
ls = OverlappingSolverBackend(go, part, dist) if world > 1 else None


def solve(solver, precond):
    z.zero_()
    r = b.clone()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if ls is None:
        res = go.solve(z, r, 1e-30, solver=solver, precond=precond, maxiter=iters)
    else:
        ls.solver, ls.precond, ls.maxiter = solver, precond, iters
        res = ls.apply(z, r, 1e-30)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), res


for name, solver, precond in (("CG", abi.SOLVER_CG, abi.PRECOND_NONE), ("CG+BlockJacobi", abi.SOLVER_CG, abi.PRECOND_BLOCK_JACOBI),
                              ("BiCGSTAB", abi.SOLVER_BICGSTAB, abi.PRECOND_NONE)):
    solve(solver, precond)  # allocates the work vectors
    ms, res = solve(solver, precond)
    if rank == 0:
        print(json.dumps(dict(solver=name, n_gpus=world, cells_per_gpu=[C] * 3, global_dofs=owned * world,
                              iterations=res["iterations"], ms_per_iteration=round(ms / max(res["iterations"], 1), 4),
                              dof_iterations_per_s=owned * world / (ms / max(res["iterations"], 1) * 1e-3),
                              reduction=res["reduction"],
                              comm="halo: NVLink mailboxes hidden behind the interior tiles; sums: peer mailboxes" if world > 1 else "none")),
              flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

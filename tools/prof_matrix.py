"""One assembled-Jacobian + pattern + SpMV + FEM residual pass on conforming Q2 3D (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
cells = (N, N, N)
nc = N ** 3
g = torch.Generator(device="cuda").manual_seed(0)
kappa = 10.0 ** (2.0 * torch.rand(nc, dtype=torch.float64, device="cuda", generator=g) - 1.0)
f = torch.rand(nc * 27, dtype=torch.float64, device="cuda", generator=g)
spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=2, a_mode=abi.A_SCALAR, A=kappa, f=f)
go = GridOperator(spec)
nr, nnz = go.pattern_size()
x = torch.rand(nr, dtype=torch.float64, device="cuda", generator=g)
r = torch.zeros_like(x)
vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
rowptr = torch.empty(nr + 1, dtype=torch.int64, device="cuda")
colidx = torch.empty(nnz, dtype=torch.int32, device="cuda")
for _ in range(2):
    go.fill_pattern(rowptr=rowptr, colidx=colidx, index32=True)
    go.jacobian(x, vals, fresh=True)
    go.csr_mv(vals, x, r)
    go.residual(x, r)
torch.cuda.synchronize()
print("ok", nr, nnz)

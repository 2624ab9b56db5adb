#!/usr/bin/env python
"""tuning helper: cfg2 with convection and reaction (b = (1, 0.5, 0.25), c = 1) and without, same process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dune-pdelab_b200", "python"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
from bench import _time_events, ALPHA
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator

dev = torch.device("cuda:0")
def rand(n, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    return torch.rand(n, dtype=torch.float64, device=dev, generator=g)
cells = (128, 128, 128); nc = 128 ** 3
kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
bvec = torch.tensor([1.0, 0.5, 0.25], dtype=torch.float64, device=dev).repeat(nc, 1).contiguous()
out = []
for name, kw in (("b=0", {}), ("b,c", dict(b=bvec, c=torch.ones(nc, dtype=torch.float64, device=dev)))):
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=2, alpha=ALPHA, a_mode=abi.A_SCALAR, A=kappa, **kw)
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    n = spec.num_dofs
    x, r = rand(n, 2), torch.zeros(n, dtype=torch.float64, device=dev)
    ms = _time_events(torch, lambda: go.apply(x, r), 100, warm=10)
    out.append("%s %.4f ms" % (name, ms))
print(os.environ.get("PDB200_LIB", "head").split("_")[-1], " | ".join(out), flush=True)

#!/usr/bin/env python
"""Times the non-headline configurations of BASELINE.json (cfg1, cfg3, cfg4 and the FEM vector
kernels) on one GPU with CUDA events and reports achieved algorithmic GB/s.
Not the driver's bench (that is bench.py); output: one JSON object per line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch

from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, alg_bytes, units, unit_name, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    line = dict(config=name, ms=round(ms, 4), units_per_s=units / (ms * 1e-3), unit=unit_name,
                algorithmic_GBs=round(gbs, 1), frac_of_measured_hbm_peak=round(gbs / PEAK, 4), **extra)
    print(json.dumps(line), flush=True)


def rand(n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand(n, dtype=torch.float64, device="cuda", generator=g)


def fem_case(name, cells, k, reps):
    nc = int(np.prod(cells))
    kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=k, a_mode=abi.A_SCALAR, A=kappa,
                           f=rand(nc * (k + 1) ** len(cells), 1))
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    n = spec.num_dofs
    x, r = rand(n, 2), torch.zeros(n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: go.residual(x, r), reps)
    report(name + " residual", ms, 24.0 * n + 8.0 * nc * (1 + (k + 1) ** len(cells)), n, "DOF/s", dofs=n,
           note="accumulate form: read x, read+write r (24 B/DOF) + kappa and f per cell")
    ms = timeit(lambda: go.apply(x, r), reps)
    report(name + " jacobian_apply (y = J x)", ms, 16.0 * n + 8.0 * nc, n, "DOF/s", dofs=n)
    nr, nnz = go.pattern_size()
    rowptr = torch.empty(nr + 1, dtype=torch.int64, device="cuda")
    colidx = torch.empty(nnz, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: go.fill_pattern(rowptr=rowptr, colidx=colidx, index32=True), max(1, reps // 4), warm=1)
    report(name + " pattern (colidx u32 + rowptr u64)", ms, 4.0 * nnz + 8.0 * (nr + 1), nnz, "nnz/s", nnz=nnz)
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: go.jacobian(x, vals, fresh=True), max(1, reps // 4), warm=1)
    report(name + " jacobian (A = 0; jacobian)", ms, 8.0 * nnz + 8.0 * nc, nnz, "nnz/s", nnz=nnz)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: go.csr_mv(vals, x, y), max(1, reps // 4), warm=1)
    report(name + " SpMV with arithmetic column decode", ms, 8.0 * nnz + 16.0 * n, nnz, "nnz/s")
    del go, vals, colidx


def dg_case(name, cells, k, reps, residual):
    nc = int(np.prod(cells))
    kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
    kw = dict(f=rand(nc * (k + 1) ** len(cells), 1)) if residual else {}
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=k, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa, **kw)
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    n = spec.num_dofs
    x, r = rand(n, 2), torch.zeros(n, dtype=torch.float64, device="cuda")
    if residual:
        ms = timeit(lambda: go.residual(x, r), reps)
        report(name + " residual", ms, 24.0 * n + 8.0 * nc + 8.0 * n, n, "DOF/s", dofs=n, kernel=go.last_kernel(),
               note="read x, read+write r, f at the (k+1)^3 points per cell, kappa per cell")
    ms = timeit(lambda: go.apply(x, r), reps)
    report(name + " jacobian_apply (y = J x)", ms, 16.0 * n + 8.0 * nc, n, "DOF/s", dofs=n, kernel=go.last_kernel())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="cfg1,fem3d,cfg3,dg2d,cfg4")
    args = ap.parse_args()
    which = args.which.split(",")
    if "cfg1" in which:
        fem_case("cfg1 Q1 2D 256^2", (256, 256), 1, 200)
        fem_case("Q1 2D 4096^2", (4096, 4096), 1, 20)
    if "fem3d" in which:
        fem_case("Q1 3D 256^3", (256, 256, 256), 1, 10)
        fem_case("Q2 3D 96^3", (96, 96, 96), 2, 10)
    if "cfg3" in which:
        dg_case("cfg3 DG k=4 3D 64^3", (64, 64, 64), 4, 5, True)
        dg_case("DG k=1 3D 128^3", (128, 128, 128), 1, 5, True)
        dg_case("DG k=2 3D 128^3 generic", (128, 128, 128), 2, 5, True)
    if "dg2d" in which:  # the reference's own DG test configurations (k = 1, 2-D), at size
        dg_case("DG k=1 2D 4096^2", (4096, 4096), 1, 10, True)
        dg_case("DG k=2 2D 2048^2", (2048, 2048), 2, 10, True)
    if "cfg4" in which:
        fem_case("cfg4 Q2 3D 160^3", (160, 160, 160), 2, 4)


if __name__ == "__main__":
    main()

"""Times the fast DG kernel on one GPU for arbitrary local boxes / side kinds (tuning helper)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator

def run(cells, side_kind, reps=200):
    nc = int(np.prod(cells))
    g = torch.Generator(device="cuda").manual_seed(0)
    kappa = 10.0 ** (2.0 * torch.rand(nc, dtype=torch.float64, device="cuda", generator=g) - 1.0)
    spec = abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa, side_kind=side_kind)
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    z = torch.rand(spec.num_dofs, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty_like(z)
    for _ in range(10):
        go.apply(z, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go.apply(z, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"cells": cells, "side_kind": side_kind, "ms": round(ms, 4), "kernel": go.last_kernel()}), flush=True)

D = [[0, 0], [0, 0], [0, 0]]
CASES = [((128, 128, 128), D), ((128, 128, 129), D), ((128, 128, 129), [[0, 0], [0, 0], [0, 1]]),
         ((128, 128, 129), [[0, 0], [0, 0], [1, 0]]), ((128, 128, 130), [[0, 0], [0, 0], [1, 1]]), ((128, 128, 132), D),
         ((128, 130, 130), [[0, 0], [1, 1], [1, 1]]), ((128, 129, 130), [[0, 0], [0, 1], [1, 1]]), ((128, 128, 128), D),
         ((128, 130, 130), D), ((128, 130, 128), [[0, 0], [1, 1], [0, 0]]), ((128, 132, 128), D), ((128, 128, 132), D)]
# the kernel runs into sw_power_cap after ~0.1 s: compare cases from a cold start, one case per process
# (python tools/bench_box.py 2; sleep 5; python tools/bench_box.py 0; ...)
for i in ([int(a) for a in sys.argv[1:]] or range(len(CASES))):
    run(*CASES[i])

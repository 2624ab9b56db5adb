"""Does the placement of z / y in memory matter?  Same kernel, same sizes, different offsets."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
cells = (128, 128, 128)
nc = int(np.prod(cells))
g = torch.Generator(device="cuda").manual_seed(0)
kappa = 10.0 ** (2.0 * torch.rand(nc, dtype=torch.float64, device="cuda", generator=g) - 1.0)
spec = abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa)
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
n = spec.num_dofs
pool = torch.empty(3 * n + (1 << 24), dtype=torch.float64, device="cuda")
pool.uniform_(generator=g)
def t(zoff, yoff, reps=200):
    z = pool[zoff:zoff + n]; y = pool[yoff:yoff + n]
    for _ in range(10): go.apply(z, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go.apply(z, y)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("pool base % 2MB:", pool.data_ptr() % (1 << 21))
for rep in range(2):
    for gap in (0, 2, 1 << 10, 1 << 14, 1 << 17, 1 << 18, 1 << 20, (1 << 20) + (1 << 13), 3 << 19, 1 << 22):
        print(json.dumps({"gap_doubles": gap, "ms": round(t(0, n + gap), 4)}), flush=True)

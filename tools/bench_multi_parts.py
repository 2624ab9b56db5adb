"""Where does a multi-GPU step go?  (run under torchrun)  Times, in the sustained regime: the full kernel,
the interior and boundary parts, the p2p exchange alone and the overlapped apply."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch, torch.distributed as dist
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
from pdelab_b200.partition import OverlappingPartition, P2PHaloExchanger
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
part = OverlappingPartition.weak((128, 128, 128), world, rank, overlap=1)
ncl = int(np.prod(part.local_cells))
g = torch.Generator(device=dev).manual_seed(42 + rank)
kappa = 10.0 ** (2.0 * torch.rand(ncl, dtype=torch.float64, device=dev, generator=g) - 1.0)
spec = abi.ProblemSpec(part.local_cells, degree=2, lower=part.local_lower, upper=part.local_upper, alpha=3.0,
                       a_mode=abi.A_SCALAR, A=kappa, side_kind=part.side_kind, device=lr)
go = GridOperator(spec); go.set_stream(torch.cuda.current_stream().cuda_stream)
halo = P2PHaloExchanger(go, part, dist)
z = torch.rand(spec.num_dofs, dtype=torch.float64, device=dev, generator=g); y = torch.empty_like(z)
def timeit(fn, reps=100):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return round(mn.item(), 4), round(mx.item(), 4)
t0 = time.time()
while time.time() - t0 < 0.6:   # reach the power-capped steady state
    for _ in range(50): go.apply(z, y)
    torch.cuda.synchronize()
res = {}
for name, fn in [("all", lambda: go.apply(z, y)), ("interior", lambda: go.apply_part(z, y, abi.PART_INTERIOR)),
                 ("boundary", lambda: go.apply_part(z, y, abi.PART_BOUNDARY)), ("exchange", lambda: halo.exchange(z)),
                 ("apply_p2p", lambda: halo.apply(z, y)), ("all_again", lambda: go.apply(z, y))]:
    res[name] = timeit(fn)
go.synchronize()
if rank == 0: print(json.dumps({"world": world, "procs": part.procs, "min_max_ms": res}))
dist.destroy_process_group()

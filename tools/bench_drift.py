"""Burst vs sustained: time 100-step chunks of the headline kernel for a few seconds and log NVML clocks."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dune-pdelab_b200/python", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch, pynvml
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
cells = (128, 128, 128)
nc = int(np.prod(cells))
g = torch.Generator(device="cuda").manual_seed(0)
kappa = 10.0 ** (2.0 * torch.rand(nc, dtype=torch.float64, device="cuda", generator=g) - 1.0)
spec = abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa)
go = GridOperator(spec)
go.set_stream(torch.cuda.current_stream().cuda_stream)
z = torch.rand(spec.num_dofs, dtype=torch.float64, device="cuda", generator=g)
y = torch.empty_like(z)
for _ in range(3): go.apply(z, y)
torch.cuda.synchronize()
t00 = time.time()
for chunk in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): go.apply(z, y)
    e1.record()
    sm = pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM); mem = pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_MEM)
    pw = pynvml.nvmlDeviceGetPowerUsage(hd) / 1000; rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hd)
    torch.cuda.synchronize()
    print(f"t={time.time()-t00:6.3f}s chunk {chunk:2d}: {e0.elapsed_time(e1)/100:.4f} ms  sm {sm} mem {mem} MHz  {pw:.0f} W reasons {rs:#x}", flush=True)
    if chunk == 19:
        time.sleep(2.0); print("-- slept 2 s --")

"""Time the fused one-step stage operator (csrc/onestep.cu) against the unfused equivalent on config 2
(DG k=2, 128^3 cells): one stage apply  y = (b_rr dt J0 + M) z
  fused    : one launch of dg_fast_q2_3d on the stage operator (mass term in the reaction slot)
  unfused  : J0 z, M z with the two operators + one axpby pass (what running the two assemblers side by side costs)
and the pre-stage / residual calls.  Prints one JSON line per measurement."""
import json
import sys

import torch

sys.path.insert(0, "dune-pdelab_b200/python")
sys.path.insert(0, "tests")
from pdelab_b200 import onestep as osm  # noqa: E402
from pdelab_b200.capi import GridOperator  # noqa: E402
from problems import dg_problem  # noqa: E402


def timeit(fn, warm=10, reps=100):
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


def main():
    n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    spec0 = dg_problem((n1, n1, n1), degree=2, a="scalar", with_f=True)
    go0, go1 = GridOperator(spec0), GridOperator(osm.l2_spec(spec0))
    igo = osm.OneStepGridOperator(go0, go1)
    method = osm.Alexander2Parameter()
    n = spec0.num_dofs
    g = torch.Generator("cuda").manual_seed(1)
    z = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    x1 = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    y, y0, y1 = torch.empty_like(z), torch.empty_like(z), torch.empty_like(z)
    dt = 1e-3
    igo.preStep(method, 0.0, dt)
    igo.preStage(1, [z])
    w0 = method.b(1, 1) * dt

    def unfused():
        go0.apply(z, y0)
        go1.apply(z, y1)
        torch.add(y1, y0, alpha=w0, out=y)

    out = dict(config=f"DG k=2 {n1}^3", dofs=n)
    out["stage_apply_fused_ms"] = timeit(lambda: igo.apply(z, y))
    out["stage_apply_unfused_ms"] = timeit(unfused)
    out["stationary_apply_ms"] = timeit(lambda: go0.apply(z, y))
    r = torch.zeros_like(z)
    out["stage_residual_ms"] = timeit(lambda: igo.residual(z, r), warm=3, reps=30)
    out["pre_stage_2_ms"] = timeit(lambda: igo.preStage(2, [z, x1]), warm=2, reps=10)
    out["dof_per_s_fused"] = n / (out["stage_apply_fused_ms"] * 1e-3)
    out["kernel"] = igo.stage_operator().last_kernel()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

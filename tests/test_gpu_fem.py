"""Parity of the CUDA conforming-Qk path (residual, exact jacobian_apply, constraints, DOF
numbering) with the CPU oracle (-m gpu, through the C ABI)."""
import numpy as np
import pytest

from pdelab_b200 import abi
from problems import fem_problem, mt_vector, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    dict(cells=(5, 4), degree=1), dict(cells=(17, 9), degree=1, a="full", with_b=True, with_c=True),
    dict(cells=(5, 4), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(33, 18), degree=2, a="diagonal", extent=(1.0, 0.7)),
    dict(cells=(1, 1), degree=1, a="identity"), dict(cells=(1, 1, 1), degree=2, a="identity"),
    dict(cells=(3, 2, 1), degree=1, a="full", with_b=True, with_c=True),
    dict(cells=(9, 5, 6), degree=1, a="scalar", extent=(1.0, 0.7, 1.3)),
    dict(cells=(4, 3, 2), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(9, 5, 6), degree=2, a="scalar"),
    dict(cells=(6, 5), degree=1, a="scalar", bc="mixed", with_b=True),
    dict(cells=(18, 7), degree=2, a="scalar", bc="mixed", with_b=True),
    dict(cells=(5, 4, 3), degree=1, a="full", bc="mixed", with_b=True, with_c=True),
    dict(cells=(9, 5, 5), degree=2, a="diagonal", bc="mixed", with_b=True),
    dict(cells=(4, 3, 2), degree=2, a="scalar", intorderadd=1),
]
_id = lambda c: "-".join(f"{k}={v}" for k, v in c.items())


def _ops(spec):
    from oracle import Oracle
    from pdelab_b200.capi import GridOperator
    return GridOperator(spec), Oracle(spec)


@pytest.mark.parametrize("case", CASES, ids=_id)
def test_fem_residual_and_apply_match_oracle(cuda_lib, case):
    spec = fem_problem(**case)
    go, orc = _ops(spec)
    n = spec.num_dofs
    assert go.globalSizeU() == orc.num_dofs == n
    x = mt_vector(n)
    r0 = mt_vector(n, seed=9)
    assert rel_err(go.residual(x, r0.copy()), orc.residual(x, r0.copy())) < TOL
    kron = case.get("a", "scalar") != "full" and not case.get("with_b", False)   # Kronecker cell integral + cached R(0)
    assert go.last_kernel() == ("fem_kron+r0" if kron else "fem_residual")
    assert rel_err(go.residual(x, r0.copy()), orc.residual(x, r0.copy())) < TOL   # second call: cached R(0)
    assert rel_err(go.jacobian_apply(x, r0.copy()), orc.jacobian_apply(x, r0.copy())) < TOL
    y = go.apply(x, np.full(n, np.nan))            # OnTheFlyOperator::apply: y = J x
    assert rel_err(y, orc.jacobian_apply(x)) < TOL


@pytest.mark.parametrize("case", [dict(cells=(5, 4), degree=2), dict(cells=(4, 3, 2), degree=2, bc="mixed"),
                                  dict(cells=(3, 3, 3), degree=1)], ids=_id)
def test_fem_dof_numbering_and_constraints_bit_exact(cuda_lib, case):
    spec = fem_problem(**case)
    go, orc = _ops(spec)
    for cell in range(spec.ncells):
        assert np.array_equal(go.cell_dof_indices(cell), orc.cell_dof_indices(cell))
    assert np.array_equal(go.constrained_dofs(), orc.constrained_dofs())


def test_fem_processor_sides_are_constrained(cuda_lib):
    spec = fem_problem((6, 5, 4), degree=2, a="scalar")
    spec = spec.replace(side_kind=[[abi.SIDE_DOMAIN, abi.SIDE_PROCESSOR], [abi.SIDE_PROCESSOR, abi.SIDE_DOMAIN],
                                   [abi.SIDE_DOMAIN, abi.SIDE_DOMAIN]])
    go, orc = _ops(spec)
    x = mt_vector(spec.num_dofs)
    assert np.array_equal(go.constrained_dofs(), orc.constrained_dofs())
    assert rel_err(go.residual(x, np.zeros_like(x)), orc.residual(x)) < TOL


def test_fem_device_tensors_and_affinity_at_size(cuda_lib):
    """cfg1-sized (Q1 2D 256^2) and a larger Q2 3D case on device tensors: residual is affine,
    R(a u + (1-a) v) = a R(u) + (1-a) R(v)."""
    import torch
    from pdelab_b200.capi import GridOperator
    for cells, k in (((256, 256), 1), ((48, 40, 32), 2)):
        spec = fem_problem(cells, degree=k, a="scalar", with_c=True)
        go = GridOperator(spec)
        n = spec.num_dofs
        g = torch.Generator(device="cuda").manual_seed(1)
        u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        v = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        Ru, Rv, Rw = (torch.zeros_like(u) for _ in range(3))
        go.residual(u, Ru), go.residual(v, Rv), go.residual(0.3 * u + 0.7 * v, Rw)
        go.synchronize()
        err = (Rw - (0.3 * Ru + 0.7 * Rv)).abs().max() / Rw.abs().max()
        assert err.item() < TOL

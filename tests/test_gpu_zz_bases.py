"""QkDG in the Legendre and Gauss-Lobatto bases (finiteelementmap/qkdg.hh:15, finiteelement/qkdglegendre.hh,
qkdglobatto.hh) on the CUDA path: the table-driven reference-order kernels against the oracle (-m gpu, through the C ABI).
Runs last in the suite (file name) — the newest path."""
import numpy as np
import pytest

from pdelab_b200 import abi
from problems import dg_problem, mt_vector, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    dict(cells=(5, 4), degree=2, a="full", with_b=True, with_c=True, basis=abi.BASIS_LEGENDRE),
    dict(cells=(4, 3, 2), degree=2, a="scalar", bc="mixed", basis=abi.BASIS_LEGENDRE),
    dict(cells=(3, 3), degree=4, a="diagonal", with_c=True, basis=abi.BASIS_LEGENDRE),
    dict(cells=(4, 3), degree=3, a="full", with_b=True, basis=abi.BASIS_LOBATTO),
    dict(cells=(2, 2, 2), degree=4, a="scalar", basis=abi.BASIS_LOBATTO),
]


def check_case(case):
    """Shared with tools/check_bases_gpu.py (a torch-free runner of the same checks)."""
    from oracle import Oracle
    from pdelab_b200.capi import GridOperator
    bc = case.get("bc", "dirichlet_g")
    spec = dg_problem(with_f=True, **dict(case, bc=bc))
    go, orc = GridOperator(spec), Oracle(spec)
    n = spec.num_dofs
    z = mt_vector(n) - 0.5
    errs = {}
    errs["jacobian_apply"] = rel_err(go.jacobian_apply(z, np.zeros(n)), orc.jacobian_apply(z))
    assert go.last_kernel() == "dg_generic_jacobian_apply", go.last_kernel()
    errs["residual"] = rel_err(go.residual(z, np.zeros(n)), orc.residual(z))
    rowptr, colidx = go.fill_pattern()
    orp, oci, ov = orc.jacobian(z)
    assert np.array_equal(rowptr, orp) and np.array_equal(colidx, oci)
    errs["jacobian"] = rel_err(go.jacobian(z, np.zeros(colidx.size)), ov)
    return errs


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_other_bases_match_oracle(cuda_lib, case):
    errs = check_case(case)
    assert max(errs.values()) < TOL, errs


def test_fast_kernels_are_lagrange_only(cuda_lib):
    from pdelab_b200.capi import GridOperator, PDELabError
    spec = dg_problem((8, 4, 4), degree=2, a="scalar", basis=abi.BASIS_LEGENDRE, kernel=abi.KERNEL_FAST)
    go = GridOperator(spec)
    with pytest.raises(PDELabError, match="no fast kernel"):
        go.apply(np.zeros(spec.num_dofs), np.zeros(spec.num_dofs))
    with pytest.raises(PDELabError, match="Lagrange"):
        GridOperator(abi.ProblemSpec((4, 4), space=abi.SPACE_QK, degree=1, basis=abi.BASIS_LEGENDRE))

"""Coefficients sampled per quadrature point (pdb200_problem::pointwise, layout (2) of include/pdelab_b200.h): a
rotating velocity b(x) = (-y, x, ..), c(x) varying inside the cells, A(x) with permeabilityIsConstantPerCell() == false
and, for QkDG, a boundary type that changes inside boundary faces — evaluated exactly where the reference evaluates the
call-backs (convectiondiffusiondg.hh:143-146,178,181,367-382,426,752-763,797; convectiondiffusionfem.hh:97-100,127-129,
254).  CUDA path against the oracle (which test_oracle_vs_numpy.py checks against the analytic call-backs)."""
import numpy as np
import pytest

from pdelab_b200 import abi
from problems import dg_problem, fem_problem, mt_vector, pointwise_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
_id = lambda c: "-".join(f"{k}={v}" for k, v in c.items())

DG_CASES = [
    dict(cells=(5, 4), degree=1, a="scalar"),
    dict(cells=(4, 3), degree=2, a="full", which=("A", "b", "c")),
    dict(cells=(3, 3), degree=3, a="diagonal", weights=abi.DG_WEIGHTS_OFF, method=abi.DG_NIPG),
    dict(cells=(3, 2, 2), degree=1, a="full"),
    dict(cells=(4, 3, 2), degree=2, a="scalar", which=("b", "c")),          # cfg2's space with convection only
    dict(cells=(3, 2, 2), degree=2, a="diagonal", which=("A", "bctype")),
    dict(cells=(4, 2, 3), degree=2, a="full", extent=(1.0, 0.7, 1.3)),
    dict(cells=(2, 1, 2), degree=3, a="scalar"),
    dict(cells=(2, 2, 1), degree=4, a="scalar", which=("A", "b")),
]
FEM_CASES = [
    dict(cells=(6, 5), degree=1, a="scalar"),
    dict(cells=(4, 3), degree=2, a="full"),
    dict(cells=(3, 3, 2), degree=1, a="diagonal", bc="mixed", with_b=True),
    dict(cells=(3, 2, 2), degree=2, a="scalar", which=("b", "c")),
    dict(cells=(3, 2, 2), degree=2, a="full", bc="mixed", with_b=True),
]


def _ops(spec):
    from oracle import Oracle
    from pdelab_b200.capi import GridOperator
    return GridOperator(spec), Oracle(spec)


def _check(spec, dg):
    go, orc = _ops(spec)
    n = spec.num_dofs
    z, y0 = mt_vector(n), mt_vector(n, seed=7)
    assert rel_err(go.jacobian_apply(z, y0.copy()), orc.jacobian_apply(z, y0.copy())) < TOL
    assert go.last_kernel() in ("dg_generic_jacobian_apply", "fem_jacobian_apply")   # no Kronecker form for these
    assert rel_err(go.residual(z, y0.copy()), orc.residual(z, y0.copy())) < TOL
    assert rel_err(go.apply(z, np.full(n, np.nan)), orc.jacobian_apply(z)) < TOL
    rp_o, ci_o, va_o = orc.jacobian()
    rp, ci = go.fill_pattern()
    assert np.array_equal(rp, rp_o) and np.array_equal(ci, ci_o)
    va = go.jacobian(z, np.zeros(ci.size))
    assert rel_err(va, va_o) < TOL
    if dg:
        nbr, nblocks = go.pattern_size(block=True)
        nl = spec.local_size
        vb = go.jacobian(z, np.zeros(nblocks * nl * nl), layout=abi.LAYOUT_BCSR)
        assert abs(np.abs(vb).sum() - np.abs(va_o).sum()) <= 1e-10 * np.abs(va_o).sum()


@pytest.mark.parametrize("case", DG_CASES, ids=_id)
def test_dg_pointwise_coefficients_match_oracle(cuda_lib, case):
    kw = dict(case)
    which = kw.pop("which", ("A", "b", "c", "bctype"))
    _check(pointwise_problem(dg_problem(with_f=True, bc="dirichlet_g", **kw), which), True)


@pytest.mark.parametrize("case", FEM_CASES, ids=_id)
def test_fem_pointwise_coefficients_match_oracle(cuda_lib, case):
    kw = dict(case)
    which = kw.pop("which", ("A", "b", "c"))
    _check(pointwise_problem(fem_problem(**kw), which), False)


def test_pointwise_layout_filled_with_cellwise_fields_equals_the_cellwise_layout(cuda_lib):
    base = dg_problem((4, 3, 2), degree=2, a="full", with_b=True, with_c=True, with_f=True, bc="mixed",
                      kernel=abi.KERNEL_GENERIC)
    NP, nq, nfq = base.points_per_cell, base.nq, base.nfq
    A, b, c, bct = (base.arrays[k] for k in ("A", "b", "c", "bctype"))
    pw = base.replace(A=np.repeat(A[:, None], NP, axis=1), b=np.repeat(b[:, None], NP, axis=1),
                      c=np.repeat(c[:, None], nq, axis=1), bctype=np.repeat(bct[:, None], nfq, axis=1),
                      pointwise=abi.POINTWISE_A | abi.POINTWISE_B | abi.POINTWISE_C | abi.POINTWISE_BCTYPE)
    from pdelab_b200.capi import GridOperator
    g0, g1 = GridOperator(base), GridOperator(pw)
    z = mt_vector(base.num_dofs)
    assert np.array_equal(g0.residual(z, np.zeros_like(z)), g1.residual(z, np.zeros_like(z)))
    nnz = g0.pattern_size()[1]
    assert np.array_equal(g0.jacobian(z, np.zeros(nnz)), g1.jacobian(z, np.zeros(nnz)))


def test_pointwise_errors(cuda_lib):
    from pdelab_b200.capi import GridOperator, PDELabError
    spec = fem_problem((3, 3), degree=1, bc="mixed", with_b=True)
    bct = np.repeat(spec.arrays["bctype"][:, None], spec.nfq, axis=1)
    with pytest.raises(PDELabError, match="face centre"):     # convectiondiffusionfem.hh:226-229
        GridOperator(spec.replace(bctype=bct, pointwise=abi.POINTWISE_BCTYPE))
    with pytest.raises(PDELabError, match="unknown bits"):
        GridOperator(spec.replace(pointwise=64))
    pw = pointwise_problem(dg_problem((4, 2, 2), degree=2, a="scalar"), ("b",))
    with pytest.raises(PDELabError, match="no fast kernel"):
        GridOperator(pw.replace(kernel=abi.KERNEL_FAST)).apply(np.zeros(pw.num_dofs), np.zeros(pw.num_dofs))


@pytest.mark.parametrize("kernel", [abi.KERNEL_FAST, abi.KERNEL_GENERIC])
@pytest.mark.parametrize("cells,degree", [((8, 4, 4), 2), ((6, 5), 1), ((5, 4), 2), ((4, 3, 2), 1), ((4, 2, 2), 3)])
def test_weights_off_keeps_the_penalty_where_A_vanishes(cuda_lib, cells, degree, kernel):
    """weightsOff: the penalty alpha/h_F k(k+d-1) does not contain A (harmonic_average = 1,
    convectiondiffusiondg.hh:334-338, 724-727), so cells with A == 0 keep it on interior and Dirichlet faces."""
    spec = dg_problem(cells, degree=degree, a="scalar", weights=abi.DG_WEIGHTS_OFF, alpha=2.0, bc="mixed", kernel=kernel)
    A = spec.arrays["A"].copy()
    A[::3] = 0.0
    spec = spec.replace(A=A)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    assert rel_err(go.apply(z, np.zeros_like(z)), orc.jacobian_apply(z)) < TOL
    if kernel == abi.KERNEL_FAST:
        assert go.last_kernel() in ("dg_fast_q2_3d", "dg_small", "dg_kron_3d")

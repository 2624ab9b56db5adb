"""The C++ oracle (oracle/pdelab_oracle.cc) against an independent dense numpy assembly of the
same weak forms (tests/numpy_assembly.py) and against the Kronecker derivation
(tests/kron_reference.py).  CPU only."""
import numpy as np
import pytest

from numpy_assembly import Grid, apply_constraints, assemble, dof_map
from oracle import Oracle
from pdelab_b200 import abi
from problems import dg_problem, fem_problem, mt_vector, rel_err
import kron_reference

TOL = 1e-12

DG_CASES = [
    dict(cells=(3, 2), degree=1, a="scalar"),
    dict(cells=(3, 2), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(2, 2), degree=3, a="diagonal", with_b=True),
    dict(cells=(3, 2, 2), degree=1, a="full", with_b=True, with_c=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(2, 2, 2), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(3, 2, 2), degree=2, a="scalar", bc="mixed", with_b=True),
    dict(cells=(3, 2, 1), degree=2, a="diagonal", method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF),
    dict(cells=(2, 2, 1), degree=2, a="scalar", method=abi.DG_IIPG, intorderadd=1),
    dict(cells=(1, 1, 1), degree=2, a="identity"),
]
FEM_CASES = [
    dict(cells=(4, 3), degree=1, a="scalar"),
    dict(cells=(3, 3), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(3, 2, 2), degree=1, a="full", with_b=True, with_c=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(2, 2, 2), degree=2, a="diagonal", with_c=True),
    dict(cells=(3, 2, 2), degree=2, a="scalar", bc="mixed", with_b=True),
    dict(cells=(4, 3), degree=2, a="scalar", bc="mixed", with_b=True),
    dict(cells=(1, 1, 1), degree=2, a="identity"),
]
_id = lambda c: "-".join(f"{k}={v}" for k, v in c.items())


def _csr_to_dense(rowptr, colidx, values, n):
    J = np.zeros((n, n))
    for r in range(n):
        s, e = int(rowptr[r]), int(rowptr[r + 1])
        J[r, colidx[s:e].astype(np.int64)] = values[s:e]
    return J


def _check(spec):
    orc = Oracle(spec)
    J, r0, con = assemble(spec)
    n = J.shape[0]
    assert orc.num_dofs == n
    # DOF numbering (bit-exact) and constraint set
    dmap, _ = dof_map(Grid(spec))
    for cell in range(spec.ncells):
        assert np.array_equal(orc.cell_dof_indices(cell).astype(np.int64), dmap[cell])
    assert np.array_equal(orc.constrained_dofs().astype(np.int64), np.flatnonzero(con))
    z = mt_vector(n)
    scale = np.abs(J @ z).max()
    # jacobian_apply: y += J z, constrained rows zero
    want = J @ z
    want[con] = 0.0
    y0 = mt_vector(n, seed=3)
    y0c = y0.copy()
    y0c[con] = 0.0          # constrain_residual zeroes the whole entry, including what was in y
    got = orc.jacobian_apply(z, y0.copy())
    assert np.abs(got - (want + y0c)).max() / scale < TOL
    # residual: affine
    want_r = J @ z + r0
    want_r[con] = 0.0
    got_r = orc.residual(z)
    assert np.abs(got_r - want_r).max() / max(scale, np.abs(want_r).max()) < TOL
    # assembled matrix: pattern covers J, values equal, constrained rows are unit rows
    rowptr, colidx, values = orc.jacobian()
    Jd = _csr_to_dense(rowptr, colidx, values, n)
    assert np.abs(Jd - apply_constraints(J, r0, con)).max() / np.abs(J).max() < TOL
    for r in range(n):
        cols = colidx[int(rowptr[r]):int(rowptr[r + 1])]
        assert np.all(np.diff(cols.astype(np.int64)) > 0)      # ascending, no duplicates
    return orc, J


@pytest.mark.parametrize("case", DG_CASES, ids=_id)
def test_dg_oracle_matches_numpy_assembly(case):
    kw = dict(case)
    if kw.get("bc", "dirichlet") == "dirichlet":
        kw["bc"] = "dirichlet_g"
    spec = dg_problem(with_f=True, **kw)
    orc, J = _check(spec)
    # the pattern is exactly: own block + face-neighbour blocks
    rowptr, colidx = orc.pattern()
    n = spec.local_size
    N = spec.cells
    nnz = 0
    for e in range(spec.ncells):
        c = np.unravel_index(e, N[::-1])[::-1]
        nb = 1 + sum(int(c[d] > 0) + int(c[d] < N[d] - 1) for d in range(spec.dim))
        nnz += nb * n * n
    assert colidx.size == nnz


@pytest.mark.parametrize("case", FEM_CASES, ids=_id)
def test_fem_oracle_matches_numpy_assembly(case):
    spec = fem_problem(**case)
    orc, J = _check(spec)
    # conforming pattern = DOFs sharing a cell
    rowptr, colidx = orc.pattern()
    dmap, n = dof_map(Grid(spec))
    adj = np.zeros((n, n), dtype=bool)
    for ids in dmap:
        adj[np.ix_(ids, ids)] = True
    assert colidx.size == adj.sum()
    for r in range(n):
        assert np.array_equal(colidx[int(rowptr[r]):int(rowptr[r + 1])].astype(np.int64), np.flatnonzero(adj[r]))


def test_fem_finite_difference_apply_is_close_to_exact():
    """The reference's FEM jacobian_apply is a finite-difference mixin (epsilon = 1e-7): it agrees
    with the exact J z only to ~1e-7 (documented deviation)."""
    spec = fem_problem((4, 3), degree=2, a="scalar", with_c=True)
    orc = Oracle(spec)
    z = mt_vector(spec.num_dofs)
    exact = orc.jacobian_apply(z)
    fd = orc.fem_jacobian_apply_fd(z)
    assert 1e-13 < rel_err(fd, exact) < 1e-5


@pytest.mark.parametrize("case", [
    dict(cells=(4, 3, 2), extent=(1.0, 0.7, 1.3), a="diagonal", with_c=True),
    dict(cells=(3, 3, 3), a="scalar", weights=abi.DG_WEIGHTS_OFF, method=abi.DG_NIPG),
    dict(cells=(5, 4), a="scalar", degree=3),
    dict(cells=(2, 2, 2), a="scalar", degree=4),
], ids=_id)
def test_dg_oracle_matches_kronecker_form(case):
    """Second derivation: sum of Kronecker products of exactly integrated 1-D matrices."""
    kw = dict(case)
    degree = kw.pop("degree", 2)
    spec = dg_problem(degree=degree, **kw)
    dim = spec.dim
    A = np.asarray(spec.arrays["A"])
    Ad = np.repeat(A.reshape(-1, 1), dim, axis=1) if spec.a_mode == abi.A_SCALAR else A.reshape(-1, dim)
    z = mt_vector(spec.num_dofs)
    want = kron_reference.dg_apply_kron(spec.cells, [spec.upper[d] - spec.lower[d] for d in range(dim)], degree, Ad, z,
                                        spec.alpha, theta={abi.DG_SIPG: -1.0, abi.DG_NIPG: 1.0, abi.DG_IIPG: 0.0}[spec.method],
                                        weights_on=spec.weights == abi.DG_WEIGHTS_ON, c=spec.arrays["c"])
    # the numpy derivation inverts the 1-D mass matrix in double precision (cond ~ 1e3 at k = 4)
    assert rel_err(Oracle(spec).jacobian_apply(z), want) < (TOL if degree <= 2 else 1e-11)


def test_quadrature_matches_numpy_leggauss():
    for k in (1, 2, 3, 4):
        for add in (0, 1):
            spec = dg_problem((2, 2), degree=k, intorderadd=add)
            x, w = Oracle(spec).quadrature()
            xr, wr = np.polynomial.legendre.leggauss(spec.m)
            assert np.abs(x - 0.5 * (xr + 1)).max() < 1e-15 and np.abs(w - 0.5 * wr).max() < 1e-15


# ---- other QkDG bases (finiteelementmap/qkdg.hh:15: QkDGBasisPolynomial legendre / lobatto) -------------------------
BASIS_CASES = [
    dict(cells=(3, 2), degree=1, a="scalar", basis=abi.BASIS_LEGENDRE),
    dict(cells=(3, 2), degree=2, a="full", with_b=True, with_c=True, basis=abi.BASIS_LEGENDRE),
    dict(cells=(2, 2), degree=3, a="diagonal", with_b=True, basis=abi.BASIS_LEGENDRE),
    dict(cells=(2, 2, 2), degree=2, a="full", with_b=True, with_c=True, basis=abi.BASIS_LEGENDRE),
    dict(cells=(3, 2, 2), degree=2, a="scalar", bc="mixed", with_b=True, basis=abi.BASIS_LEGENDRE),
    dict(cells=(2, 2), degree=4, a="scalar", basis=abi.BASIS_LEGENDRE),
    dict(cells=(3, 2), degree=2, a="full", with_b=True, with_c=True, basis=abi.BASIS_LOBATTO),
    dict(cells=(2, 2), degree=3, a="diagonal", with_b=True, basis=abi.BASIS_LOBATTO),
    dict(cells=(2, 2, 1), degree=3, a="scalar", bc="mixed", basis=abi.BASIS_LOBATTO),
    dict(cells=(2, 2), degree=4, a="scalar", basis=abi.BASIS_LOBATTO),
]


@pytest.mark.parametrize("case", BASIS_CASES, ids=_id)
def test_dg_oracle_matches_numpy_assembly_in_other_bases(case):
    """Legendre (finiteelement/qkdglegendre.hh) and Gauss-Lobatto Lagrange (finiteelement/qkdglobatto.hh) QkDG bases:
    the oracle's recurrences / closed-form nodes against numpy.polynomial (independent) through the dense assembly."""
    _check(dg_problem(with_f=True, **case))


def test_lobatto_equals_lagrange_up_to_degree_two_and_legendre_spans_the_same_space():
    """k <= 2: the Gauss-Lobatto points are the equidistant ones, so the two nodal bases coincide.  Any k: a change of
    basis C (Legendre coefficients -> Lagrange coefficients, cell-wise Kronecker) maps the Legendre operator onto the
    Lagrange one:  J_leg = C^T J_lag C."""
    base = dict(cells=(3, 2), a="diagonal", with_c=True, with_f=True)
    for k in (1, 2):
        lag = Oracle(dg_problem(degree=k, **base))
        lob = Oracle(dg_problem(degree=k, basis=abi.BASIS_LOBATTO, **base))
        z = mt_vector(lag.num_dofs)
        assert rel_err(lob.jacobian_apply(z), lag.jacobian_apply(z)) < 1e-13
        assert rel_err(lob.residual(z), lag.residual(z)) < 1e-13
    from numpy.polynomial import legendre as Lg
    k = 3
    lag = Oracle(dg_problem(degree=k, **base))
    leg = Oracle(dg_problem(degree=k, basis=abi.BASIS_LEGENDRE, **base))
    nodes = np.arange(k + 1) / k
    C1 = np.array([[Lg.legval(2 * x - 1, np.eye(k + 1)[n]) for n in range(k + 1)] for x in nodes])   # nodal values of P_n
    C = np.kron(np.eye(6), np.kron(C1, C1))                                                             # y (x) x per cell
    zl = mt_vector(leg.num_dofs) - 0.5
    assert rel_err(leg.jacobian_apply(zl), C.T @ lag.jacobian_apply(C @ zl)) < 1e-11


# ---- spatially varying coefficients: A, b, c (and the DG boundary type) evaluated per quadrature point -------------
# (convectiondiffusiondg.hh:143-146,178,181,367-382,426,752-763; convectiondiffusionfem.hh:97-100,127-129,254)
from problems import pointwise_problem  # noqa: E402

PW_DG_CASES = [
    dict(cells=(3, 2), degree=1, a="scalar"),
    dict(cells=(3, 3), degree=2, a="full", which=("A", "b", "c")),
    dict(cells=(2, 3), degree=2, a="diagonal", weights=abi.DG_WEIGHTS_OFF, method=abi.DG_NIPG),
    dict(cells=(2, 2, 2), degree=1, a="full"),
    dict(cells=(2, 2, 2), degree=2, a="scalar", which=("b", "c")),
    dict(cells=(3, 2, 2), degree=2, a="diagonal", which=("A", "bctype")),
    dict(cells=(2, 1, 2), degree=3, a="scalar", which=("A", "b", "c", "bctype")),
]
PW_FEM_CASES = [
    dict(cells=(4, 3), degree=1, a="scalar"),
    dict(cells=(3, 3), degree=2, a="full"),
    dict(cells=(2, 2, 2), degree=1, a="diagonal", bc="mixed", with_b=True),
    dict(cells=(2, 2, 2), degree=2, a="scalar", which=("b", "c")),
]


@pytest.mark.parametrize("case", PW_DG_CASES, ids=_id)
def test_dg_oracle_pointwise_coefficients_match_numpy_assembly(case):
    kw = dict(case)
    which = kw.pop("which", ("A", "b", "c", "bctype"))
    spec = pointwise_problem(dg_problem(with_f=True, bc="dirichlet_g", **kw), which)
    _check(spec)


@pytest.mark.parametrize("case", PW_FEM_CASES, ids=_id)
def test_fem_oracle_pointwise_coefficients_match_numpy_assembly(case):
    kw = dict(case)
    which = kw.pop("which", ("A", "b", "c"))
    spec = pointwise_problem(fem_problem(**kw), which)
    _check(spec)


def test_pointwise_layout_with_cellwise_constant_fields_equals_the_cellwise_layout():
    """Filling the point-wise arrays with a cell-wise constant field must reproduce the cell-wise path bit for bit."""
    base = dg_problem((3, 2, 2), degree=2, a="full", with_b=True, with_c=True, with_f=True, bc="mixed")
    NP, nq, nfq = base.points_per_cell, base.nq, base.nfq
    A, b, c, bct = (base.arrays[k] for k in ("A", "b", "c", "bctype"))
    pw = base.replace(A=np.repeat(A[:, None], NP, axis=1), b=np.repeat(b[:, None], NP, axis=1),
                      c=np.repeat(c[:, None], nq, axis=1), bctype=np.repeat(bct[:, None], nfq, axis=1),
                      pointwise=abi.POINTWISE_A | abi.POINTWISE_B | abi.POINTWISE_C | abi.POINTWISE_BCTYPE)
    z = mt_vector(base.num_dofs)
    o0, o1 = Oracle(base), Oracle(pw)
    assert np.array_equal(o0.residual(z), o1.residual(z))
    assert np.array_equal(o0.jacobian_apply(z), o1.jacobian_apply(z))
    assert np.array_equal(o0.jacobian()[2], o1.jacobian()[2])

"""The reference's own integration tests for this path, restated against the CPU oracle.

dune-pdelab holds no golden vectors for residual / jacobian_apply / jacobian; what
its tests pin are error thresholds of manufactured problems and structural equalities.  Each test
below names the reference test it restates and uses its grid, space, parameters and threshold.
tests/test_gpu_reference_invariants.py runs the same cases through the CUDA path."""
import numpy as np
import pytest

from manufactured import OracleOps, bicgstab, l2_error_squared, node_coordinates, sample_data, solve_stationary
from oracle import Oracle
from pdelab_b200 import abi
from pdelab_b200.abi import ProblemSpec


def u_centre(X):   # test/testconvectiondiffusiondg.cc:33-41, test/testmatrixfree.cc:33-40
    return np.exp(-np.sum((X - 0.5) ** 2, axis=1))


def f_centre(X):   # test/testconvectiondiffusiondg.cc:14-24: 4 (1 - c) g in 2D
    c = np.sum((X - 0.5) ** 2, axis=1)
    return (2.0 * X.shape[1] - 4.0 * c) * np.exp(-c)


def u_origin(X):   # test/testfastdgassembler.cc:82-88
    return np.exp(-np.sum(X ** 2, axis=1))


def f_origin(X):   # test/testfastdgassembler.cc:66-72
    n2 = np.sum(X ** 2, axis=1)
    return (2.0 * X.shape[1] - 4.0 * n2) * np.exp(-n2)


CASES = {
    # name: (spec kwargs, u, f, threshold on the squared L2 error)
    "testconvectiondiffusiondg": (dict(cells=(16, 16), space=abi.SPACE_QKDG, degree=1, method=abi.DG_SIPG,
                                       weights=abi.DG_WEIGHTS_ON, alpha=1.0), u_centre, f_centre, 1e-6),
    "testfastdgassembler": (dict(cells=(32, 32), space=abi.SPACE_QKDG, degree=1, method=abi.DG_SIPG,
                                 weights=abi.DG_WEIGHTS_ON, alpha=2.0), u_origin, f_origin, 1e-8),
    "matrix_free_linear": (dict(cells=(16, 16), space=abi.SPACE_QKDG, degree=1, method=abi.DG_SIPG,
                                weights=abi.DG_WEIGHTS_ON, alpha=3.0), u_centre, f_centre, 1e-6),
    "testmatrixfree": (dict(cells=(32, 32), space=abi.SPACE_QK, degree=2), u_centre, f_centre, 1e-7),
}


def make_case(name):
    kw, u, f, thr = CASES[name]
    spec = sample_data(ProblemSpec(**kw), u, f)
    x0 = np.zeros(spec.num_dofs)
    if spec.space == abi.SPACE_QK:
        # interpolate(g, gfs, x): Lagrange interpolation of the Dirichlet extension
        x0 = u(node_coordinates(spec))
    return spec, x0, u, thr


@pytest.mark.parametrize("name", ["testconvectiondiffusiondg", "testfastdgassembler", "matrix_free_linear"])
def test_dg_manufactured_solution_matrix_based(name):
    spec, x0, u, thr = make_case(name)
    x, _ = solve_stationary(OracleOps(spec), x0)
    err = l2_error_squared(spec, x, u)
    assert np.isfinite(err) and err <= thr, err


def test_testmatrixfree_q2_matrix_based_and_matrix_free():
    """test/testmatrixfree.cc:122-178: both solves must reach err^2 <= 1e-7."""
    spec, x0, u, thr = make_case("testmatrixfree")
    ops = OracleOps(spec)
    x, _ = solve_stationary(ops, x0)
    assert l2_error_squared(spec, x, u) <= thr
    xmf, its = solve_stationary(ops, x0, matrix_free=True, reduction=1e-10)
    assert l2_error_squared(spec, xmf, u) <= thr
    assert its > 1


def test_matrix_free_linear_same_iteration_count():
    """test/matrixfree/matrix_free_linear.cc:390-393: the Krylov solver takes the same number of
    iterations with the assembled operator and with the matrix-free one, and err^2 <= 1e-6."""
    spec, x0, u, thr = make_case("matrix_free_linear")
    ops = OracleOps(spec)
    r = ops.residual(x0)
    J = ops.matrix()
    z_mb, it_mb = bicgstab(lambda v: J @ v, r, 1e-10)
    z_mf, it_mf = bicgstab(ops.jacobian_apply, r, 1e-10)
    assert it_mb == it_mf
    assert l2_error_squared(spec, x0 - z_mf, u) <= thr


def test_blocked_and_flat_dg_orderings_coincide():
    """test/test-blocked-istl-ordering.cc:65-72,82-121: QkDG k=2 on 4^3 cells — the flat container
    index of local DOF i of cell e is e*n + i, i.e. flat and Blocking::fixed vectors share memory."""
    spec = ProblemSpec((4, 4, 4), space=abi.SPACE_QKDG, degree=2)
    orc = Oracle(spec)
    n = spec.local_size
    for e in range(spec.ncells):
        assert np.array_equal(orc.cell_dof_indices(e), np.arange(e * n, (e + 1) * n, dtype=np.uint64))


def test_sipg_jacobian_is_symmetric_and_annihilates_constants_in_the_interior():
    from problems import dg_problem
    spec = dg_problem((4, 4, 3), degree=2, a="scalar")
    ops = OracleOps(spec)
    J = ops.matrix().toarray()
    assert np.abs(J - J.T).max() / np.abs(J).max() < 1e-13
    y = J @ np.ones(ops.n)
    n = spec.local_size
    interior = [e for e in range(spec.ncells)
                if all(0 < c < N - 1 for c, N in zip(np.unravel_index(e, spec.cells[::-1])[::-1], spec.cells))]
    assert interior
    for e in interior:
        assert np.abs(y[e * n:(e + 1) * n]).max() / np.abs(J).max() < 1e-13

"""Device-resident Krylov solvers (csrc/krylov.cu) against the host restatement of dune-istl's
iterations and the reference's solver-level acceptance tests:
test/matrixfree/matrix_free_linear.cc:390-393 (equal iteration counts assembled vs matrix-free,
err^2 <= 1e-6), test/testmatrixfree.cc:175-178 (Q2, err^2 <= 1e-7)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from manufactured import GpuOps, OracleOps, bicgstab, l2_error_squared
from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator
from problems import dg_problem, fem_problem, mt_vector, rel_err
from test_reference_invariants import make_case

pytestmark = pytest.mark.gpu


def host_cg(apply, b, reduction, dinv=None, maxit=5000):
    """dune-istl CGSolver::apply with SeqJac (w = 1) or Richardson(1.0), x0 = 0."""
    x = np.zeros_like(b)
    r = b.copy()
    def0 = np.linalg.norm(r)
    p = r * dinv if dinv is not None else r.copy()
    rholast = p @ r
    for i in range(1, maxit + 1):
        q = apply(p)
        lam = rholast / (p @ q)
        x += lam * p
        r -= lam * q
        if np.linalg.norm(r) < reduction * def0:
            return x, i
        z = r * dinv if dinv is not None else r
        rho = z @ r
        p = z + (rho / rholast) * p
        rholast = rho
    raise RuntimeError("CG did not converge")


def test_matrix_free_bicgstab_matches_host_restatement_and_reference_test(cuda_lib):
    """ISTLBackend_SEQ_MatrixFree_BCGS_Richardson on matrix_free_linear.cc's problem."""
    spec, x0, u, thr = make_case("matrix_free_linear")
    go = GridOperator(spec)
    r = go.residual(x0, np.zeros(spec.num_dofs))
    z_host, it_host = bicgstab(OracleOps(spec).jacobian_apply, r, 1e-10)
    z = np.zeros(spec.num_dofs)
    rr = r.copy()
    res = go.solve(z, rr, 1e-10)                      # matrix-free, on the device
    assert res["converged"] == 1
    assert abs(res["iterations"] - it_host) <= 1, (res, it_host)
    assert rel_err(z, z_host) < 1e-7
    assert res["reduction"] < 1e-10 and abs(res["first_defect"] - np.linalg.norm(r)) < 1e-12 * np.linalg.norm(r)
    # the defect handed back in r is b - A z
    assert np.linalg.norm(rr) <= 1.0000001 * res["defect"] + 1e-300
    assert l2_error_squared(spec, x0 - z, u) <= thr
    # assembled operator, same solver: same iteration count (matrix_free_linear.cc:390-393)
    import torch
    nr, nnz = go.pattern_size()
    vals = torch.zeros(nnz, dtype=torch.float64, device="cuda")
    go.jacobian(torch.zeros(nr, dtype=torch.float64, device="cuda"), vals, fresh=True)
    z2 = np.zeros(spec.num_dofs)
    res2 = go.solve(z2, r.copy(), 1e-10, values=vals)
    # BiCGSTAB amplifies rounding: the assembled product (row gather) and the Kronecker kernel sum in
    # different orders, which moves the stopping test by a few of ~60 iterations (the CPU suite, where
    # both operators share one arithmetic, asserts equality)
    assert res2["converged"] == 1 and abs(res2["iterations"] - res["iterations"]) <= 5
    assert rel_err(z2, z_host) < 1e-7


@pytest.mark.parametrize("name", ["testconvectiondiffusiondg", "testfastdgassembler", "matrix_free_linear",
                                  "testmatrixfree"])
@pytest.mark.parametrize("matrix_free", [True, False])
def test_solve_stationary_reaches_the_reference_thresholds(cuda_lib, name, matrix_free):
    """StationaryLinearProblemSolver::apply entirely on the device."""
    spec, x0, u, thr = make_case(name)
    go = GridOperator(spec)
    x = x0.copy()
    res = go.solve_stationary(x, reduction=1e-10, matrix_free=matrix_free,
                              precond=abi.PRECOND_NONE if matrix_free else abi.PRECOND_JACOBI)
    assert res["converged"] == 1, res
    err = l2_error_squared(spec, x, u)
    assert np.isfinite(err) and err <= thr, err
    assert res["defect"] <= 1e-10 * res["first_defect"] * 1.0000001


def test_cg_jacobi_on_assembled_q1_poisson_matches_host_cg_and_direct_solve(cuda_lib):
    import torch
    spec = fem_problem((12, 10, 8), degree=1, a="scalar")
    ops = GpuOps(spec)
    J = ops.matrix().tocsr()
    go = ops.go
    n = spec.num_dofs
    b = mt_vector(n, seed=3)
    b[go.constrained_dofs().astype(np.int64)] = 0.0   # consistent right-hand side (constrained rows are unit rows)
    dinv = 1.0 / J.diagonal()
    x_host, it_host = host_cg(lambda v: J @ v, b, 1e-9, dinv)
    vals = torch.from_numpy(np.ascontiguousarray(J.data)).cuda()
    z = np.zeros(n)
    res = go.solve(z, b.copy(), 1e-9, solver=abi.SOLVER_CG, precond=abi.PRECOND_JACOBI, values=vals)
    assert res["converged"] == 1 and abs(res["iterations"] - it_host) <= 1, (res, it_host)
    x_direct = spla.spsolve(J.tocsc(), b)
    assert rel_err(z, x_direct) < 1e-7
    # matrix-free CG (Richardson) on the same operator
    z2 = np.zeros(n)
    res2 = go.solve(z2, b.copy(), 1e-9, solver=abi.SOLVER_CG)
    assert res2["converged"] == 1
    assert rel_err(z2, x_direct) < 1e-6


def test_solver_on_device_tensors_large_dg_and_not_converged_is_reported(cuda_lib):
    """Device pointers in place; maxiter too small -> converged == 0 (no exception), like
    InverseOperatorResult."""
    import torch
    spec = dg_problem((32, 16, 16), degree=2, a="scalar")
    go = GridOperator(spec)
    n = spec.num_dofs
    g = torch.Generator(device="cuda").manual_seed(1)
    b = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    z = torch.zeros_like(b)
    res = go.solve(z, b.clone(), 1e-8, solver=abi.SOLVER_CG)     # SIPG with b = 0 is SPD
    assert res["converged"] == 1 and go.last_kernel() == "dg_fast_q2_3d"
    y = torch.empty_like(b)
    go.apply(z, y)
    assert float((y - b).norm() / b.norm()) < 2e-8
    z.zero_()
    res = go.solve(z, b.clone(), 1e-12, maxiter=3)
    assert res["converged"] == 0 and res["iterations"] == 3
    with pytest.raises(Exception, match="unknown preconditioner"):
        go.solve(z, b.clone(), 1e-8, precond=7)


BJ_CASES = [
    dict(cells=(6, 5), degree=1, a="scalar"), dict(cells=(5, 4), degree=2, a="diagonal", with_c=True),
    dict(cells=(4, 3, 3), degree=1, a="diagonal", bc="mixed"), dict(cells=(4, 4, 3), degree=2, a="scalar", extent=(1.0, 0.7, 1.3)),
    dict(cells=(1, 1, 1), degree=2, a="identity"),
]


@pytest.mark.parametrize("case", BJ_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_block_jacobi_is_the_exact_inverse_of_the_diagonal_blocks(cuda_lib, case):
    """AssembledBlockJacobiPreconditionerLocalOperator (assembledblockjacobipreconditioner.hh:96-230) LU-factorises
    the assembled diagonal blocks; the fast-diagonalisation kernel must give the same D^-1 r."""
    spec = dg_problem(**case)
    go = GridOperator(spec)
    n, ncell, nd = spec.local_size, spec.ncells, spec.num_dofs
    rowptr, colidx = go.fill_pattern(block=True)
    vals = go.jacobian(np.zeros(nd), np.zeros(colidx.size * n * n), layout=abi.LAYOUT_BCSR, fresh=True).reshape(-1, n, n)
    r = mt_vector(nd, seed=5)
    want = np.zeros(nd)
    for e in range(ncell):
        k = [j for j in range(int(rowptr[e]), int(rowptr[e + 1])) if int(colidx[j]) == e][0]
        want[e * n:(e + 1) * n] = np.linalg.solve(vals[k], r[e * n:(e + 1) * n])
    z = go.block_jacobi_apply(r, np.zeros(nd))
    assert rel_err(z, want) < 1e-11


def test_block_jacobi_preconditioned_krylov(cuda_lib):
    """ISTLBackend_SEQ_MatrixFree_Base with the block-Jacobi preconditioner (backends.hh:62-143): same
    solution as the unpreconditioned solve, in far fewer iterations, on a strongly heterogeneous field."""
    import torch
    spec = dg_problem((24, 16, 16), degree=2, a="scalar")
    go = GridOperator(spec)
    nd = spec.num_dofs
    g = torch.Generator(device="cuda").manual_seed(4)
    b = torch.rand(nd, dtype=torch.float64, device="cuda", generator=g)
    z0, z1, z2 = (torch.zeros_like(b) for _ in range(3))
    plain = go.solve(z0, b.clone(), 1e-8, solver=abi.SOLVER_CG)
    bj = go.solve(z1, b.clone(), 1e-8, solver=abi.SOLVER_CG, precond=abi.PRECOND_BLOCK_JACOBI)
    bj2 = go.solve(z2, b.clone(), 1e-8, solver=abi.SOLVER_BICGSTAB, precond=abi.PRECOND_BLOCK_JACOBI)
    assert plain["converged"] == bj["converged"] == bj2["converged"] == 1
    assert bj["iterations"] < 0.6 * plain["iterations"], (plain, bj)
    y = torch.empty_like(b)
    for z in (z1, z2):
        go.apply(z, y)
        assert float((y - b).norm() / b.norm()) < 5e-8
    assert float((z1 - z0).norm() / z0.norm()) < 1e-5
    # the manufactured DG problem of matrix_free_linear.cc with the preconditioned matrix-free back-end
    spec, x0, u, thr = make_case("matrix_free_linear")
    go = GridOperator(spec)
    x = x0.copy()
    res = go.solve_stationary(x, reduction=1e-10, matrix_free=True, precond=abi.PRECOND_BLOCK_JACOBI)
    assert res["converged"] == 1 and l2_error_squared(spec, x, u) <= thr


DIAG_CASES = [
    ("fem", dict(cells=(7, 6), degree=1, a="scalar")), ("fem", dict(cells=(5, 4), degree=2, a="diagonal", with_c=True)),
    ("fem", dict(cells=(5, 4, 3), degree=1, a="diagonal")), ("fem", dict(cells=(4, 3, 3), degree=2, a="scalar", with_c=True)),
    ("fem", dict(cells=(35, 3, 14), degree=2, a="scalar")),
    ("dg", dict(cells=(6, 5), degree=1, a="scalar")), ("dg", dict(cells=(5, 4), degree=2, a="diagonal", with_c=True)),
    ("dg", dict(cells=(4, 3, 3), degree=1, a="diagonal", bc="mixed")), ("dg", dict(cells=(4, 4, 3), degree=2, a="scalar")),
]


@pytest.mark.parametrize("kind,case", DIAG_CASES, ids=lambda c: c if isinstance(c, str) else "-".join(f"{k}={v}" for k, v in c.items()))
def test_matrix_free_point_diagonal_equals_the_assembled_diagonal(cuda_lib, kind, case):
    """PointDiagonalLocalOperatorWrapper (localoperator/pointdiagonalwrapper.hh): diag(J) without a matrix."""
    spec = fem_problem(**case) if kind == "fem" else dg_problem(**case)
    ops = GpuOps(spec)
    J = ops.matrix().tocsr()
    d = ops.go.point_diagonal(np.zeros(spec.num_dofs))
    assert rel_err(d, J.diagonal()) < 1e-12


def test_matrix_free_jacobi_cg_on_the_testmatrixfree_problem(cuda_lib):
    """Q2 conforming problem of test/testmatrixfree.cc solved matrix-free with point-Jacobi CG on the device."""
    spec, x0, u, thr = make_case("testmatrixfree")
    go = GridOperator(spec)
    x_plain, x_jac = x0.copy(), x0.copy()
    plain = go.solve_stationary(x_plain, reduction=1e-10, solver=abi.SOLVER_CG, matrix_free=True)
    jac = go.solve_stationary(x_jac, reduction=1e-10, solver=abi.SOLVER_CG, matrix_free=True, precond=abi.PRECOND_JACOBI)
    assert plain["converged"] == 1 and jac["converged"] == 1
    assert l2_error_squared(spec, x_jac, u) <= thr
    assert rel_err(x_jac, x_plain) < 1e-6
    # heterogeneous coefficient: the diagonal scaling pays
    spec = fem_problem((24, 20, 16), degree=1, a="scalar")
    go = GridOperator(spec)
    b = mt_vector(spec.num_dofs, seed=3)
    b[go.constrained_dofs().astype(np.int64)] = 0.0
    r1 = go.solve(np.zeros_like(b), b.copy(), 1e-8, solver=abi.SOLVER_CG)
    r2 = go.solve(np.zeros_like(b), b.copy(), 1e-8, solver=abi.SOLVER_CG, precond=abi.PRECOND_JACOBI)
    assert r1["converged"] == r2["converged"] == 1 and r2["iterations"] < r1["iterations"]


# ---- block diagonal / off-diagonal wrappers and block SOR (backend/istl/matrixfree/blocksorpreconditioner.hh) -------

def _dense_blocks(go, spec):
    """Assembled Jacobian of the device path as a dense matrix (small problems) + the cell-block size."""
    n, nd = spec.local_size, spec.num_dofs
    rowptr, colidx = go.fill_pattern()
    vals = go.jacobian(np.zeros(nd), np.zeros(colidx.size), fresh=True)
    import scipy.sparse as sp
    return sp.csr_matrix((vals, colidx.astype(np.int64), rowptr.astype(np.int64)), shape=(nd, nd)).toarray(), n


SOR_CASES = [dict(cells=(6, 5), degree=1, a="scalar"), dict(cells=(5, 4), degree=2, a="diagonal", with_c=True),
             dict(cells=(4, 3, 3), degree=1, a="diagonal", bc="mixed"), dict(cells=(4, 3, 2), degree=2, a="scalar"),
             dict(cells=(3, 4, 5), degree=2, a="diagonal", with_c=True, extent=(1.0, 0.7, 1.3))]


@pytest.mark.parametrize("case", SOR_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_block_diagonal_and_offdiagonal_wrappers(cuda_lib, case):
    """BlockDiagonalLocalOperatorWrapper / BlockOffDiagonalLocalOperatorWrapper (localoperator/blockdiagonalwrapper.hh,
    blockoffdiagonalwrapper.hh): D z and (J - D) z, matrix-free, against the blocks of the assembled Jacobian."""
    spec = dg_problem(**case)
    go = GridOperator(spec)
    J, n = _dense_blocks(go, spec)
    nd = spec.num_dofs
    D = np.zeros_like(J)
    for e in range(spec.ncells):
        s = slice(e * n, (e + 1) * n)
        D[s, s] = J[s, s]
    z = mt_vector(nd, seed=8) - 0.5
    assert rel_err(go.block_diagonal_apply(z, np.full(nd, np.nan)), D @ z) < 1e-12
    assert rel_err(go.block_offdiagonal_apply(z, np.full(nd, np.nan)), (J - D) @ z) < 1e-12


def _sor_reference(J, n, d, v, omega, backward=False):
    """The algorithm of blocksorpreconditioner.hh:43-52, sequential, in index-set order."""
    v = v.copy()
    cells = range(J.shape[0] // n)
    for e in (reversed(cells) if backward else cells):
        s = slice(e * n, (e + 1) * n)
        a = d[s] - J[s, :] @ v + J[s, s] @ v[s]
        v[s] = (1.0 - omega) * v[s] + omega * np.linalg.solve(J[s, s], a)
    return v


@pytest.mark.parametrize("case", SOR_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_block_sor_sweep_is_the_sequential_sweep(cuda_lib, case):
    """The hyperplane-wavefront sweep reproduces the reference's sequential block SOR sweep (same order of updates):
    forward from v = 0 (preconditioner use), with relaxation, backward, and continuing from a given iterate."""
    spec = dg_problem(**case)
    go = GridOperator(spec)
    J, n = _dense_blocks(go, spec)
    nd = spec.num_dofs
    d = mt_vector(nd, seed=9) - 0.5
    zero = np.zeros(nd)
    v1 = go.block_sor_apply(d, np.full(nd, np.nan))
    assert rel_err(v1, _sor_reference(J, n, d, zero, 1.0)) < 1e-11
    v2 = go.block_sor_apply(d, np.full(nd, np.nan), omega=1.3)
    assert rel_err(v2, _sor_reference(J, n, d, zero, 1.3)) < 1e-11
    v3 = go.block_sor_apply(d, v2.copy(), omega=0.9, backward=True, keep_iterate=True)
    assert rel_err(v3, _sor_reference(J, n, d, v2, 0.9, backward=True)) < 1e-11
    v4 = go.block_sor_apply(d, v3.copy(), omega=1.0, keep_iterate=True)
    assert rel_err(v4, _sor_reference(J, n, d, v3, 1.0)) < 1e-11


def test_block_sor_preconditioned_krylov(cuda_lib):
    """ISTLBackend_SEQ_MatrixFree_Base with BlockSORPreconditionerLocalOperator (backends.hh:62-143): BiCGSTAB + block
    SOR and CG + the symmetric sweep converge to the solution of the unpreconditioned solve in fewer iterations than
    block Jacobi."""
    import torch
    spec = dg_problem((16, 12, 12), degree=2, a="scalar")
    go = GridOperator(spec)
    nd = spec.num_dofs
    g = torch.Generator(device="cuda").manual_seed(4)
    b = torch.rand(nd, dtype=torch.float64, device="cuda", generator=g)
    z0, z1, z2, z3 = (torch.zeros_like(b) for _ in range(4))
    bj = go.solve(z0, b.clone(), 1e-8, solver=abi.SOLVER_BICGSTAB, precond=abi.PRECOND_BLOCK_JACOBI)
    sor = go.solve(z1, b.clone(), 1e-8, solver=abi.SOLVER_BICGSTAB, precond=abi.PRECOND_BLOCK_SOR)
    ssor = go.solve(z2, b.clone(), 1e-8, solver=abi.SOLVER_CG, precond=abi.PRECOND_BLOCK_SSOR)
    go.set_relaxation(1.2)
    sor12 = go.solve(z3, b.clone(), 1e-8, solver=abi.SOLVER_BICGSTAB, precond=abi.PRECOND_BLOCK_SOR)
    assert bj["converged"] == sor["converged"] == ssor["converged"] == sor12["converged"] == 1
    assert sor["iterations"] < bj["iterations"], (sor, bj)
    y = torch.empty_like(b)
    for z in (z1, z2, z3):
        go.apply(z, y)
        assert float((y - b).norm() / b.norm()) < 5e-8
        assert float((z - z0).norm() / z0.norm()) < 1e-5

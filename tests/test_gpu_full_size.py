"""BASELINE.json configurations at FULL size through size-independent properties (the oracle only
finishes small grids in seconds): the Kronecker kernels against the reference-order (generic) kernel,
which is oracle-checked at small sizes; linearity; SIPG symmetry; constants in the kernel of interior
rows; assembled Jacobian times vector == matrix-free apply; closed-form pattern size and ordering."""
import numpy as np
import pytest
import torch

from pdelab_b200 import abi
from pdelab_b200.capi import GridOperator

pytestmark = pytest.mark.gpu


def _rand(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand(n, dtype=torch.float64, device="cuda", generator=g)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _dg_spec(cells, k, kernel, with_f=False):
    nc = int(np.prod(cells))
    kappa = 10.0 ** (2.0 * _rand(nc, 42) - 1.0)
    kw = dict(f=_rand(nc * (k + 1) ** 3, 1)) if with_f else {}
    return abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=k, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa, kernel=kernel, **kw)


@pytest.mark.parametrize("cells,k,fast_name", [((128, 128, 128), 2, "dg_fast_q2_3d"), ((64, 64, 64), 4, "dg_kron_3d")])
def test_dg_fast_kernels_equal_the_reference_order_kernel_at_full_size(cuda_lib, cells, k, fast_name):
    """configs[1] (k=2, 128^3, 56.6 M DOFs) and configs[2] (k=4, 64^3, 32.8 M DOFs)."""
    fast = GridOperator(_dg_spec(cells, k, abi.KERNEL_FAST, with_f=True))
    ref = GridOperator(_dg_spec(cells, k, abi.KERNEL_GENERIC, with_f=True))
    n = fast.spec.num_dofs
    x, y = _rand(n, 2), _rand(n, 3)
    jf, jr = torch.empty_like(x), torch.empty_like(x)
    fast.apply(x, jf)
    ref.apply(x, jr)
    assert fast.last_kernel() == fast_name and ref.last_kernel() == "dg_generic_jacobian_apply"
    assert _rel(jf, jr) < 1e-12
    # residual (accumulate form, source term): r0 + J x + R(0)
    r0 = _rand(n, 4)
    rf, rr = r0.clone(), r0.clone()
    fast.residual(x, rf)
    ref.residual(x, rr)
    assert fast.last_kernel() == fast_name + "+r0"
    assert _rel(rf, rr) < 1e-12
    # linearity and SIPG symmetry (b = 0): x^T J y == y^T J x
    jy, jc = torch.empty_like(x), torch.empty_like(x)
    fast.apply(y, jy)
    fast.apply(2.0 * x - 3.0 * y, jc)
    assert _rel(jc, 2.0 * jf - 3.0 * jy) < 1e-12
    sxy, syx = float(torch.dot(x, jy)), float(torch.dot(y, jf))
    assert abs(sxy - syx) <= 1e-12 * abs(sxy)
    # constants are in the kernel of every row whose cell and face neighbours are off the Dirichlet boundary
    ones = torch.ones_like(x)
    fast.apply(ones, jc)
    nl = (k + 1) ** 3
    j1 = jc.view(cells[2], cells[1], cells[0], nl)[1:-1, 1:-1, 1:-1]
    assert float(j1.abs().max()) < 1e-11 * float(jf.abs().max())


def test_cfg4_q2_160_assembled_jacobian_is_consistent_with_the_matrix_free_operator(cuda_lib):
    """configs[3]: Q2 on 160^3 cells — 33,076,161 rows, 1281^3 non-zeros (16.8 GB of values)."""
    C = 160
    nc = C ** 3
    kappa = 10.0 ** (2.0 * _rand(nc, 42) - 1.0)
    go = GridOperator(abi.ProblemSpec((C, C, C), space=abi.SPACE_QK, degree=2, a_mode=abi.A_SCALAR, A=kappa))
    n = go.spec.num_dofs
    nr, nnz = go.pattern_size()
    assert nr == n == 321 ** 3 and nnz == 1281 ** 3
    rowptr = torch.empty(nr + 1, dtype=torch.int64, device="cuda")
    colidx = torch.empty(nnz, dtype=torch.int32, device="cuda")
    go.fill_pattern(rowptr=rowptr, colidx=colidx, index32=True)
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == nnz
    # columns ascending inside every row (dune-istl setIndices): checked on a window of 2e8 entries
    lo, hi = 7 * 10 ** 8, 9 * 10 ** 8
    d = (colidx[lo + 1:hi].to(torch.int64) - colidx[lo:hi - 1].to(torch.int64))
    starts = torch.zeros(hi - lo - 1, dtype=torch.bool, device="cuda")
    rs = rowptr[(rowptr > lo) & (rowptr < hi)] - lo - 1
    starts[rs] = True
    assert bool(((d > 0) | starts).all())
    del d, starts, rs, colidx
    x = _rand(n, 2)
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    go.jacobian(x, vals, fresh=True)
    y_mf, y_mv = torch.empty_like(x), torch.empty_like(x)
    go.apply(x, y_mf)
    assert go.last_kernel() == "fem_kron"
    go.csr_mv(vals, x, y_mv)
    # constrained rows: unit rows in the matrix (y = x there), zero rows in the matrix-free operator
    con = torch.from_numpy(go.constrained_dofs().astype(np.int64)).cuda()
    assert con.numel() == 321 ** 3 - 319 ** 3
    assert torch.equal(y_mv[con], x[con]) and float(y_mf[con].abs().max()) == 0.0
    y_mv[con] = 0.0
    assert _rel(y_mv, y_mf) < 1e-12
    # constants are in the kernel of the interior rows: row sums vanish
    ones = torch.ones_like(x)
    go.csr_mv(vals, ones, y_mv)
    y_mv[con] = 0.0
    assert float(y_mv.abs().max()) < 1e-10 * float(vals.abs().max())

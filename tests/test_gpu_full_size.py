"""BASELINE.json configurations at FULL size.

Direct parity: every configuration is compared with the CPU oracle itself at its full size (the multi-threaded
oracle needs seconds for the vector entry points; pattern + Jacobian of cfg4 are compared on 1e5 sampled rows through
oracle_jacobian_rows, which is bit-identical to the full oracle assembly) — no GPU-vs-GPU link in the chain.
Size-independent properties on top: linearity; SIPG symmetry; constants in the kernel of interior rows; assembled
Jacobian times vector == matrix-free apply; closed-form pattern size and ordering."""
import os
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


# ---- direct parity with the oracle at the BASELINE sizes ------------------------------------------------------------
def _threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _np_spec(spec):
    """the same problem with host (numpy) coefficient arrays, for the oracle"""
    return spec.replace(**{k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in spec.arrays.items() if v is not None})


def _rel_np(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_cfg2_dg_k2_128_jacobian_apply_equals_the_oracle(cuda_lib):
    """configs[1], the headline: DG k=2 on 128^3 cells, y = J z against oracle_jacobian_apply_mt (all host threads)."""
    from oracle import Oracle
    spec = _dg_spec((128, 128, 128), 2, abi.KERNEL_AUTO)
    go = GridOperator(spec)
    z = _rand(spec.num_dofs, 2)
    y = torch.empty_like(z)
    go.apply(z, y)
    assert go.last_kernel() == "dg_fast_q2_3d"
    want = Oracle(_np_spec(spec)).jacobian_apply(z.cpu().numpy(), threads=_threads())
    assert _rel_np(y.cpu().numpy(), want) < 1e-12


def test_cfg2_with_convection_and_reaction_equals_the_oracle(cuda_lib):
    """SURVEY 8d's second cfg2 variant: constant b = (1, 0.5, 0.25), c = 1 — through the Kronecker kernel at 128^3."""
    from oracle import Oracle
    spec = _dg_spec((128, 128, 128), 2, abi.KERNEL_AUTO)
    nc = spec.ncells
    b = torch.tensor([1.0, 0.5, 0.25], dtype=torch.float64, device="cuda").repeat(nc, 1).contiguous()
    c = torch.ones(nc, dtype=torch.float64, device="cuda")
    spec = spec.replace(b=b, c=c)
    go = GridOperator(spec)
    z = _rand(spec.num_dofs, 2)
    y = torch.empty_like(z)
    go.apply(z, y)
    assert go.last_kernel() == "dg_fast_q2_3d"
    want = Oracle(_np_spec(spec)).jacobian_apply(z.cpu().numpy(), threads=_threads())
    assert _rel_np(y.cpu().numpy(), want) < 1e-12


def test_cfg3_dg_k4_64_residual_equals_the_oracle(cuda_lib):
    """configs[2]: DG k=4 on 64^3 cells, r += R(x) with a source term."""
    from oracle import Oracle
    spec = _dg_spec((64, 64, 64), 4, abi.KERNEL_AUTO, with_f=True)
    go = GridOperator(spec)
    x, r = _rand(spec.num_dofs, 2), _rand(spec.num_dofs, 4)
    r0 = r.cpu().numpy().copy()
    go.residual(x, r)
    assert go.last_kernel() == "dg_kron_3d+r0"
    want = Oracle(_np_spec(spec)).residual(x.cpu().numpy(), r0, threads=_threads())
    assert _rel_np(r.cpu().numpy(), want) < 1e-12


def test_cfg4_q2_160_vectors_and_sampled_matrix_rows_equal_the_oracle(cuda_lib):
    """configs[3]: Q2 on 160^3 cells.  residual and jacobian_apply against the oracle at full size; pattern (bit-exact)
    and Jacobian (1e-12) on 1e5 rows: random interior rows of every sub-entity group plus rows on boundary faces, edges
    and corners (constrained unit rows) and their first interior neighbours."""
    from oracle import Oracle
    C = 160
    nc = C ** 3
    kappa = 10.0 ** (2.0 * _rand(nc, 42) - 1.0)
    f = _rand(nc * 27, 1)
    spec = abi.ProblemSpec((C, C, C), space=abi.SPACE_QK, degree=2, a_mode=abi.A_SCALAR, A=kappa, f=f)
    go = GridOperator(spec)
    orc = Oracle(_np_spec(spec))
    n = spec.num_dofs
    x, r = _rand(n, 2), _rand(n, 4)
    y = torch.empty_like(x)
    go.apply(x, y)
    assert go.last_kernel() == "fem_kron"
    assert _rel_np(y.cpu().numpy(), orc.jacobian_apply(x.cpu().numpy(), threads=_threads())) < 1e-12
    r0 = r.cpu().numpy().copy()
    go.residual(x, r)
    assert _rel_np(r.cpu().numpy(), orc.residual(x.cpu().numpy(), r0, threads=_threads())) < 1e-12
    del y, r
    # ---- sampled rows of pattern + Jacobian
    rng = np.random.default_rng(5)
    L = 2 * C + 1                                     # lattice points per direction
    lat = rng.integers(0, L, size=(70000, 3))         # random rows (mostly interior, all entity groups)
    bnd = rng.integers(0, L, size=(30000, 3))         # rows on / next to the boundary: faces, edges, corners
    for d in range(3):
        sel = rng.random(30000) < 0.45
        bnd[sel, d] = rng.choice(np.array([0, 1, 2, L - 3, L - 2, L - 1]), size=int(sel.sum()))
    lat = np.concatenate([lat, bnd, np.array([[0, 0, 0], [L - 1, L - 1, L - 1], [0, L - 1, 1], [1, 1, 1], [2, 2, 2]])])
    # container index of a lattice point through the C ABI's own cell -> DOF map
    cells = np.minimum(lat // 2, C - 1)
    loc = lat - 2 * cells
    rows = np.empty(len(lat), dtype=np.uint64)
    cache = {}
    for i, (c, l) in enumerate(zip(cells, loc)):
        e = int(c[0] + C * (c[1] + C * c[2]))
        if e not in cache:
            cache[e] = go.cell_dof_indices(e)
        rows[i] = cache[e][int(l[0] + 3 * (l[1] + 3 * l[2]))]
    rows = np.unique(rows)
    rowlen, cols, vals = orc.jacobian_rows(rows, threads=_threads())
    nr, nnz = go.pattern_size()
    rowptr = torch.empty(nr + 1, dtype=torch.int64, device="cuda")
    colidx = torch.empty(nnz, dtype=torch.int32, device="cuda")
    go.fill_pattern(rowptr=rowptr, colidx=colidx, index32=True)
    values = torch.empty(nnz, dtype=torch.float64, device="cuda")
    go.jacobian(x, values, fresh=True)
    rows_t = torch.from_numpy(rows.astype(np.int64)).cuda()
    start = rowptr[rows_t]
    length = (rowptr[rows_t + 1] - start).cpu().numpy()
    assert np.array_equal(length, rowlen)                                  # row lengths: bit-exact
    maxlen = cols.shape[1]
    k = torch.arange(maxlen, device="cuda")[None, :]
    valid = k < torch.from_numpy(rowlen).cuda()[:, None]
    idx = torch.where(valid, start[:, None] + k, torch.zeros_like(k))
    got_c = torch.where(valid, colidx[idx].to(torch.int64), torch.zeros_like(idx)).cpu().numpy()
    got_v = torch.where(valid, values[idx], torch.zeros_like(values[idx])).cpu().numpy()
    assert np.array_equal(got_c.astype(np.uint64), cols)                   # column indices: bit-exact
    assert float(np.abs(got_v - vals).max() / np.abs(vals).max()) < 1e-12
    ncon = int(((vals == 1.0).sum(axis=1) == 1).sum() - 0)
    assert ncon > 1000                                                     # the sample does hold constrained unit rows


def test_cfg2_block_rows_of_the_dg_jacobian_equal_the_oracle_at_64(cuda_lib):
    """QkDG k=2 assembled Jacobian (block CSR) at 64^3 cells (51 GB at 128^3 would not fit beside the tests): sampled
    block rows — interior, boundary faces, edges, corners — against oracle_jacobian_rows."""
    from oracle import Oracle
    C = 64
    spec = _dg_spec((C, C, C), 2, abi.KERNEL_AUTO)
    go = GridOperator(spec)
    orc = Oracle(_np_spec(spec))
    rng = np.random.default_rng(3)
    cc = np.concatenate([rng.integers(0, C, size=(300, 3)),
                         np.array([[0, 0, 0], [C - 1, C - 1, C - 1], [0, 5, C - 1], [7, 0, 9], [C - 1, 3, 4]])])
    cells = np.unique(cc[:, 0] + C * (cc[:, 1] + C * cc[:, 2]))
    rows = (cells[:, None] * 27 + np.arange(27)[None, :]).reshape(-1).astype(np.uint64)
    rowlen, cols, vals = orc.jacobian_rows(rows, threads=_threads())
    nr, nnz = go.pattern_size()
    rowptr = torch.empty(nr + 1, dtype=torch.int64, device="cuda")
    colidx = torch.empty(nnz, dtype=torch.int32, device="cuda")
    go.fill_pattern(rowptr=rowptr, colidx=colidx, index32=True)
    values = torch.empty(nnz, dtype=torch.float64, device="cuda")
    go.jacobian(_rand(spec.num_dofs, 2), values, fresh=True)
    rp = rowptr.cpu().numpy()
    for s, r in enumerate(rows.astype(np.int64)):
        a, b = int(rp[r]), int(rp[r + 1])
        assert b - a == rowlen[s]
        assert np.array_equal(colidx[a:b].cpu().numpy().astype(np.uint64), cols[s, :b - a])
        assert _rel_np(values[a:b].cpu().numpy(), vals[s, :b - a]) < 1e-12

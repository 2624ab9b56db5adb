"""Sparsity pattern (bit-exact), assembled Jacobian (1e-12) and SpMV of the CUDA path against the
CPU oracle (-m gpu, through the C ABI)."""
import numpy as np
import pytest

from pdelab_b200 import abi
from problems import dg_problem, fem_problem, mt_vector, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
_id = lambda c: "-".join(f"{k}={v}" for k, v in c.items())

FEM_CASES = [
    dict(cells=(5, 4), degree=1), dict(cells=(7, 5), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(1, 1), degree=2, a="identity"), dict(cells=(2, 1, 1), degree=1, a="identity"),
    dict(cells=(4, 3, 3), degree=1, a="full", with_b=True, with_c=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(4, 3, 2), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(5, 4, 3), degree=2, a="scalar"),
    dict(cells=(6, 5), degree=2, a="scalar", bc="mixed", with_b=True),
    dict(cells=(4, 3, 3), degree=2, a="diagonal", bc="mixed", with_b=True, with_c=True),
    dict(cells=(5, 4, 3), degree=1, a="scalar", bc="mixed", with_b=True),
]
DG_CASES = [
    dict(cells=(5, 4), degree=1), dict(cells=(4, 3), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(3, 2), degree=3, a="diagonal"), dict(cells=(1, 1, 1), degree=2, a="scalar"),
    dict(cells=(3, 2, 2), degree=1, a="full", with_b=True, with_c=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(4, 3, 2), degree=2, a="scalar"),
    dict(cells=(3, 2, 2), degree=2, a="full", with_b=True, bc="mixed"),
    dict(cells=(3, 3, 2), degree=2, a="scalar", method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF),
]


def _ops(spec):
    from oracle import Oracle
    from pdelab_b200.capi import GridOperator
    return GridOperator(spec), Oracle(spec)


def _check(spec):
    go, orc = _ops(spec)
    rp_o, ci_o, va_o = orc.jacobian()
    nr, nnz = go.pattern_size()
    assert (nr, nnz) == (rp_o.size - 1, ci_o.size)
    rp, ci = go.fill_pattern()
    assert np.array_equal(rp, rp_o) and np.array_equal(ci, ci_o)          # bit-exact
    rp32, ci32 = go.fill_pattern(index32=True)
    assert np.array_equal(rp32, rp_o) and np.array_equal(ci32.astype(np.uint64), ci_o)
    x = mt_vector(spec.num_dofs)
    va = go.jacobian(x, np.zeros(nnz))
    assert rel_err(va, va_o) < TOL
    # accumulate semantics: values += J, except constrained rows which are reset to unit rows
    v0 = mt_vector(nnz, seed=5)
    va2 = go.jacobian(x, v0.copy())
    con = np.zeros(nr, dtype=bool)
    con[orc.constrained_dofs().astype(np.int64)] = True
    rows = np.repeat(np.arange(nr), np.diff(rp.astype(np.int64)))
    want = np.where(con[rows], va_o, va_o + v0)
    assert rel_err(va2, want) < TOL
    va3 = go.jacobian(x, np.full(nnz, np.nan), fresh=True)
    assert rel_err(va3, va_o) < TOL
    return go, orc, (rp, ci, va)


@pytest.mark.parametrize("case", FEM_CASES, ids=_id)
def test_fem_pattern_and_jacobian_match_oracle(cuda_lib, case):
    _check(fem_problem(**case))


@pytest.mark.parametrize("lines", [2, 3, 8])
@pytest.mark.parametrize("case", [dict(cells=(9, 7, 6), degree=2, a="scalar"), dict(cells=(6, 11, 5), degree=1, a="scalar"),
                                  dict(cells=(12, 19), degree=2, a="scalar"), dict(cells=(5, 9, 4), degree=2, a="identity")],
                         ids=_id)
def test_multi_line_values_kernel_on_small_grids(cuda_lib, case, lines, monkeypatch):
    """The interior values kernel walks several x-lines per CTA on big grids only (one line per CTA below ~600 lines);
    PDB200_QKV_LINES forces the walk - line counts that are no multiple of it, the double-buffered coefficient rows and
    the accumulate form - on grids the oracle assembles in full."""
    monkeypatch.setenv("PDB200_QKV_LINES", str(lines))
    _check(fem_problem(**case))


@pytest.mark.parametrize("case", DG_CASES, ids=_id)
def test_dg_pattern_and_jacobian_match_oracle(cuda_lib, case):
    spec = dg_problem(**case)
    go, orc, (rp, ci, va) = _check(spec)
    # block CSR (Blocking::fixed): same entries, blocks row-major and contiguous
    nbr, nblocks = go.pattern_size(block=True)
    assert nbr == spec.ncells
    brp, bci = go.fill_pattern(block=True)
    n = spec.local_size
    vb = go.jacobian(np.zeros(spec.num_dofs), np.zeros(nblocks * n * n), layout=abi.LAYOUT_BCSR)
    rpi = rp.astype(np.int64)
    for e in range(spec.ncells):
        cols = bci[int(brp[e]):int(brp[e + 1])].astype(np.int64)
        assert np.all(np.diff(cols) > 0)
        for s, cn in enumerate(cols):
            blk = vb[(int(brp[e]) + s) * n * n:(int(brp[e]) + s + 1) * n * n].reshape(n, n)
            for i in range(n):
                row = e * n + i
                seg = slice(rpi[row] + s * n, rpi[row] + (s + 1) * n)
                assert np.array_equal(ci[seg].astype(np.int64), cn * n + np.arange(n))
                assert np.array_equal(blk[i], va[seg])


@pytest.mark.parametrize("make,case", [(fem_problem, dict(cells=(9, 7, 5), degree=2, a="scalar", with_c=True)),
                                       (fem_problem, dict(cells=(40, 30), degree=1, a="full", with_b=True)),
                                       (dg_problem, dict(cells=(5, 4, 3), degree=2, a="full", with_b=True, with_c=True))],
                         ids=["q2-3d", "q1-2d", "dg-q2-3d"])
def test_assembled_matrix_times_vector_equals_jacobian_apply(cuda_lib, make, case):
    """jacobian_apply(z) == J z on unconstrained rows, with J z computed by the device SpMV."""
    import torch
    spec = make(**case)
    go, orc = _ops(spec)
    n = spec.num_dofs
    _, nnz = go.pattern_size()
    z = torch.from_numpy(mt_vector(n)).cuda()
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    go.jacobian(z, vals, fresh=True)
    y_mv, y_ap = torch.empty_like(z), torch.empty_like(z)
    go.csr_mv(vals, z, y_mv)
    go.apply(z, y_ap)
    go.synchronize()
    con = go.constrained_dofs().astype(np.int64)
    y_mv, y_ap = y_mv.cpu().numpy(), y_ap.cpu().numpy()
    assert np.allclose(y_mv[con], z.cpu().numpy()[con], rtol=0, atol=0)   # unit rows
    y_mv[con] = 0.0
    assert rel_err(y_mv, y_ap) < TOL
    if spec.space == abi.SPACE_QKDG:
        nb = go.pattern_size(block=True)[1]
        vb = torch.empty(nb * spec.local_size ** 2, dtype=torch.float64, device="cuda")
        go.jacobian(z, vb, layout=abi.LAYOUT_BCSR, fresh=True)
        yb = torch.empty_like(z)
        go.csr_mv(vb, z, yb, layout=abi.LAYOUT_BCSR)
        go.synchronize()
        assert rel_err(yb.cpu().numpy(), y_ap) < TOL


def test_cfg1_poisson_q1_2d_256(cuda_lib):
    """BASELINE.json configs[0]: Poisson Q1 on 256^2 — 66,049 DOFs, 591,361 non-zeros; pattern
    bit-exact and values/residual within 1e-12 of the oracle at full size."""
    spec = fem_problem((256, 256), degree=1, a="identity", with_f=True)
    go, orc = _ops(spec)
    assert go.pattern_size() == (66049, 591361)
    rp_o, ci_o, va_o = orc.jacobian()
    rp, ci = go.fill_pattern()
    assert np.array_equal(rp, rp_o) and np.array_equal(ci, ci_o)
    x = mt_vector(spec.num_dofs)
    assert rel_err(go.jacobian(x, np.zeros(ci.size)), va_o) < TOL
    assert rel_err(go.residual(x, np.zeros_like(x)), orc.residual(x)) < TOL


def test_q2_3d_pattern_counts_closed_form(cuda_lib):
    """Conforming Q2 on N^3 cells has (8N+1)^3 non-zeros (cfg4: 160^3 -> 1281^3; here N = 24)."""
    N = 24
    spec = fem_problem((N, N, N), degree=2, a="scalar", with_f=False)
    from pdelab_b200.capi import GridOperator
    go = GridOperator(spec)
    nr, nnz = go.pattern_size()
    assert nr == (2 * N + 1) ** 3 and nnz == (8 * N + 1) ** 3

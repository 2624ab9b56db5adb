"""oracle_jacobian_rows (sampled rows of pattern + Jacobian, used for parity at BASELINE sizes) is bit-identical
to the corresponding rows of the full oracle assembly on small grids."""
import numpy as np
import pytest

from oracle import Oracle
from problems import dg_problem, fem_problem

CASES = [
    ("dg_k1_2d", lambda: dg_problem((5, 4), degree=1, a="full", with_b=True, with_c=True, bc="mixed")),
    ("dg_k2_3d", lambda: dg_problem((3, 3, 2), degree=2, a="scalar")),
    ("dg_k2_3d_proc", lambda: dg_problem((3, 2, 3), degree=2, a="diagonal").replace(side_kind=[[0, 0], [1, 0], [0, 1]])),
    ("q1_2d", lambda: fem_problem((6, 5), degree=1, a="scalar", with_b=True, with_c=True, bc="mixed")),
    ("q2_2d", lambda: fem_problem((4, 3), degree=2, a="full", with_c=True)),
    ("q1_3d", lambda: fem_problem((3, 4, 2), degree=1, a="diagonal")),
    ("q2_3d", lambda: fem_problem((3, 2, 3), degree=2, a="scalar", with_b=True, bc="mixed")),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_sampled_rows_equal_the_full_assembly(name, make):
    spec = make()
    orc = Oracle(spec)
    rowptr, colidx, values = orc.jacobian()
    n = orc.num_dofs
    rows = np.arange(n, dtype=np.uint64)
    rowlen, cols, vals = orc.jacobian_rows(rows, threads=2)
    assert np.array_equal(rowlen, np.diff(rowptr.astype(np.int64)))
    for r in range(n):
        a, b = int(rowptr[r]), int(rowptr[r + 1])
        assert np.array_equal(cols[r, :b - a], colidx[a:b]), (name, r)
        assert np.array_equal(vals[r, :b - a], values[a:b]), (name, r)

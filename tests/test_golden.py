"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the
independent numpy assembly) against the CPU oracle, and with -m gpu against the CUDA path."""
import os

import numpy as np
import pytest

from problems import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES  # noqa: E402

TOL = 1e-12


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def _dense(rowptr, colidx, values, n):
    J = np.zeros((n, n))
    for r in range(n):
        s, e = int(rowptr[r]), int(rowptr[r + 1])
        J[r, colidx[s:e].astype(np.int64)] = values[s:e]
    return J


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    from oracle import Oracle
    spec, gold = CASES[name](), _load(name)
    orc = Oracle(spec)
    z = gold["z"]
    assert rel_err(orc.jacobian_apply(z), gold["jacobian_apply"]) < TOL
    assert rel_err(orc.residual(z), gold["residual"]) < TOL
    assert np.array_equal(orc.constrained_dofs().astype(np.int64), gold["constrained"])
    rowptr, colidx, values = orc.jacobian()
    assert rel_err(_dense(rowptr, colidx, values, z.size), gold["jacobian_dense"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_golden(cuda_lib, name):
    from pdelab_b200.capi import GridOperator
    spec, gold = CASES[name](), _load(name)
    go = GridOperator(spec)
    z = gold["z"]
    n = z.size
    assert rel_err(go.jacobian_apply(z, np.zeros(n)), gold["jacobian_apply"]) < TOL
    assert rel_err(go.residual(z, np.zeros(n)), gold["residual"]) < TOL
    assert np.array_equal(go.constrained_dofs().astype(np.int64), gold["constrained"])
    rowptr, colidx = go.fill_pattern()
    values = go.jacobian(z, np.zeros(colidx.size))
    assert rel_err(_dense(rowptr, colidx, values, n), gold["jacobian_dense"]) < TOL

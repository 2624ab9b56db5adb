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


# ---- one-step (Runge-Kutta stage) operator: fixtures of tests/golden/make_golden_onestep.py ---------------------------
from make_golden_onestep import CASES as ONESTEP_CASES  # noqa: E402


def _onestep_setup(name, make_ops):
    from pdelab_b200 import onestep as osm
    make, method_cls, stage = ONESTEP_CASES[name]
    gold = _load(name)
    spec0 = make()
    spec1 = osm.l2_spec(spec0, float(gold["scaling"]))
    op = make_ops(spec0, spec1)
    op.preStep(method_cls(), float(gold["time"]), float(gold["dt"]))
    op.preStage(stage, list(gold["xs"]))
    return op, gold, spec0


def _check_onestep(op, gold, residual, jacobian_apply, dense, const):
    x = gold["x"]
    scale = np.abs(gold["const_residual"]).max()
    assert np.abs(const - gold["const_residual"]).max() / scale < TOL
    assert rel_err(residual, gold["residual"]) < TOL
    assert rel_err(jacobian_apply, gold["jacobian_apply"]) < TOL
    assert rel_err(dense, gold["jacobian_dense"]) < TOL


@pytest.mark.parametrize("name", sorted(ONESTEP_CASES))
def test_onestep_oracle_reproduces_golden(name):
    from onestep_oracle import OneStepOracle
    op, gold, _ = _onestep_setup(name, lambda s0, s1: OneStepOracle(s0, s1))
    x = gold["x"]
    _check_onestep(op, gold, op.residual(x), op.jacobian_apply(x), op.matrix().toarray(), op.const)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ONESTEP_CASES))
def test_onestep_cuda_path_reproduces_golden(cuda_lib, name):
    from pdelab_b200 import onestep as osm
    from pdelab_b200.capi import GridOperator
    keep = []

    def make(s0, s1):
        keep.extend([GridOperator(s0), GridOperator(s1)])
        return osm.OneStepGridOperator(keep[0], keep[1])
    op, gold, spec0 = _onestep_setup(name, make)
    x = gold["x"]
    n = x.size
    rowptr, colidx = op.fill_pattern()
    values = op.jacobian(x, np.zeros(colidx.size))
    _check_onestep(op, gold, op.residual(x, np.zeros(n)), op.jacobian_apply(x, np.zeros(n)),
                   _dense(rowptr, colidx, values, n), op.const_residual(np.zeros(n)))

#!/usr/bin/env python
"""Generates the one-step golden fixtures tests/golden/onestep_*.npz.

Like make_golden.py these vectors come from the independent dense numpy assembly of the weak forms
(tests/numpy_assembly.py: spatial operator J0, R0(0) and the mass matrix of the L2 operator), combined with the stage
weights of gridoperator/onestep/{prestageengine,residualengine,jacobianengine}.hh written out by hand below — no code
shared with the oracle restatement (tests/onestep_oracle.py) or the CUDA path (csrc/onestep.cu).

    python tests/golden/make_golden_onestep.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "..", "dune-pdelab_b200", "python")]
from numpy_assembly import assemble  # noqa: E402
from pdelab_b200 import onestep as osm  # noqa: E402
from problems import dg_problem, fem_problem, mt_vector  # noqa: E402

CASES = {
    "onestep_dg_k2_3d_4x3x2_alexander3": (lambda: dg_problem((4, 3, 2), degree=2, a="scalar", with_f=True, bc="dirichlet_g"),
                                          osm.Alexander3Parameter, 3),
    "onestep_dg_k1_2d_5x4_fractionalstep": (lambda: dg_problem((5, 4), degree=1, a="diagonal", with_c=True, with_f=True,
                                                                bc="dirichlet_g"), osm.FractionalStepParameter, 2),
    "onestep_fem_q2_2d_4x3_alexander2": (lambda: fem_problem((4, 3), degree=2, a="diagonal"), osm.Alexander2Parameter, 2),
}
TIME, DT, SCALING = 0.25, 0.0625, 1.5


def stage_inputs(n, stage):
    return [mt_vector(n, seed=11 + i) - 0.5 for i in range(stage)], mt_vector(n, seed=31) - 0.5


def main():
    for name, (make, method_cls, stage) in CASES.items():
        spec0 = make()
        spec1 = osm.l2_spec(spec0, SCALING)
        J0, r00, con = assemble(spec0)
        M, r10, _ = assemble(spec1)
        assert np.abs(r10).max() == 0.0
        m = method_cls()
        n = J0.shape[0]
        xs, x = stage_inputs(n, stage)
        const = np.zeros(n)
        for i in range(stage):                                   # MultiplyOperator0ByDT: dt_factor0 = dt, dt_factor1 = 1
            a, b = m.a(stage, i), m.b(stage, i)
            if abs(b) > 1e-6:
                const += b * DT * (J0 @ xs[i] + r00)
            if abs(a) > 1e-6:
                const += a * (M @ xs[i])
        const[con] = 0.0
        w0 = m.b(stage, stage) * DT
        Js = w0 * J0 + M
        r = Js @ x + w0 * r00 + const
        r[con] = 0.0
        y = Js @ x
        y[con] = 0.0
        Jc = Js.copy()
        Jc[con, :] = 0.0
        Jc[con, con] = 1.0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), xs=np.array(xs), x=x, const_residual=const, residual=r,
                            jacobian_apply=y, jacobian_dense=Jc, constrained=np.flatnonzero(con),
                            stage=stage, time=TIME, dt=DT, scaling=SCALING)
        print(name, n, "dofs")


if __name__ == "__main__":
    main()

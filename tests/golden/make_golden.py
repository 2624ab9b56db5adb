#!/usr/bin/env python
"""Generates the golden fixtures tests/golden/*.npz.

The reference (C++, DUNE core modules absent) cannot be run here, so these vectors do NOT come from
dune-pdelab itself; they come from the independent dense numpy assembly of the weak forms
(tests/numpy_assembly.py), which shares no code with the oracle or the CUDA path.  Each file holds
the problem description, inputs and J z / residual / dense Jacobian so that both the oracle
(CPU suite) and the CUDA path (-m gpu) are checked against a committed artefact.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "..", "dune-pdelab_b200", "python")]
from numpy_assembly import apply_constraints, assemble  # noqa: E402
from problems import dg_problem, fem_problem, mt_vector  # noqa: E402

CASES = {
    "dg_k2_3d_4x3x2_scalar": lambda: dg_problem((4, 3, 2), degree=2, a="scalar", with_f=True, bc="dirichlet_g"),
    "dg_k2_3d_4x2x2_full_b_c_mixed": lambda: dg_problem((4, 2, 2), degree=2, a="full", with_b=True, with_c=True,
                                                        with_f=True, bc="mixed", extent=(1.0, 0.7, 1.3)),
    "dg_k1_2d_5x4_diag": lambda: dg_problem((5, 4), degree=1, a="diagonal", with_f=True, bc="dirichlet_g"),
    "dg_k4_3d_2x2x1_scalar": lambda: dg_problem((2, 2, 1), degree=4, a="scalar", with_f=True, bc="dirichlet_g"),
    "fem_q1_2d_6x5": lambda: fem_problem((6, 5), degree=1, a="scalar", with_c=True),
    "fem_q2_3d_3x2x2_mixed": lambda: fem_problem((3, 2, 2), degree=2, a="full", with_b=True, with_c=True, bc="mixed"),
    "fem_q2_2d_4x3": lambda: fem_problem((4, 3), degree=2, a="diagonal"),
}


def main():
    for name, make in CASES.items():
        spec = make()
        J, r0, con = assemble(spec)
        n = J.shape[0]
        z = mt_vector(n)
        y = J @ z
        y[con] = 0.0
        r = J @ z + r0
        r[con] = 0.0
        Jc = apply_constraints(J, r0, con)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), z=z, jacobian_apply=y, residual=r,
                            jacobian_dense=Jc, constrained=np.flatnonzero(con))
        print(name, n, "dofs")


if __name__ == "__main__":
    main()

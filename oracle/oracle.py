"""ctypes wrapper of oracle/pdelab_oracle.cc — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker and the CPU baseline, never the product.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "dune-pdelab_b200", "python"))
from pdelab_b200.abi import Problem, ProblemSpec  # noqa: E402  (data definitions only)

_u64p = C.POINTER(C.c_uint64)
_dp = C.c_void_p


def build(native=False):
    target = "_build/liboracle_native.so" if native else "_build/liboracle.so"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)
    return os.path.join(_HERE, target)


def load(native=False):
    path = os.path.join(_HERE, "_build", "liboracle_native.so" if native else "liboracle.so")
    src = os.path.join(_HERE, "pdelab_oracle.cc")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        path = build(native)
    lib = C.CDLL(path)
    lib.oracle_last_error.restype = C.c_char_p
    return lib


class Oracle:
    def __init__(self, spec: ProblemSpec, native=False):
        self.lib = load(native)
        self.spec = spec
        self.p = spec.c_struct()

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.oracle_last_error().decode())

    @property
    def num_dofs(self):
        n = C.c_uint64()
        self._chk(self.lib.oracle_num_dofs(C.byref(self.p), C.byref(n)))
        return n.value

    def quadrature(self):
        m = self.spec.m
        x, w = np.zeros(m), np.zeros(m)
        self._chk(self.lib.oracle_quadrature(C.byref(self.p), _dp(x.ctypes.data), _dp(w.ctypes.data)))
        return x, w

    def cell_dof_indices(self, cell):
        idx = np.zeros(self.spec.local_size, dtype=np.uint64)
        self._chk(self.lib.oracle_cell_dof_indices(C.byref(self.p), C.c_uint64(cell), _dp(idx.ctypes.data)))
        return idx

    def constrained_dofs(self):
        n = C.c_uint64()
        self._chk(self.lib.oracle_constrained_dofs(C.byref(self.p), C.byref(n), None))
        idx = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._chk(self.lib.oracle_constrained_dofs(C.byref(self.p), C.byref(n), _dp(idx.ctypes.data)))
        return idx

    def _vec(self, fn, x, out, *extra):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(self.num_dofs) if out is None else out
        assert x.size == out.size == self.num_dofs
        self._chk(fn(C.byref(self.p), _dp(x.ctypes.data), _dp(out.ctypes.data), *extra))
        return out

    def residual(self, x, r=None, threads=1):
        if threads > 1:
            return self._vec(self.lib.oracle_residual_mt, x, r, C.c_int(threads))
        return self._vec(self.lib.oracle_residual, x, r)

    def jacobian_apply(self, z, y=None, threads=1):
        if threads > 1:
            return self._vec(self.lib.oracle_jacobian_apply_mt, z, y, C.c_int(threads))
        return self._vec(self.lib.oracle_jacobian_apply, z, y)

    def fem_jacobian_apply_fd(self, x, y=None):
        return self._vec(self.lib.oracle_fem_jacobian_apply_fd, x, y)

    def max_threads(self):
        return self.lib.oracle_max_threads()

    def pattern(self):
        nrows, nnz = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.oracle_pattern(C.byref(self.p), C.byref(nrows), C.byref(nnz), None, None))
        rowptr = np.zeros(nrows.value + 1, dtype=np.uint64)
        colidx = np.zeros(nnz.value, dtype=np.uint64)
        self._chk(self.lib.oracle_pattern(C.byref(self.p), C.byref(nrows), C.byref(nnz),
                                          _dp(rowptr.ctypes.data), _dp(colidx.ctypes.data)))
        return rowptr, colidx

    def jacobian_rows(self, rows, threads=1):
        """Pattern and Jacobian of the sampled rows: list of (sorted column indices, values) per row,
        bit-identical to the corresponding rows of jacobian()."""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        s = self.spec
        maxlen = (2 * s.dim + 1) * s.local_size if s.space == 0 else (2 * s.degree + 1) ** s.dim
        rowlen = np.zeros(rows.size, dtype=np.uint64)
        colidx = np.zeros((rows.size, maxlen), dtype=np.uint64)
        values = np.zeros((rows.size, maxlen))
        self._chk(self.lib.oracle_jacobian_rows(C.byref(self.p), _dp(rows.ctypes.data), C.c_uint64(rows.size),
                                                C.c_int(threads), C.c_uint64(maxlen), _dp(rowlen.ctypes.data),
                                                _dp(colidx.ctypes.data), _dp(values.ctypes.data)))
        return rowlen.astype(np.int64), colidx, values

    def jacobian(self, x=None, values=None):
        rowptr, colidx = self.pattern()
        values = np.zeros(colidx.size) if values is None else values
        x = np.zeros(self.num_dofs) if x is None else np.ascontiguousarray(x, dtype=np.float64)
        self._chk(self.lib.oracle_jacobian(C.byref(self.p), _dp(x.ctypes.data), _dp(values.ctypes.data)))
        return rowptr, colidx, values

"""ctypes view of include/pdelab_b200.h (struct pdb200_problem and the enums).

Pure data definitions: no compute, no oracle, no torch.  Shared by the product binding
(`pdelab_b200.capi`) and by the test-only oracle wrapper (`oracle/oracle.py`).
"""
import ctypes as C

import numpy as np

SPACE_QKDG, SPACE_QK = 0, 1
DG_NIPG, DG_SIPG, DG_IIPG = 0, 1, 2
DG_WEIGHTS_ON, DG_WEIGHTS_OFF = 0, 1
BC_DIRICHLET, BC_NEUMANN, BC_OUTFLOW, BC_NONE = 1, -1, -2, -3
A_IDENTITY, A_SCALAR, A_DIAGONAL, A_FULL = 0, 1, 2, 3
SIDE_DOMAIN, SIDE_PROCESSOR = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FAST = 0, 1, 2
BASIS_LAGRANGE, BASIS_LEGENDRE, BASIS_LOBATTO = 0, 1, 2
LAYOUT_CSR, LAYOUT_BCSR = 0, 1
PART_ALL, PART_INTERIOR, PART_BOUNDARY = 0, 1, 2
SOLVER_BICGSTAB, SOLVER_CG = 0, 1
PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_JACOBI, PRECOND_BLOCK_SOR, PRECOND_BLOCK_SSOR = 0, 1, 2, 3, 4
SOR_BACKWARD, SOR_KEEP_ITERATE = 1, 2
POINTWISE_A, POINTWISE_B, POINTWISE_C, POINTWISE_BCTYPE = 1, 2, 4, 8


class Problem(C.Structure):
    """struct pdb200_problem (include/pdelab_b200.h)."""

    _fields_ = [
        ("dim", C.c_int32),
        ("cells", C.c_int32 * 3),
        ("lower", C.c_double * 3),
        ("upper", C.c_double * 3),
        ("space", C.c_int32),
        ("degree", C.c_int32),
        ("dg_method", C.c_int32),
        ("dg_weights", C.c_int32),
        ("dg_alpha", C.c_double),
        ("intorderadd", C.c_int32),
        ("a_mode", C.c_int32),
        ("A", C.c_void_p),
        ("b", C.c_void_p),
        ("c", C.c_void_p),
        ("f", C.c_void_p),
        ("bctype", C.c_void_p),
        ("g", C.c_void_p),
        ("j", C.c_void_p),
        ("o", C.c_void_p),
        ("side_kind", (C.c_int32 * 2) * 3),
        ("device", C.c_int32),
        ("kernel", C.c_int32),
        ("basis", C.c_int32),
        ("pointwise", C.c_int32),
    ]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (host or device)
        return a.data_ptr()
    raise TypeError(type(a))


class ProblemSpec:
    """Host-side description of one operator: grid, space, DG parameters and coefficient arrays.

    Mirrors what a PDELab program fixes through its typedefs (YaspGrid, QkDG/Qk finite element
    map, ConvectionDiffusionDG/FEM constructor arguments, parameter class).  Arrays are numpy
    (host) or torch (host/device) and are kept alive by this object.
    """

    def __init__(self, cells, space=SPACE_QKDG, degree=2, lower=None, upper=None,
                 method=DG_SIPG, weights=DG_WEIGHTS_ON, alpha=1.0, intorderadd=0,
                 a_mode=A_IDENTITY, A=None, b=None, c=None, f=None, bctype=None, g=None, j=None,
                 o=None, side_kind=None, device=0, kernel=KERNEL_AUTO, basis=BASIS_LAGRANGE, pointwise=0):
        self.cells = tuple(int(v) for v in cells)
        self.dim = len(self.cells)
        assert self.dim in (2, 3)
        self.space, self.degree = int(space), int(degree)
        self.lower = tuple(lower) if lower is not None else (0.0,) * self.dim
        self.upper = tuple(upper) if upper is not None else (1.0,) * self.dim
        self.method, self.weights, self.alpha = int(method), int(weights), float(alpha)
        self.intorderadd = int(intorderadd)
        self.a_mode = int(a_mode)
        self.arrays = dict(A=A, b=b, c=c, f=f, bctype=bctype, g=g, j=j, o=o)
        for k, v in self.arrays.items():
            if isinstance(v, np.ndarray):
                want = np.int8 if k == "bctype" else np.float64
                self.arrays[k] = np.ascontiguousarray(v, dtype=want)
        self.side_kind = side_kind if side_kind is not None else [[SIDE_DOMAIN] * 2 for _ in range(3)]
        self.device, self.kernel = int(device), int(kernel)
        self.basis = int(basis)
        self.pointwise = int(pointwise)  # POINTWISE_* bits: A, b, c, bctype sampled per quadrature point

    # sizes -------------------------------------------------------------------------------
    @property
    def ncells(self):
        return int(np.prod(self.cells))

    @property
    def local_size(self):
        return (self.degree + 1) ** self.dim

    @property
    def m(self):
        return (2 * self.degree + self.intorderadd) // 2 + 1

    @property
    def nq(self):
        return self.m ** self.dim

    @property
    def nfq(self):
        return self.m ** (self.dim - 1)

    @property
    def points_per_cell(self):
        """NP of the point-wise layouts: volume points, then the face points of the 2 dim faces."""
        return self.nq + 2 * self.dim * self.nfq

    def face_point(self, d, side, q):
        return self.nq + (2 * d + side) * self.nfq + q

    @property
    def num_boundary_faces(self):
        n = self.ncells
        return sum(2 * (n // self.cells[d]) for d in range(self.dim))

    def boundary_face_offset(self, d, side):
        off = 0
        for dd in range(self.dim):
            for s in range(2):
                if dd == d and s == side:
                    return off
                off += self.ncells // self.cells[dd]
        raise ValueError

    @property
    def num_dofs(self):
        if self.space == SPACE_QKDG:
            return self.ncells * self.local_size
        return int(np.prod([self.degree * n + 1 for n in self.cells]))

    def replace(self, **kw):
        import copy
        q = copy.copy(self)
        q.arrays = dict(self.arrays)
        for k, v in kw.items():
            if k in q.arrays:
                q.arrays[k] = v
            else:
                setattr(q, k, v)
        return q

    def c_struct(self):
        p = Problem()
        p.dim = self.dim
        for d in range(3):
            p.cells[d] = self.cells[d] if d < self.dim else 1
            p.lower[d] = self.lower[d] if d < self.dim else 0.0
            p.upper[d] = self.upper[d] if d < self.dim else 1.0
            for s in range(2):
                p.side_kind[d][s] = int(self.side_kind[d][s])
        p.space, p.degree = self.space, self.degree
        p.dg_method, p.dg_weights, p.dg_alpha = self.method, self.weights, self.alpha
        p.intorderadd, p.a_mode = self.intorderadd, self.a_mode
        for k, v in self.arrays.items():
            setattr(p, k, _ptr(v))
        p.device, p.kernel = self.device, self.kernel
        p.basis = getattr(self, "basis", BASIS_LAGRANGE)
        p.pointwise = getattr(self, "pointwise", 0)
        return p

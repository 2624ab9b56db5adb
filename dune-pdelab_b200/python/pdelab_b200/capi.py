"""ctypes binding of libpdelab_b200.so and the host-side mirror of PDELab's GridOperator.

`GridOperator` keeps the reference's method names and argument meaning
(dune/pdelab/gridoperator/gridoperator.hh:167-205): residual(x, r), jacobian_apply(z, y),
jacobian(x, A), fill_pattern(); results are accumulated, errors are raised as exceptions
(`PDELabError` replaces Dune::Exception).  There is no CPU implementation behind it: if the
CUDA library is missing the import fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import Problem, ProblemSpec

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libpdelab_b200.so"))


class PDELabError(RuntimeError):
    """Raised where the reference throws Dune::Exception."""


_lib = None

# every symbol declared in include/pdelab_b200.h (checked by tests/test_abi.py)
SYMBOLS = [
    "pdb200_last_error", "pdb200_create", "pdb200_destroy", "pdb200_update_coefficients",
    "pdb200_num_dofs", "pdb200_local_size", "pdb200_num_boundary_faces", "pdb200_boundary_face_offset",
    "pdb200_quadrature_size", "pdb200_quadrature", "pdb200_gauss_legendre", "pdb200_cell_dof_indices", "pdb200_constrained_dofs",
    "pdb200_residual", "pdb200_jacobian_apply", "pdb200_onthefly_apply", "pdb200_jacobian_apply_nonlinear",
    "pdb200_pattern_size", "pdb200_pattern", "pdb200_pattern_i32", "pdb200_block_pattern_size",
    "pdb200_block_pattern", "pdb200_jacobian", "pdb200_jacobian_fresh", "pdb200_csr_mv",
    "pdb200_solve", "pdb200_solve_stationary", "pdb200_block_jacobi_apply", "pdb200_point_diagonal",
    "pdb200_halo_layer_size", "pdb200_halo_pack", "pdb200_halo_unpack", "pdb200_set_stream",
    "pdb200_gather_dofs", "pdb200_scatter_dofs",
    "pdb200_onthefly_apply_part", "pdb200_halo_p2p_create", "pdb200_halo_p2p_connect",
    "pdb200_halo_exchange_p2p", "pdb200_onthefly_apply_p2p",
    "pdb200_synchronize", "pdb200_launch_count", "pdb200_last_kernel", "pdb200_version",
    "pdb200_comm_create", "pdb200_comm_connect", "pdb200_comm_sum", "pdb200_solve_ovlp",
    "pdb200_block_diagonal_apply", "pdb200_block_offdiagonal_apply", "pdb200_block_sor_apply", "pdb200_set_relaxation",
    # OneStepGridOperator (bound in pdelab_b200.onestep)
    "pdb200_onestep_create", "pdb200_onestep_destroy", "pdb200_onestep_set_method", "pdb200_onestep_set_dt_mode",
    "pdb200_onestep_pre_step", "pdb200_onestep_time_at_stage", "pdb200_onestep_pre_stage",
    "pdb200_onestep_pre_stage_begin", "pdb200_onestep_pre_stage_add", "pdb200_onestep_const_residual",
    "pdb200_onestep_explicit_stage", "pdb200_onestep_explicit_stage_begin", "pdb200_onestep_explicit_stage_add",
    "pdb200_onestep_explicit_stage_finish",
    "pdb200_onestep_residual", "pdb200_onestep_jacobian_apply", "pdb200_onestep_onthefly_apply",
    "pdb200_onestep_jacobian", "pdb200_onestep_stage_operator", "pdb200_onestep_solve_stationary",
    "pdb200_onestep_launch_count",
]


class SolveResult(C.Structure):
    """pdb200_solve_result = LinearSolverResult (backend/solver.hh:28-51) + the defect norms."""
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_uint32), ("elapsed", C.c_double),
                ("reduction", C.c_double), ("conv_rate", C.c_double), ("first_defect", C.c_double),
                ("defect", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def load_library():
    """Load the CUDA library (built in-tree by __graft_entry__.build()).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PDB200_LIB", LIB_PATH)  # tuning aid: A/B of two builds on the same box (tools/)
    if not os.path.exists(path):
        raise PDELabError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(pdelab_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    lib.pdb200_last_error.restype = C.c_char_p
    lib.pdb200_last_kernel.restype = C.c_char_p
    lib.pdb200_last_kernel.argtypes = [C.c_void_p]
    lib.pdb200_version.restype = C.c_char_p
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    lib.pdb200_create.argtypes = [C.POINTER(Problem), C.POINTER(C.c_void_p)]
    for name in ("pdb200_residual", "pdb200_jacobian_apply", "pdb200_onthefly_apply"):
        getattr(lib, name).argtypes = [vp, vp, vp]
    lib.pdb200_jacobian_apply_nonlinear.argtypes = [vp, vp, vp, vp]
    lib.pdb200_jacobian.argtypes = [vp, vp, vp, C.c_int]
    lib.pdb200_jacobian_fresh.argtypes = [vp, vp, vp, C.c_int]
    lib.pdb200_csr_mv.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.pdb200_block_jacobi_apply.argtypes = [vp, vp, vp]
    lib.pdb200_point_diagonal.argtypes = [vp, vp]
    lib.pdb200_block_diagonal_apply.argtypes = [vp, vp, vp]
    lib.pdb200_block_offdiagonal_apply.argtypes = [vp, vp, vp]
    lib.pdb200_block_sor_apply.argtypes = [vp, vp, vp, C.c_double, C.c_int]
    lib.pdb200_set_relaxation.argtypes = [vp, C.c_double]
    lib.pdb200_solve.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_double, C.c_uint32,
                                 C.POINTER(SolveResult)]
    lib.pdb200_solve_stationary.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_double, C.c_double, C.c_uint32,
                                            C.POINTER(SolveResult)]
    lib.pdb200_pattern_size.argtypes = [vp, u64p, u64p]
    lib.pdb200_block_pattern_size.argtypes = [vp, u64p, u64p]
    for name in ("pdb200_pattern", "pdb200_pattern_i32", "pdb200_block_pattern"):
        getattr(lib, name).argtypes = [vp, vp, vp]
    lib.pdb200_halo_layer_size.argtypes = [vp, C.c_int, u64p]
    lib.pdb200_halo_pack.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.pdb200_halo_unpack.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.pdb200_set_stream.argtypes = [vp, vp]
    lib.pdb200_gather_dofs.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.pdb200_scatter_dofs.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.pdb200_onthefly_apply_part.argtypes = [vp, vp, vp, C.c_int]
    lib.pdb200_halo_p2p_create.argtypes = [vp, vp]
    lib.pdb200_halo_p2p_connect.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.pdb200_halo_exchange_p2p.argtypes = [vp, vp]
    lib.pdb200_onthefly_apply_p2p.argtypes = [vp, vp, vp]
    lib.pdb200_comm_create.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.pdb200_comm_connect.argtypes = [vp, C.c_int, vp]
    lib.pdb200_comm_sum.argtypes = [vp, vp, C.c_int]
    lib.pdb200_solve_ovlp.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_double, C.c_uint32,
                                      C.POINTER(SolveResult)]
    lib.pdb200_cell_dof_indices.argtypes = [vp, C.c_uint64, vp]
    lib.pdb200_constrained_dofs.argtypes = [vp, u64p, vp]
    lib.pdb200_quadrature.argtypes = [vp, vp, vp]
    lib.pdb200_gauss_legendre.argtypes = [C.c_int, vp, vp]
    lib.pdb200_boundary_face_offset.argtypes = [vp, C.c_int, C.c_int, u64p]
    for name in ("pdb200_destroy", "pdb200_synchronize"):
        getattr(lib, name).argtypes = [vp]
    lib.pdb200_update_coefficients.argtypes = [vp, C.POINTER(Problem)]
    lib.pdb200_num_dofs.argtypes = [vp, u64p]
    lib.pdb200_num_boundary_faces.argtypes = [vp, u64p]
    lib.pdb200_launch_count.argtypes = [vp, u64p]
    lib.pdb200_local_size.argtypes = [vp, C.POINTER(C.c_uint32)]
    lib.pdb200_quadrature_size.argtypes = [vp, C.POINTER(C.c_uint32)]
    _lib = lib
    return lib


def _ptr(a):
    """Raw address of a numpy array (host) or torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype in (np.float64, np.uint64, np.uint32, np.int64) and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(f"expected numpy array or torch tensor, got {type(a)}")


class GridOperator:
    """Host mirror of Dune::PDELab::GridOperator for the CUDA path.

    Vectors are flat float64 arrays in the reference's container order
    (ISTL::BlockVector, backend/istl/vector.hh): numpy arrays (host memory, staged through the
    device inside the call) or torch CUDA tensors (used in place).
    """

    def __init__(self, spec: ProblemSpec):
        self.lib = load_library()
        self.spec = spec
        self._p = spec.c_struct()
        h = C.c_void_p()
        self._h = None
        self._chk(self.lib.pdb200_create(C.byref(self._p), C.byref(h)))
        self._h = h

    def _chk(self, rc):
        if rc != 0:
            raise PDELabError(self.lib.pdb200_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.pdb200_destroy(self._h)
                self._h = None
        except Exception:
            pass

    close = __del__

    # sizes ---------------------------------------------------------------------------------
    def globalSizeU(self):
        n = C.c_uint64()
        self._chk(self.lib.pdb200_num_dofs(self._h, C.byref(n)))
        return n.value

    globalSizeV = globalSizeU

    def quadrature(self):
        m = C.c_uint32()
        self._chk(self.lib.pdb200_quadrature_size(self._h, C.byref(m)))
        x, w = np.zeros(m.value), np.zeros(m.value)
        self._chk(self.lib.pdb200_quadrature(self._h, x.ctypes.data, w.ctypes.data))
        return x, w

    def boundary_face_offset(self, d, side):
        n = C.c_uint64()
        self._chk(self.lib.pdb200_boundary_face_offset(self._h, d, side, C.byref(n)))
        return n.value

    def cell_dof_indices(self, cell):
        idx = np.zeros(self.spec.local_size, dtype=np.uint64)
        self._chk(self.lib.pdb200_cell_dof_indices(self._h, cell, idx.ctypes.data))
        return idx

    def constrained_dofs(self):
        n = C.c_uint64()
        self._chk(self.lib.pdb200_constrained_dofs(self._h, C.byref(n), None))
        idx = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._chk(self.lib.pdb200_constrained_dofs(self._h, C.byref(n), idx.ctypes.data))
        return idx

    # the hot path --------------------------------------------------------------------------
    def residual(self, x, r):
        """r += R(x); constrained rows := 0   (gridoperator.hh:176-181)."""
        self._chk(self.lib.pdb200_residual(self._h, _ptr(x), _ptr(r)))
        return r

    def jacobian_apply(self, *args):
        """jacobian_apply(z, y): y += J z.  jacobian_apply(u, z, y) raises for these linear
        local operators exactly like the reference (gridoperator.hh:192-205)."""
        if len(args) == 3:
            u, z, y = args
            self._chk(self.lib.pdb200_jacobian_apply_nonlinear(self._h, _ptr(u), _ptr(z), _ptr(y)))
            return y
        z, y = args
        self._chk(self.lib.pdb200_jacobian_apply(self._h, _ptr(z), _ptr(y)))
        return y

    def apply(self, x, y):
        """OnTheFlyOperator::apply: y = J x  (backend/istl/seqistlsolverbackend.hh:66-76)."""
        self._chk(self.lib.pdb200_onthefly_apply(self._h, _ptr(x), _ptr(y)))
        return y

    def update_coefficients(self, **arrays):
        spec = self.spec.replace(**{k: None for k in self.spec.arrays})
        for k, v in arrays.items():
            spec.arrays[k] = v
        p = spec.c_struct()
        self._chk(self.lib.pdb200_update_coefficients(self._h, C.byref(p)))

    # matrix --------------------------------------------------------------------------------
    def pattern_size(self, block=False):
        nr, nnz = C.c_uint64(), C.c_uint64()
        fn = self.lib.pdb200_block_pattern_size if block else self.lib.pdb200_pattern_size
        self._chk(fn(self._h, C.byref(nr), C.byref(nnz)))
        return nr.value, nnz.value

    def fill_pattern(self, block=False, rowptr=None, colidx=None, index32=False):
        """Sparsity pattern (CSR over DOFs, or block CSR over cells) as (rowptr, colidx)."""
        nr, nnz = self.pattern_size(block)
        if rowptr is None:
            rowptr = np.zeros(nr + 1, dtype=np.uint64)
        if colidx is None:
            colidx = np.zeros(nnz, dtype=np.uint32 if index32 else np.uint64)
        if block:
            fn = self.lib.pdb200_block_pattern
        else:
            fn = self.lib.pdb200_pattern_i32 if index32 else self.lib.pdb200_pattern
        self._chk(fn(self._h, _ptr(rowptr), _ptr(colidx)))
        return rowptr, colidx

    def jacobian(self, x, values, layout=abi.LAYOUT_CSR, fresh=False):
        """values += dR/dx (fresh=False, gridoperator.hh:184-189) or values = dR/dx (fresh=True:
        `A = 0; go.jacobian(x, A)`, stationary/linearproblem.hh:221-226)."""
        fn = self.lib.pdb200_jacobian_fresh if fresh else self.lib.pdb200_jacobian
        self._chk(fn(self._h, _ptr(x), _ptr(values), layout))
        return values

    def csr_mv(self, values, x, y, layout=abi.LAYOUT_CSR):
        self._chk(self.lib.pdb200_csr_mv(self._h, _ptr(values), layout, _ptr(x), _ptr(y)))
        return y

    # linear solvers (device-resident Krylov loops) -----------------------------------------
    def solve(self, z, r, reduction, solver=abi.SOLVER_BICGSTAB, precond=abi.PRECOND_NONE, values=None,
              layout=abi.LAYOUT_CSR, maxiter=5000):
        """Solve J z = r like the reference's ISTL back-ends (seqistlsolverbackend.hh): matrix-free
        (values=None: ISTLBackend_SEQ_MatrixFree_BCGS_Richardson) or with the assembled matrix
        (ISTLBackend_SEQ_BCGS_Jac / _CG_Jac).  z: initial guess in, solution out; r: defect out."""
        res = SolveResult()
        self._chk(self.lib.pdb200_solve(self._h, solver, precond, _ptr(values), layout, _ptr(z), _ptr(r),
                                        float(reduction), int(maxiter), C.byref(res)))
        return res.as_dict()

    def block_jacobi_apply(self, r, z):
        """z = D^-1 r, D = block diagonal of the QkDG Jacobian (AssembledBlockJacobiPreconditionerLocalOperator,
        backend/istl/matrixfree/assembledblockjacobipreconditioner.hh:96-230), matrix-free."""
        self._chk(self.lib.pdb200_block_jacobi_apply(self._h, _ptr(r), _ptr(z)))
        return z

    def block_diagonal_apply(self, z, y):
        """y = D z (BlockDiagonalLocalOperatorWrapper, localoperator/blockdiagonalwrapper.hh), matrix-free."""
        self._chk(self.lib.pdb200_block_diagonal_apply(self._h, _ptr(z), _ptr(y)))
        return y

    def block_offdiagonal_apply(self, z, y):
        """y = (J - D) z (BlockOffDiagonalLocalOperatorWrapper, localoperator/blockoffdiagonalwrapper.hh)."""
        self._chk(self.lib.pdb200_block_offdiagonal_apply(self._h, _ptr(z), _ptr(y)))
        return y

    def block_sor_apply(self, d, v, omega=1.0, backward=False, keep_iterate=False):
        """One block SOR sweep in index-set order (BlockSORPreconditionerLocalOperator,
        backend/istl/matrixfree/blocksorpreconditioner.hh), matrix-free, in place in v."""
        flags = (abi.SOR_BACKWARD if backward else 0) | (abi.SOR_KEEP_ITERATE if keep_iterate else 0)
        self._chk(self.lib.pdb200_block_sor_apply(self._h, _ptr(d), _ptr(v), float(omega), flags))
        return v

    def set_relaxation(self, omega):
        self._chk(self.lib.pdb200_set_relaxation(self._h, float(omega)))

    def point_diagonal(self, d):
        """d = diag(J), matrix-free (PointDiagonalLocalOperatorWrapper, localoperator/pointdiagonalwrapper.hh);
        1 on constrained rows."""
        self._chk(self.lib.pdb200_point_diagonal(self._h, _ptr(d)))
        return d

    def solve_stationary(self, x, reduction=1e-10, min_defect=1e-99, solver=abi.SOLVER_BICGSTAB,
                         precond=abi.PRECOND_NONE, matrix_free=True, maxiter=5000):
        """StationaryLinearProblemSolver::apply (stationary/linearproblem.hh:188-302): x -= J^-1 R(x)."""
        res = SolveResult()
        self._chk(self.lib.pdb200_solve_stationary(self._h, solver, precond, 1 if matrix_free else 0, _ptr(x),
                                                   float(reduction), float(min_defect), int(maxiter), C.byref(res)))
        return res.as_dict()

    # halo ----------------------------------------------------------------------------------
    def halo_layer_size(self, d):
        n = C.c_uint64()
        self._chk(self.lib.pdb200_halo_layer_size(self._h, d, C.byref(n)))
        return n.value

    def halo_pack(self, x, d, side, buf):
        self._chk(self.lib.pdb200_halo_pack(self._h, _ptr(x), d, side, _ptr(buf)))

    def halo_unpack(self, x, d, side, buf):
        self._chk(self.lib.pdb200_halo_unpack(self._h, _ptr(x), d, side, _ptr(buf)))

    def apply_part(self, x, y, part):
        """y = J x restricted to the INTERIOR or BOUNDARY tiles of the local box (abi.PART_*)."""
        self._chk(self.lib.pdb200_onthefly_apply_part(self._h, _ptr(x), _ptr(y), part))
        return y

    def gather_dofs(self, x, idx, buf):
        """buf[i] = x[idx[i]] on the device (idx: int64 CUDA tensor)."""
        self._chk(self.lib.pdb200_gather_dofs(self._h, _ptr(x), idx.data_ptr(), idx.numel(), _ptr(buf)))

    def scatter_dofs(self, buf, idx, x):
        """x[idx[i]] = buf[i] on the device."""
        self._chk(self.lib.pdb200_scatter_dofs(self._h, _ptr(buf), idx.data_ptr(), idx.numel(), _ptr(x)))

    def halo_p2p_create(self):
        """Create this rank's mailbox; returns the 64-byte IPC handle to give to the neighbours."""
        buf = C.create_string_buffer(64)
        self._chk(self.lib.pdb200_halo_p2p_create(self._h, buf))
        return buf.raw

    def halo_p2p_connect(self, d, side, handle_bytes):
        buf = C.create_string_buffer(bytes(handle_bytes), 64)
        self._chk(self.lib.pdb200_halo_p2p_connect(self._h, d, side, buf))

    def halo_exchange_p2p(self, x):
        self._chk(self.lib.pdb200_halo_exchange_p2p(self._h, _ptr(x)))

    def apply_p2p(self, x, y):
        """y = J x on the overlapping partition, exchange hidden behind the interior tiles."""
        self._chk(self.lib.pdb200_onthefly_apply_p2p(self._h, _ptr(x), _ptr(y)))
        return y

    # overlapping solvers ---------------------------------------------------------------------
    def comm_create(self, rank, size):
        """Create this rank's reduction mailbox; returns the 64-byte IPC handle for the other ranks."""
        buf = C.create_string_buffer(64)
        self._chk(self.lib.pdb200_comm_create(self._h, rank, size, buf))
        return buf.raw

    def comm_connect(self, peer_rank, handle_bytes):
        buf = C.create_string_buffer(bytes(handle_bytes), 64)
        self._chk(self.lib.pdb200_comm_connect(self._h, peer_rank, buf))

    def comm_sum(self, values):
        """gridView().comm().sum of one or two float64 values (numpy array, in place), collective."""
        self._chk(self.lib.pdb200_comm_sum(self._h, _ptr(values), int(values.size)))
        return values

    def solve_ovlp(self, z, r, reduction, solver=abi.SOLVER_BICGSTAB, precond=abi.PRECOND_NONE, values=None,
                   layout=abi.LAYOUT_CSR, maxiter=5000):
        """pdb200_solve on the overlapping partition (OverlappingOperator + OverlappingScalarProduct,
        backend/istl/ovlpistlsolverbackend.hh:30-134): collective, device vectors, z consistent on return."""
        res = SolveResult()
        self._chk(self.lib.pdb200_solve_ovlp(self._h, solver, precond, _ptr(values), layout, _ptr(z), _ptr(r),
                                             float(reduction), int(maxiter), C.byref(res)))
        return res.as_dict()

    # misc ----------------------------------------------------------------------------------
    def set_stream(self, stream_ptr):
        self._chk(self.lib.pdb200_set_stream(self._h, stream_ptr))

    def synchronize(self):
        self._chk(self.lib.pdb200_synchronize(self._h))

    def launch_count(self):
        n = C.c_uint64()
        self._chk(self.lib.pdb200_launch_count(self._h, C.byref(n)))
        return n.value

    def last_kernel(self):
        return self.lib.pdb200_last_kernel(self._h).decode()

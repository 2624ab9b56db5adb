/*
 * pdelab_b200.h — C ABI of the B200-native operator-evaluation path.
 *
 * This is the drop-in boundary.  The reference (dune-pdelab) has no FFI of its
 * own: the path lives behind the C++ class template Dune::PDELab::GridOperator
 * (dune/pdelab/gridoperator/gridoperator.hh:30-35).  Each entry point below names the reference
 * member it replaces; the C++ mirror in dune-pdelab_b200/host/gridoperator.hh forwards to these
 * symbols, and INTEGRATION.md shows the binding a PDELab maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only, no C++/torch types;
 *  - every function returns 0 on success, non-zero on error; pdb200_last_error() gives the text
 *    (the reference throws Dune::Exception, e.g. gridoperator.hh:195; the C++ mirror rethrows);
 *  - vector/matrix pointers may be HOST or DEVICE memory (detected with
 *    cudaPointerGetAttributes).  Host pointers are staged through pinned buffers and the call is
 *    synchronous on return; device pointers run on the handle's stream and are asynchronous;
 *  - results are ACCUMULATED into r / y / values exactly like the reference engines
 *    (gridoperator/default/residualengine.hh:184-188, jacobianapplyengine.hh:197-202);
 *  - there is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 *
 * Data model of the coefficient fields.  The reference takes a C++ parameter class with call-backs
 * (localoperator/convectiondiffusionparameter.hh:135-209).  Across a C ABI these become arrays.
 * Two layouts exist per field, selected by the bits of pdb200_problem::pointwise:
 *
 * (1) cell-/face-wise layout (bit clear) — for fields that are constant on every cell / boundary face, which is
 *     what the Kronecker fast paths need:
 *   A       per cell (cell centre; permeabilityIsConstantPerCell()==true, :139-142)
 *   b, c    per cell (cell-wise constant fields: the value the call-back returns at EVERY point of the cell)
 *   bctype  per boundary face (int8: Dirichlet=1, Neumann=-1, Outflow=-2, None=-3, :111-115)
 * (2) point-wise layout (bit set) — sampled exactly where the reference evaluates the call-backs.  A cell has
 *     NP = nq + 2*dim*nfq sample points: first its nq = m^dim volume quadrature points, then for every face
 *     (dir, side) in the order 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z its nfq = m^(dim-1) face quadrature points,
 *         pt = q                                 volume point   (convectiondiffusiondg.hh:143-146,178,181)
 *         pt = nq + (2*dir + side)*nfq + q       face point, in the local coordinates of THIS cell
 *                                                (geo_in_inside / geo_in_outside.global(ip), :370-371,426,755,797)
 *   A       A[(cell*NP + pt) * {1, dim, dim*dim}]   permeabilityIsConstantPerCell()==false: volume points, and on
 *                                                   every face the points seen from both adjacent cells
 *   b       b[(cell*NP + pt) * dim]                 volume points; face points of the lower faces (the assembler
 *                                                   visits an interior face from the larger-index cell, :426) and
 *                                                   of boundary faces; the other slots are never read
 *   c       c[cell*nq + q]                          volume points
 *   bctype  bctype[bface*nfq + q]                   QkDG only (:763,979); ConvectionDiffusionFEM evaluates the
 *                                                   type at the face centre (convectiondiffusionfem.hh:226-229)
 *     Point-wise fields run through the reference-order kernels (no Kronecker form exists for them).
 * Always per point:
 *   f       per cell and volume quadrature point  f[cell*nq + q],  q = q0 + m*(q1 + m*q2)
 *   g,j,o   per boundary face and face quadrature point  g[bface*nfq + q], q = t0 + m*t1 over the
 *           tangential directions in increasing order
 * with m = (2*degree+intorderadd)/2 + 1 Gauss-Legendre points per direction in ascending order
 * (convectiondiffusiondg.hh:139-140).  Boundary faces are numbered direction-major:
 * for d in 0..dim-1, for side in {0 (lower), 1 (upper)}: faces lexicographic in the tangential
 * cell coordinates; pdb200_boundary_face_offset() returns the first index of each (d,side) group.
 */
#ifndef PDELAB_B200_H
#define PDELAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* finite element space: finiteelementmap/qkdg.hh:36-76 (QkDG, Lagrange basis) and
 * finiteelementmap/qkfem.hh:17-78 (conforming Qk, k in {1,2}) */
enum { PDB200_SPACE_QKDG = 0, PDB200_SPACE_QK = 1 };
/* ConvectionDiffusionDGMethod::Type, convectiondiffusiondg.hh:31 */
enum { PDB200_DG_NIPG = 0, PDB200_DG_SIPG = 1, PDB200_DG_IIPG = 2 };
/* ConvectionDiffusionDGWeights::Type, convectiondiffusiondg.hh:36 */
enum { PDB200_DG_WEIGHTS_ON = 0, PDB200_DG_WEIGHTS_OFF = 1 };
/* ConvectionDiffusionBoundaryConditions::Type, convectiondiffusionparameter.hh:113 */
enum { PDB200_BC_DIRICHLET = 1, PDB200_BC_NEUMANN = -1, PDB200_BC_OUTFLOW = -2, PDB200_BC_NONE = -3 };
/* layout of the diffusion tensor array */
enum { PDB200_A_IDENTITY = 0, PDB200_A_SCALAR = 1, PDB200_A_DIAGONAL = 2, PDB200_A_FULL = 3 };
/* kind of the six outer sides of the (local) grid box */
enum { PDB200_SIDE_DOMAIN = 0,     /* physical boundary: alpha_boundary is integrated            */
       PDB200_SIDE_PROCESSOR = 1   /* processor boundary of an overlapping partition: nothing is
                                      integrated (default/assembler.hh:239-250) and the cells that
                                      touch it are constrained (constraints/p0.hh:31-41)          */ };
/* kernel selection */
enum { PDB200_KERNEL_AUTO = 0, PDB200_KERNEL_GENERIC = 1, PDB200_KERNEL_FAST = 2 };
/* QkDGBasisPolynomial, finiteelementmap/qkdg.hh:15 — the 1-D polynomials of the QkDG tensor-product basis:
 *   LAGRANGE  equidistant nodes j/k                         finiteelement/qkdglagrange.hh:55-79
 *   LEGENDRE  shifted Legendre polynomials P_n(2x - 1)      finiteelement/qkdglegendre.hh:76-139 (not normalised)
 *   LOBATTO   Lagrange polynomials on the Gauss-Lobatto points of [0,1], ascending   finiteelement/qkdglobatto.hh:20-105
 * Local DOF i <-> multi-index (i mod (k+1), ...), x fastest, for all three.  Conforming Qk spaces are Lagrange. */
enum { PDB200_BASIS_LAGRANGE = 0, PDB200_BASIS_LEGENDRE = 1, PDB200_BASIS_LOBATTO = 2 };
/* pdb200_problem::pointwise — which coefficient arrays use the point-wise layout (header comment, layout (2)) */
enum { PDB200_POINTWISE_A = 1, PDB200_POINTWISE_B = 2, PDB200_POINTWISE_C = 4, PDB200_POINTWISE_BCTYPE = 8 };

typedef struct pdb200_problem {
  int32_t dim;            /* 2 or 3                                                              */
  int32_t cells[3];       /* YaspGrid cells per direction (of the LOCAL box incl. overlap)       */
  double  lower[3];       /* lower-left corner of the local box                                  */
  double  upper[3];       /* upper-right corner of the local box                                 */
  int32_t space;          /* PDB200_SPACE_*                                                      */
  int32_t degree;         /* k                                                                   */
  int32_t dg_method;      /* PDB200_DG_*      (ConvectionDiffusionDG ctor, :85-102)              */
  int32_t dg_weights;     /* PDB200_DG_WEIGHTS_*                                                 */
  double  dg_alpha;       /* penalty constant alpha                                              */
  int32_t intorderadd;    /* extra quadrature order                                              */
  int32_t a_mode;         /* PDB200_A_*                                                          */
  const double* A;        /* [cells * {1, dim, dim*dim}] (row-major tensor) or NULL; point-wise: [cells * NP * ...] */
  const double* b;        /* [cells * dim] or NULL (= 0); point-wise: [cells * NP * dim]          */
  const double* c;        /* [cells] or NULL (= 0); point-wise: [cells * m^dim]                   */
  const double* f;        /* [cells * m^dim] or NULL (= 0)                                       */
  const int8_t* bctype;   /* [boundary faces] or NULL (= all Dirichlet); point-wise: [faces * m^(dim-1)] */
  const double* g;        /* [boundary faces * m^(dim-1)] or NULL (= 0)                          */
  const double* j;        /* same layout, Neumann flux, or NULL                                  */
  const double* o;        /* same layout, outflow flux, or NULL                                  */
  int32_t side_kind[3][2];/* PDB200_SIDE_* for (direction, lower/upper)                          */
  int32_t device;         /* CUDA device ordinal                                                 */
  int32_t kernel;         /* PDB200_KERNEL_*                                                     */
  int32_t basis;          /* PDB200_BASIS_* (QkDG only; 0 = Lagrange)                            */
  int32_t pointwise;      /* bit mask of PDB200_POINTWISE_*: arrays in the point-wise layout     */
} pdb200_problem;

typedef struct pdb200_operator* pdb200_handle;

/* text of the last error on the calling thread (replaces Dune::Exception::what()) */
const char* pdb200_last_error(void);

/* GridOperator::GridOperator(gfsu,cu,gfsv,cv,lop,mb), gridoperator.hh:76-89 — builds the DOF
 * numbering (ordering/leafgridviewordering.hh:120-197), the constraint set
 * (constraints/conforming.hh:53-93, constraints/p0.hh:31-41) and uploads the coefficient fields.
 * All arrays in *p may be host or device memory; they are copied. */
int pdb200_create(const pdb200_problem* p, pdb200_handle* out);
int pdb200_destroy(pdb200_handle h);

/* re-upload coefficient arrays (same shapes); NULL members of *p are left unchanged */
int pdb200_update_coefficients(pdb200_handle h, const pdb200_problem* p);

/* GridOperator::globalSizeU()/globalSizeV(), gridoperator.hh:104-113 */
int pdb200_num_dofs(pdb200_handle h, uint64_t* n);
int pdb200_local_size(pdb200_handle h, uint32_t* n);          /* DOFs per cell, (k+1)^dim      */
int pdb200_num_boundary_faces(pdb200_handle h, uint64_t* n);
int pdb200_boundary_face_offset(pdb200_handle h, int dir, int side, uint64_t* first);
int pdb200_quadrature_size(pdb200_handle h, uint32_t* m);     /* Gauss points per direction    */
/* ascending Gauss-Legendre abscissae on [0,1] and weights, m entries each (host arrays) */
int pdb200_quadrature(pdb200_handle h, double* points, double* weights);

/* The same rule without an operator handle (pure host table; no device needed): callers use it to
 * sample f, g, j, o at the points the reference evaluates them (common/quadraturerules.hh:117-120).
 * m = (2*degree + intorderadd)/2 + 1; ascending abscissae on [0,1]. */
int pdb200_gauss_legendre(int m, double* points, double* weights);

/* container index of local DOF i of cell `cell`: LFSIndexCache::containerIndex,
 * gridfunctionspace/lfsindexcache.hh:603-633 + ordering/leaforderingbase.hh:97-203.
 * Writes (k+1)^dim entries (host array). */
int pdb200_cell_dof_indices(pdb200_handle h, uint64_t cell, uint64_t* idx);

/* constrained DOFs (constraints(bctype,gfs,cc), constraints/common/constraints.hh:588-687).
 * Two-call protocol: idx == NULL returns the count. Sorted ascending. */
int pdb200_constrained_dofs(pdb200_handle h, uint64_t* count, uint64_t* idx);

/* GridOperator::residual(x, r), gridoperator.hh:176-181:   r += R(x), constrained rows := 0 */
int pdb200_residual(pdb200_handle h, const double* x, double* r);

/* GridOperator::jacobian_apply(z, y) (linear variant), gridoperator.hh:192-197:  y += J z,
 * constrained rows := 0 (jacobianapplyengine.hh:249-254) */
int pdb200_jacobian_apply(pdb200_handle h, const double* z, double* y);

/* OnTheFlyOperator::apply(x, y), backend/istl/seqistlsolverbackend.hh:66-76:  y = 0; y += J x.
 * The zeroing is fused into the kernel (y is written, never read): this is the call a Krylov
 * solver makes once per iteration and the one bench.py times. */
int pdb200_onthefly_apply(pdb200_handle h, const double* x, double* y);

/* GridOperator::jacobian_apply(u, z, y) (non-linear variant), gridoperator.hh:200-205.  Both
 * supported local operators are linear, so like the reference this always fails (:202-203). */
int pdb200_jacobian_apply_nonlinear(pdb200_handle h, const double* u, const double* z, double* y);

/* BCRSMatrixBackend::buildPattern -> GridOperator::fill_pattern(p), gridoperator.hh:168-173,
 * backend/istl/bcrsmatrixbackend.hh:90-121,225-241.  Scalar CSR over DOFs, column indices
 * ascending inside a row (dune-istl setIndices).  size_t-compatible output (host or device). */
int pdb200_pattern_size(pdb200_handle h, uint64_t* nrows, uint64_t* nnz);
int pdb200_pattern(pdb200_handle h, uint64_t* rowptr, uint64_t* colidx);
/* same pattern with 32-bit column indices (device-native layout) */
int pdb200_pattern_i32(pdb200_handle h, uint64_t* rowptr, uint32_t* colidx);

/* DG with ISTL::VectorBackend<Blocking::fixed,n>: block CSR over cells, blocks are row-major
 * FieldMatrix<double,n,n> stored contiguously (backend/common/aliasedmatrixview.hh:93-97). */
int pdb200_block_pattern_size(pdb200_handle h, uint64_t* nblockrows, uint64_t* nblocks);
int pdb200_block_pattern(pdb200_handle h, uint64_t* rowptr, uint64_t* colidx);

/* GridOperator::jacobian(x, A), gridoperator.hh:184-189: values += dR/dx in the order of
 * pdb200_pattern (layout 0) or pdb200_block_pattern (layout 1); afterwards constrained rows are
 * cleared and get a unit diagonal (assemblerutilities.hh:666-684, bcrsmatrix.hh:254-258). */
enum { PDB200_LAYOUT_CSR = 0, PDB200_LAYOUT_BCSR = 1 };
int pdb200_jacobian(pdb200_handle h, const double* x, double* values, int layout);
/* `*A = 0.0; go.jacobian(x, *A);` as issued by StationaryLinearProblemSolver::apply
 * (stationary/linearproblem.hh:221-226) with the zeroing fused: values are written, never read. */
int pdb200_jacobian_fresh(pdb200_handle h, const double* x, double* values, int layout);

/* y = A x with the assembled matrix (dune-istl BCRSMatrix::mv as used by
 * backend/istl/seqistlsolverbackend.hh MatrixAdapter) — used to check jacobian against
 * jacobian_apply on the device. */
int pdb200_csr_mv(pdb200_handle h, const double* values, int layout, const double* x, double* y);

/* ---- linear solvers on the device (the consumers of jacobian_apply / jacobian) ---------------
 * pdb200_solve: the reference's sequential ISTL back-ends with every vector resident on the GPU.
 *   values == NULL : matrix-free, ISTLBackend_SEQ_MatrixFree_BCGS_Richardson
 *                    (backend/istl/seqistlsolverbackend.hh:157-203,1039-1050): OnTheFlyOperator (:44-100)
 *                    + Richardson(1.0) + BiCGSTAB (or CG), precond PDB200_PRECOND_NONE; or
 *                    ISTLBackend_SEQ_MatrixFree_Base (backend/istl/matrixfree/backends.hh:62-143) with the exact
 *                    block-Jacobi preconditioner (AssembledBlockJacobiPreconditionerLocalOperator,
 *                    backend/istl/matrixfree/assembledblockjacobipreconditioner.hh:96-230), precond
 *                    PDB200_PRECOND_BLOCK_JACOBI (QkDG, k = 1, 2, SIPG, diagonal A, b = 0); or point Jacobi on the
 *                    matrix-free diagonal (pdb200_point_diagonal), precond PDB200_PRECOND_JACOBI
 *   values != NULL : assembled matrix in `layout` (MatrixAdapter), ISTLBackend_SEQ_BCGS_Jac / _CG_Jac
 *                    (:401-416,538-553; SeqJac, one step, w = 1; scalar CSR layout only) or no
 *                    preconditioner
 * z: initial guess in, solution out;  r: right-hand side in, final defect out (dune-istl overwrites
 * the right-hand side).  Host or device pointers.  Stopping rule of dune-istl:
 * |defect| < reduction * |defect_0|  (two-norm, SequentialNorm).  Not converged within maxiter is
 * reported through res->converged == 0, not as an error (LinearSolverResult, backend/solver.hh:28-51). */
enum { PDB200_SOLVER_BICGSTAB = 0, PDB200_SOLVER_CG = 1 };
enum { PDB200_PRECOND_NONE = 0, PDB200_PRECOND_JACOBI = 1, PDB200_PRECOND_BLOCK_JACOBI = 2,
       PDB200_PRECOND_BLOCK_SOR = 3,   /* one forward block SOR sweep from v = 0 (matrix-free, BiCGSTAB)           */
       PDB200_PRECOND_BLOCK_SSOR = 4   /* forward + backward sweep: the symmetric variant a CG needs                */ };
typedef struct pdb200_solve_result {
  int32_t converged;
  uint32_t iterations;
  double elapsed;       /* seconds, wall clock of the call */
  double reduction;     /* achieved defect reduction */
  double conv_rate;     /* reduction^(1/iterations) */
  double first_defect;  /* |b - A z_0| */
  double defect;        /* final defect norm */
} pdb200_solve_result;
int pdb200_solve(pdb200_handle h, int solver, int precond, const double* values, int layout, double* z, double* r,
                 double reduction, uint32_t maxiter, pdb200_solve_result* res);
/* z = D^-1 r with D the block diagonal of the QkDG Jacobian (one n x n block per cell), the operation of
 * AssembledBlockJacobiPreconditionerLocalOperator::jacobian_apply_volume
 * (backend/istl/matrixfree/assembledblockjacobipreconditioner.hh:189-230) wrapped by
 * GridOperatorPreconditioner::apply (gridoperatorpreconditioner.hh:81-87).  Matrix-free: the blocks are
 * Kronecker sums and are inverted by fast diagonalisation (csrc/dg_blockjac.cu).  Host or device pointers. */
int pdb200_block_jacobi_apply(pdb200_handle h, const double* r, double* z);

/* y = D z with D the block diagonal of the QkDG Jacobian — BlockDiagonalLocalOperatorWrapper::jacobian_apply_volume /
 * _skeleton / _boundary (localoperator/blockdiagonalwrapper.hh:100-300) — and y = (J - D) z, the block off-diagonal
 * part — BlockOffDiagonalLocalOperatorWrapper (localoperator/blockoffdiagonalwrapper.hh:60-240).  Matrix-free from the
 * same per-cell fast-diagonalisation data as pdb200_block_jacobi_apply.  Host or device pointers; y is overwritten. */
int pdb200_block_diagonal_apply(pdb200_handle h, const double* z, double* y);
int pdb200_block_offdiagonal_apply(pdb200_handle h, const double* z, double* y);

/* One block SOR sweep, matrix-free — BlockSORPreconditionerLocalOperator (backend/istl/matrixfree/
 * blocksorpreconditioner.hh:36-301) wrapped by GridOperatorPreconditioner::apply(v, d):
 *   for every cell T_i in index-set order:  a_i = d_i - sum_{j != i} A_ij v_j;  solve D_i b_i = a_i;
 *                                           v_i = (1 - omega) v_i + omega b_i            (in place).
 * The sweep runs over the hyperplanes i + j + k = const of the structured grid (cells of one hyperplane are mutually
 * independent, their lower neighbours are already updated): exactly the reference's sequential order.
 * flags: PDB200_SOR_BACKWARD sweeps in reverse order; PDB200_SOR_KEEP_ITERATE starts from the v passed in instead of
 * v = 0 (smoother use).  QkDG, k = 1, 2, SIPG, diagonal A, b = 0.  Host or device pointers. */
enum { PDB200_SOR_BACKWARD = 1, PDB200_SOR_KEEP_ITERATE = 2 };
int pdb200_block_sor_apply(pdb200_handle h, const double* d, double* v, double omega, int flags);
/* omega of PDB200_PRECOND_BLOCK_SOR / _SSOR inside pdb200_solve (default 1.0) */
int pdb200_set_relaxation(pdb200_handle h, double omega);

/* d = point diagonal of the Jacobian, matrix-free (PointDiagonalLocalOperatorWrapper,
 * localoperator/pointdiagonalwrapper.hh: the diagonal a point-Jacobi preconditioner needs without an assembled
 * matrix); constrained rows carry 1 like the unit rows of the assembled matrix.  Conforming Qk and QkDG (k = 1, 2),
 * diagonal A, b = 0.  Host or device pointer.  pdb200_solve with values == NULL and PDB200_PRECOND_JACOBI uses it. */
int pdb200_point_diagonal(pdb200_handle h, double* d);

/* StationaryLinearProblemSolver::apply (stationary/linearproblem.hh:188-302):
 *   [A = 0; jacobian(x, A)]  r = 0; residual(x, r);  red = max(reduction, min_defect / |r|);
 *   solve J z = r to red;  x -= z.
 * matrix_free != 0 skips the assembly (linearSolverIsMatrixFree).  x: host or device pointer. */
int pdb200_solve_stationary(pdb200_handle h, int solver, int precond, int matrix_free, double* x, double reduction,
                            double min_defect, uint32_t maxiter, pdb200_solve_result* res);

/* ---- OneStepGridOperator: the operator of one Runge-Kutta / fractional-step stage ----------------------------
 * gridoperator/onestep.hh:30-308 with the engines of gridoperator/onestep/ (localassembler.hh, prestageengine.hh,
 * residualengine.hh, jacobianengine.hh, jacobianapplyengine.hh).  go0 is the spatial operator, go1 the temporal
 * one (L2 mass operator, localoperator/l2.hh: a handle with A = 0, c = scaling, boundary type None).  For stage r of
 * a method with coefficients a_ri, b_ri, d_i (instationary/onestepparameter.hh:43-84; r in 1..s, i in 0..r):
 *     preStage : const_residual = sum_{i<r}  b_ri dt_factor0 R0(x_i) [if |b_ri| > 1e-6] + a_ri dt_factor1 R1(x_i) [if |a_ri| > 1e-6]
 *     residual : r += b_rr dt_factor0 R0(x) + dt_factor1 R1(x) + const_residual, constrained rows := 0
 *     jacobian / jacobian_apply : the same weights on the two Jacobians
 * On the device a stage is ONE fused operator: the weighted sum of two convection-diffusion-reaction forms is the
 * form with the weighted coefficient fields (csrc/onestep.cu), so every call costs one pass of the same kernels as
 * the stationary operator.  pdb200_onestep_stage_operator hands that fused operator out as an ordinary handle:
 * pdb200_solve, pdb200_pattern, pdb200_block_jacobi_apply, ... work on it unchanged.
 * The two operator handles are referenced, not owned (like the reference, onestep.hh:300-305); re-sample
 * time-dependent coefficients with pdb200_update_coefficients(go0, ...) before the call that evaluates them. */
typedef struct pdb200_onestep* pdb200_onestep_handle;
/* OneStepLocalAssembler::DTAssemblingMode, gridoperator/onestep/localassembler.hh:140 */
enum { PDB200_ONESTEP_DIVIDE_OPERATOR1_BY_DT = 0, PDB200_ONESTEP_MULTIPLY_OPERATOR0_BY_DT = 1,
       PDB200_ONESTEP_DO_NOT_ASSEMBLE_DT = 2 };
/* OneStepGridOperator(go0, go1), onestep.hh:66-76 */
int pdb200_onestep_create(pdb200_handle go0, pdb200_handle go1, pdb200_onestep_handle* out);
int pdb200_onestep_destroy(pdb200_onestep_handle os);
/* setMethod, onestep.hh:245-248: a, b are s x (s+1) row-major (row r-1 holds a(r, 0..s)), d has s+1 entries;
 * `implicit` = method.implicit() (explicit methods switch to DoNotAssembleDT, onestep.hh:74-75) */
int pdb200_onestep_set_method(pdb200_onestep_handle os, int s, const double* a, const double* b, const double* d,
                              int implicit);
/* divideMassTermByDeltaT / multiplySpatialTermByDeltaT, onestep.hh:78-91 (takes effect at the next preStep) */
int pdb200_onestep_set_dt_mode(pdb200_onestep_handle os, int mode);
/* preStep(method, time, dt), onestep.hh:250-254 + localassembler.hh:101-130 */
int pdb200_onestep_pre_step(pdb200_onestep_handle os, double time, double dt);
/* LocalAssembler::timeAtStage(stage), localassembler.hh:148-151 */
int pdb200_onestep_time_at_stage(pdb200_onestep_handle os, int stage, double* t);
/* preStage(stage, x), onestep.hh:130-139: x[0..stage-1] are the solutions of the earlier stages (host or device).
 * begin/add is the same call split per earlier stage, so that a host with time-dependent coefficients can
 * re-sample them at t + d_i dt before add(i) (prestageengine.hh:208-211 sets that time per stage). */
int pdb200_onestep_pre_stage(pdb200_onestep_handle os, int stage, const double* const* x);
int pdb200_onestep_pre_stage_begin(pdb200_onestep_handle os, int stage);
int pdb200_onestep_pre_stage_add(pdb200_onestep_handle os, int i, const double* x);
/* One stage of an EXPLICIT method — ExplicitOneStepMethod::apply (instationary/explicitonestep.hh:332-414) with
 * OneStepGridOperator::explicit_jacobian_residual (onestep.hh:161-178):
 *   x_r = -M^-1 sum_{i<r} ( a_ri M x_i + b_ri dt R0(x_i; t + d_i dt) ).
 * QkDG spaces (block-diagonal mass matrix): exact block inverse for k <= 2, CG on the mass operator to `reduction`
 * otherwise.  x[0..stage-1]: earlier stages, xr: result (host or device).  The time-step controller of the reference
 * (CFL limit) is the caller's business: dt is the one given to pdb200_onestep_pre_step. */
int pdb200_onestep_explicit_stage(pdb200_onestep_handle os, int stage, const double* const* x, double* xr, double reduction);
/* The same stage split per earlier stage (begin, add(0) .. add(stage-1), finish = the mass solve), so that a host with
 * time-dependent coefficients can re-sample them at t + d_i dt before add(i): the explicit engine delegates to the
 * pre-stage engine, which sets the time of every earlier stage (onestep/jacobianresidualengine.hh,
 * prestageengine.hh:208-211). */
int pdb200_onestep_explicit_stage_begin(pdb200_onestep_handle os, int stage);
int pdb200_onestep_explicit_stage_add(pdb200_onestep_handle os, int i, const double* x);
int pdb200_onestep_explicit_stage_finish(pdb200_onestep_handle os, double* xr, double reduction);
/* copy of the constant part of the residual assembled by preStage (host or device destination) */
int pdb200_onestep_const_residual(pdb200_onestep_handle os, double* out);
/* residual(x, r), onestep.hh:141-149 */
int pdb200_onestep_residual(pdb200_onestep_handle os, const double* x, double* r);
/* jacobian_apply(update, result), onestep.hh:180-185: y += (b_rr dt_factor0 J0 + dt_factor1 J1) z */
int pdb200_onestep_jacobian_apply(pdb200_onestep_handle os, const double* z, double* y);
/* the same with the zeroing fused (OnTheFlyOperator::apply on the one-step operator) */
int pdb200_onestep_onthefly_apply(pdb200_onestep_handle os, const double* x, double* y);
/* jacobian(x, a), onestep.hh:151-159: values += b_rr dt_factor0 dR0/dx + dt_factor1 dR1/dx on go0's pattern */
int pdb200_onestep_jacobian(pdb200_onestep_handle os, const double* x, double* values, int layout);
/* the fused operator of the current stage (owned by `os`; valid until the weights or coefficients change) */
int pdb200_onestep_stage_operator(pdb200_onestep_handle os, pdb200_handle* stage);
/* StationaryLinearProblemSolver::apply (stationary/linearproblem.hh:188-302) on the one-step operator — the stage
 * solver of OneStepMethod::apply (instationary/implicitonestep.hh:191) for linear problems, device-resident:
 *   [A = 0; jacobian(x, A)]  r = 0; residual(x, r);  solve J z = r to max(reduction, min_defect / |r|);  x -= z.
 * Arguments as pdb200_solve_stationary. */
int pdb200_onestep_solve_stationary(pdb200_onestep_handle os, int solver, int precond, int matrix_free, double* x,
                                    double reduction, double min_defect, uint32_t maxiter, pdb200_solve_result* res);
int pdb200_onestep_launch_count(pdb200_onestep_handle os, uint64_t* n);

/* Halo exchange support for the overlapping partition (replaces the AddDataHandle/CopyDataHandle
 * communication of boilerplate/pdelab.hh:872-880 + gridfunctionspace/genericdatahandle.hh).
 * pack copies the DOFs of the owned cell layer next to side (dir,side) — the layer at distance
 * `overlap` from the box face — into a contiguous DEVICE buffer; unpack writes a received buffer
 * into the ghost layer of that side.  The transport (NCCL send/recv) is done by the caller. */
int pdb200_halo_layer_size(pdb200_handle h, int dir, uint64_t* ndoubles);
int pdb200_halo_pack(pdb200_handle h, const double* x, int dir, int side, double* buf);
int pdb200_halo_unpack(pdb200_handle h, double* x, int dir, int side, const double* buf);

/* Index gather / scatter on DEVICE vectors: buf[i] = x[idx[i]] / x[idx[i]] = buf[i] (idx: int64 device
 * array).  The pack / unpack of the conforming-Qk ghost exchange, whose lattice planes are spread over
 * the sub-entity groups of the container (python/pdelab_b200/partition.py: QkHaloExchanger). */
int pdb200_gather_dofs(pdb200_handle h, const double* x, const int64_t* idx, uint64_t n, double* buf);
int pdb200_scatter_dofs(pdb200_handle h, const double* buf, const int64_t* idx, uint64_t n, double* x);

/* Parts of the local box for overlapping communication with computation: INTERIOR = every tile
 * of cells that reads no ghost layer (can run while the halo exchange is in flight), BOUNDARY =
 * the rest.  INTERIOR followed by BOUNDARY equals ALL.  Device pointers only, y is overwritten. */
enum { PDB200_PART_ALL = 0, PDB200_PART_INTERIOR = 1, PDB200_PART_BOUNDARY = 2 };
int pdb200_onthefly_apply_part(pdb200_handle h, const double* x, double* y, int part);

/* Peer-to-peer halo exchange over NVLink / NVSwitch without NCCL or host round trips (one process
 * per GPU, CUDA IPC).  Every rank creates a MAILBOX (receive buffers for its processor sides +
 * ready/ack flags) and hands the returned handle to its face neighbours out of band (the Python
 * host uses one torch.distributed all_gather at set-up); each processor side is then connected
 * to the neighbour's mailbox.  Replaces the same CopyDataHandle communication as the pack/unpack
 * entry points above. */
typedef struct pdb200_ipc_handle { unsigned char bytes[64]; } pdb200_ipc_handle;
int pdb200_halo_p2p_create(pdb200_handle h, pdb200_ipc_handle* mine);
int pdb200_halo_p2p_connect(pdb200_handle h, int dir, int side, const pdb200_ipc_handle* neighbour);
/* owner -> ghost copy of the DEVICE vector x across all connected sides (two kernels: push into
 * the neighbours' mailboxes, wait + unpack); asynchronous on the handle's stream */
int pdb200_halo_exchange_p2p(pdb200_handle h, double* x);
/* y = J x on the overlapping partition with the exchange hidden behind the interior tiles.
 * QkDG k = 2 in 3-D whose owned extents in the split directions are multiples of the 8x4x4 tile: ONE launch —
 * push blocks (remote stores into the neighbours' mailboxes), interior tiles, then the tiles next to a processor
 * side, which wait for the neighbour's layer and read it straight from the mailbox; x's ghost layers are neither
 * read nor written.  Otherwise: side stream: push, wait + unpack, BOUNDARY;  main stream: INTERIOR, (join), and
 * x's ghost layers are updated as a side effect.  Callers must not rely on x's ghost layers afterwards (use
 * pdb200_halo_exchange_p2p for a consistent vector).  Device pointers only. */
int pdb200_onthefly_apply_p2p(pdb200_handle h, double* x, double* y);

/* ---- overlapping solvers on several GPUs --------------------------------------------------------------------
 * OverlappingOperator, OverlappingScalarProduct and the Krylov loop of the ISTLBackend_OVLP_* back-ends
 * (backend/istl/ovlpistlsolverbackend.hh:30-134, 477-560) with one process per GPU, device-resident:
 *   apply : owner -> ghost copy of the input over the NVLink mailboxes (hidden behind the interior tiles), local rows,
 *           ghost rows := 0 (set_constrained_dofs(cc, 0.0, y), :48-49)
 *   dot   : disjoint inner product + sum over all ranks (:103-108) — the sum runs through a second set of
 *           peer-mapped mailboxes (one kernel, no NCCL call, no host round trip; bit-identical on every rank)
 * Set-up: every rank creates its reduction mailbox and connects every other rank's (the handles travel out of
 * band like the halo mailboxes').  At most 16 ranks. */
int pdb200_comm_create(pdb200_handle h, int rank, int size, pdb200_ipc_handle* mine);
int pdb200_comm_connect(pdb200_handle h, int peer_rank, const pdb200_ipc_handle* peer);
/* gridView().comm().sum of one or two doubles (host or device pointer), collective over all ranks */
int pdb200_comm_sum(pdb200_handle h, double* values, int count);
/* pdb200_solve on the overlapping partition (QkDG; needs the halo and the reduction mailboxes).  values == NULL:
 * matrix-free operator; otherwise the local assembled matrix of the rank.  precond as pdb200_solve (block / point
 * Jacobi are local to a cell / a row: no extra communication).  z, r: DEVICE vectors over the local box; the ghost
 * rows of r are ignored, z is consistent (ghost layers filled) on return.  Collective: every rank calls it. */
int pdb200_solve_ovlp(pdb200_handle h, int solver, int precond, const double* values, int layout, double* z, double* r,
                      double reduction, uint32_t maxiter, pdb200_solve_result* res);

/* stream control for device-pointer calls: `stream` is a cudaStream_t */
int pdb200_set_stream(pdb200_handle h, void* stream);
int pdb200_synchronize(pdb200_handle h);

/* number of kernel launches issued by this handle so far (bench.py "gpu_launches") */
int pdb200_launch_count(pdb200_handle h, uint64_t* n);
/* name of the kernel variant the last jacobian_apply / residual used ("dg_fast_q2_3d", ...) */
const char* pdb200_last_kernel(pdb200_handle h);

/* library / build identification */
const char* pdb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PDELAB_B200_H */

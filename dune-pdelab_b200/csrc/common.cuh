// common.cuh — shared definitions of the sm_100a operator-evaluation kernels.
//
// Reference paths in comments are relative to /root/reference/dune/pdelab/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../../include/pdelab_b200.h"

namespace pdb {

constexpr int MAX_K = 4;            // highest polynomial degree with a compiled kernel
constexpr int MAX_N1 = MAX_K + 1;   // 1-D basis functions
constexpr int MAX_M = MAX_K + 1;    // Gauss points per direction (intorderadd in {0,1})
constexpr int MAX_PTS = MAX_M + 2;  // Gauss points, then xi = 0, then xi = 1

// Everything a kernel needs, passed by value (lives in the constant bank).
struct DevParams {
  int dim, k, n1, n, m, nq, nfq;
  int N[3];
  long long ncells, ndofs;
  double h[3], ih[3], area[3], vol;
  double theta, alpha;
  int weights_on, a_mode, dg;
  int basis;  // PDB200_BASIS_* of the QkDG space (the Kronecker kernels are written for the Lagrange basis)
  int pw;     // PDB200_POINTWISE_* bits: A / b / c / bctype in the point-wise layout of pdelab_b200.h
  int np;     // sample points per cell of that layout: nq + 2 dim nfq
  int side_kind[3][2];
  long long bf_off[3][2];
  const double *A, *b, *c, *f, *g, *j, *o;
  const int8_t* bctype;
  // 1-D tables: P[pt][i] = p_i(x_pt), DP[pt][i] = p_i'(x_pt)  (finiteelement/qkdglagrange.hh:55-79)
  double P[MAX_PTS * MAX_N1], DP[MAX_PTS * MAX_N1], wq[MAX_M];
};

// exactly integrated 1-D matrices of the Lagrange basis on [0,1] used by the Kronecker kernels
struct Kron1D {
  double M[MAX_N1 * MAX_N1];      // mass
  double Minv[MAX_N1 * MAX_N1];
  double d0[MAX_N1], d1[MAX_N1];  // p_i'(0), p_i'(1)
  // t_i += E0[i]*(a u'(0)) + E1[i]*(a u'(1)) + m0[i]*FL + mk[i]*FR + q0[i]*GL + q1[i]*GR   (k = 2)
  double E0[MAX_N1], E1[MAX_N1], m0[MAX_N1], mk[MAX_N1], q0[MAX_N1], q1[MAX_N1];
  double MinvK[MAX_N1 * MAX_N1];  // Minv * stiffness
  double Dn[MAX_N1 * MAX_N1];     // nodal derivative matrix: Dn[i][j] = p_j'(i / k)
};

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define PDB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      throw pdb::Error(std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
  } while (0)

__host__ __device__ inline long long cell_index(const int N[3], int x, int y, int z) {
  return x + (long long)N[0] * (y + (long long)N[1] * z);
}

// boundary-face number of face (dir, side) of the cell (x,y,z); numbering of pdelab_b200.h
__host__ __device__ inline long long bface_index(const DevParams& P, const int c[3], int dir, int side) {
  long long idx = 0, stride = 1;
  for (int d = 0; d < P.dim; d++)
    if (d != dir) {
      idx += stride * c[d];
      stride *= P.N[d];
    }
  return P.bf_off[dir][side] + idx;
}

// entry e of the tensor array: e = cell (cell-wise layout) or cell * np + pt (point-wise layout)
__device__ inline void load_A(const DevParams& P, long long e, double A[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[i][j] = 0.0;
  switch (P.a_mode) {
    case PDB200_A_IDENTITY:
      for (int i = 0; i < P.dim; i++) A[i][i] = 1.0;
      break;
    case PDB200_A_SCALAR: {
      double v = __ldg(P.A + e);
      for (int i = 0; i < P.dim; i++) A[i][i] = v;
    } break;
    case PDB200_A_DIAGONAL:
      for (int i = 0; i < P.dim; i++) A[i][i] = __ldg(P.A + e * P.dim + i);
      break;
    default:
      for (int i = 0; i < P.dim; i++)
        for (int j = 0; j < P.dim; j++) A[i][j] = __ldg(P.A + e * P.dim * P.dim + i * P.dim + j);
  }
}

// ---- the coefficient call-backs of the reference's parameter class, in either layout of pdelab_b200.h ----------
// sample-point number of face quadrature point q of face (dir, side) inside its cell
__device__ __forceinline__ int face_pt(const DevParams& P, int dir, int side, int q) {
  return P.nq + (2 * dir + side) * P.nfq + q;
}
__device__ __forceinline__ bool pw_A(const DevParams& P) { return (P.pw & PDB200_POINTWISE_A) != 0; }
// param.A(cell, x_pt), point-wise layout only (permeabilityIsConstantPerCell() == false)
__device__ inline void load_A_at(const DevParams& P, long long cell, int pt, double A[3][3]) {
  load_A(P, cell * P.np + pt, A);
}
// param.A(cell, centre); in the point-wise layout every use is preceded by load_A_at, sample 0 stands in
__device__ inline void load_A_cell(const DevParams& P, long long cell, double A[3][3]) {
  load_A(P, pw_A(P) ? cell * P.np : cell, A);
}
// param.b(cell, x_pt)
__device__ inline void load_b(const DevParams& P, long long cell, int pt, double b[3]) {
  b[0] = b[1] = b[2] = 0.0;
  if (!P.b) return;
  const long long e = (P.pw & PDB200_POINTWISE_B) ? cell * P.np + pt : cell;
  for (int d = 0; d < P.dim; d++) b[d] = __ldg(P.b + e * P.dim + d);
}
// param.c(cell, x_q) at volume point q
__device__ __forceinline__ double load_c(const DevParams& P, long long cell, int q) {
  if (!P.c) return 0.0;
  return __ldg(P.c + ((P.pw & PDB200_POINTWISE_C) ? cell * P.nq + q : cell));
}
// param.bctype(intersection, x_q): per face or per face quadrature point (QkDG, convectiondiffusiondg.hh:763)
__device__ __forceinline__ int load_bctype(const DevParams& P, long long bf, int q) {
  if (!P.bctype) return (int)PDB200_BC_DIRICHLET;
  return (int)P.bctype[(P.pw & PDB200_POINTWISE_BCTYPE) ? bf * P.nfq + q : bf];
}

// weightsOff penalty of the Kronecker kernels: alpha/h_F k(k+d-1) on every interior and Dirichlet face REGARDLESS of A
// (harmonic_average = 1, convectiondiffusiondg.hh:334-338, 724-727), none on faces without a u-dependent term.  The face
// set-up marks the latter by cs = -0.0 (cs >= +0 everywhere else), so that a cell with A == 0 keeps its penalty.
#ifdef __CUDACC__
__device__ __forceinline__ bool face_has_penalty(double cs) { return __double2hiint(cs) >= 0; }
#endif

// The Kronecker fast paths need cell-wise constant DIAGONAL A, no convection and per-face boundary types; anything
// sampled per quadrature point (P.pw) runs through the reference-order kernels.
inline bool kron_coefficients(const DevParams& P) { return P.pw == 0 && P.b == nullptr && P.a_mode != PDB200_A_FULL; }

// ---- launchers implemented in the .cu files ------------------------------------------------

struct Operator;  // operator.cu

// dg_generic.cu: reference-order gather kernels, any supported (dim,k), any coefficient mode
void launch_dg_generic(const DevParams& P, const double* x, double* y, bool residual, bool overwrite,
                       int* errflag, cudaStream_t s);
// dg_fast.cu: Kronecker-factorised kernel (k = 2, dim = 3, diagonal A, cell-wise constant b and c)
bool dg_fast_supported(const DevParams& P);
struct FastPlan;
FastPlan* dg_fast_plan_create(const DevParams& P, const Kron1D& K);
void dg_fast_plan_destroy(FastPlan*);
// part: PDB200_PART_*; returns the number of kernel launches
// r0 != nullptr: residual form  y += J x + r0  with r0 = R(0) (the operator is affine)
// [ztile_lo, ztile_hi): optional window of tile layers along z (PART_ALL only), used to pipeline
// host transfers with the computation
int launch_dg_fast(FastPlan* plan, const DevParams& P, const double* x, double* y, const double* r0,
                   bool overwrite, int part, cudaStream_t s, int* errflag, int ztile_lo = 0, int ztile_hi = 1 << 30);
int dg_fast_ztiles(const DevParams& P);
void dg_fast_ztile_layers(const DevParams& P, int lo, int hi, int* z0, int* z1);


// dg_kron.cu: Kronecker-factorised kernel for higher degree (k = 3, 4; dim = 3, diagonal A, b = 0)
bool dg_kron_supported(const DevParams& P);
struct KronPlan;
KronPlan* dg_kron_plan_create(const DevParams& P, const Kron1D& K);
void dg_kron_plan_destroy(KronPlan*);
int launch_dg_kron(KronPlan* plan, const DevParams& P, const double* x, double* y, const double* r0, bool overwrite,
                   cudaStream_t s);

// dg_small.cu: Kronecker-factorised kernel for small cells (dim = 2 with k = 1, 2; dim = 3 with k = 1), thread per cell;
// cell-wise constant diagonal A, b and c
bool dg_small_supported(const DevParams& P);
int launch_dg_small(const DevParams& P, const Kron1D& K, const double* x, double* y, const double* r0, bool overwrite,
                    cudaStream_t s, int* errflag = nullptr);

// dg_blockjac.cu: exact matrix-free block-Jacobi preconditioner z = D^-1 r (fast diagonalisation of the Kronecker-sum blocks)
struct BlockJacPlan;
bool dg_blockjac_supported(const DevParams& P);
BlockJacPlan* dg_blockjac_create(const DevParams& P, const Kron1D& K);
void dg_blockjac_destroy(BlockJacPlan*);
void dg_blockjac_invalidate(BlockJacPlan*);  // coefficients changed
int launch_dg_blockjac(BlockJacPlan*, const DevParams& P, const Kron1D& K, const double* r, double* z, cudaStream_t s);
// y = D z (mode 1) or y -= D z (mode 2) with the same per-cell data
int launch_dg_blockdiag(BlockJacPlan*, const DevParams& P, const Kron1D& K, const double* z, double* y, int mode, cudaStream_t s);
// one block SOR sweep over hyperplane wavefronts (lexicographic = index-set order), in place
int launch_dg_blocksor(BlockJacPlan*, const DevParams& P, const Kron1D& K, const double* d, double* v, double omega,
                       bool backward, bool zero_start, cudaStream_t s);

// halo.cu: pack / unpack one cell layer of a DG vector
void launch_halo_copy(const DevParams& P, double* x, double* buf, int dir, int side, bool pack, cudaStream_t s);
void launch_gather(const double* x, const long long* idx, long long n, double* buf, cudaStream_t s);
void launch_scatter(const double* buf, const long long* idx, long long n, double* x, cudaStream_t s);
// halo.cu: peer-to-peer mailbox exchange over NVLink (CUDA IPC), see pdelab_b200.h
struct P2PHalo;
P2PHalo* p2p_create(const DevParams& P, pdb200_ipc_handle* mine);
void p2p_connect(P2PHalo*, const DevParams& P, int dir, int side, const pdb200_ipc_handle* peer);
void p2p_destroy(P2PHalo*);
int p2p_push(P2PHalo*, const DevParams& P, const double* x, cudaStream_t s);       // returns launches
int p2p_wait_unpack(P2PHalo*, const DevParams& P, double* x, cudaStream_t s);
// both spaces: owner -> ghost copy (QkDG: push + wait_unpack; conforming Qk: lattice planes, direction by direction)
int p2p_exchange(P2PHalo*, const DevParams& P, double* x, cudaStream_t s);
int p2p_zero_ghosts(P2PHalo*, const DevParams& P, double* x, cudaStream_t s);  // x := 0 on everything not owned
bool p2p_is_qk(const P2PHalo*);
void p2p_check(P2PHalo*);  // throws if a spin-wait timed out
cudaStream_t p2p_stream(P2PHalo*);
cudaEvent_t p2p_event(P2PHalo*, int i);

// One-launch step of the overlapping partition (dg_fast.cu runs it, halo.cu owns the mailboxes): the table lives in
// device memory; side index = 2 * dir + side.  Flags and epochs are the ones of p2p_push / p2p_wait_unpack, so fused
// and unfused steps can alternate on the same mailboxes.
struct FusedSide {
  int active;
  long long total2, chunk2, stride2, src_off2;  // the layer sent across this side, in 16-byte elements (see copy_layer)
  double2* peer_buf;                            // receive buffer in the neighbour's mailbox
  unsigned long long *peer_ready, *peer_ack;
  const double* my_buf;                         // my receive buffer for this side: the neighbour's boundary layer
  unsigned long long *my_ready, *my_ack;
};
struct FusedTable {
  FusedSide s[6];
  unsigned int* counters;  // [12]: push blocks done per side, tiles that have consumed a side's buffer
  int* err;                // spin-wait time-out flag (p2p_check)
};
// device table once every processor side is connected and every layer moves in 16-byte elements, else nullptr
const FusedTable* p2p_fused_table(P2PHalo*, const DevParams& P);
unsigned long long p2p_next_epoch(P2PHalo*);
// dg_fast.cu: y = J x with the ghost exchange inside the launch (push blocks, interior tiles, then the tiles that
// read a neighbour's layer straight from the mailbox); x's ghost layers are neither read nor written
bool dg_fast_fused_supported(const DevParams& P);
int launch_dg_fast_fused(FastPlan* plan, const DevParams& P, const double* x, double* y, const FusedTable* table,
                         const FusedTable& table_host, unsigned long long epoch, cudaStream_t s, int* errflag);
const FusedTable& p2p_fused_table_host(P2PHalo*);

// halo.cu: all-ranks reduction of inner-product partials over peer-mapped mailboxes (overlapping solvers)
struct PeerComm;
PeerComm* comm_create(int rank, int size, pdb200_ipc_handle* mine);
void comm_connect(PeerComm*, int peer_rank, const pdb200_ipc_handle* peer);
void comm_destroy(PeerComm*);
int comm_size(const PeerComm*);
// in place: P1[0] = sum over ranks of sum(P1[0..nb)), P1[1..nb) = 0 (the same for P2 if given); one launch
void comm_allreduce_partials(PeerComm*, double* P1, double* P2, int nb, cudaStream_t s);
void comm_check(PeerComm*);  // throws if a spin-wait timed out
int launch_halo_zero(const DevParams& P, double* x, cudaStream_t s);  // ghost layers of the processor sides := 0

// fem.cu: conforming Qk residual / jacobian_apply (coloured scatter)
struct FemPlan;
FemPlan* fem_plan_create(const DevParams& P, const int8_t* bctype_dev, const Kron1D& K);
void fem_plan_invalidate(FemPlan*);  // coefficients changed: drop the cached R(0)
void fem_plan_destroy(FemPlan*);
void launch_fem_vector(FemPlan* plan, const DevParams& P, const double* x, double* y, bool residual, bool overwrite,
                       cudaStream_t s);
struct QkLayout;  // host_tables.h
// fem_kron.cu: Kronecker-form conforming Qk apply (diagonal A, b = 0): warp-shuffle / smem / register assembly
void launch_fem_kron(const DevParams& P, const QkLayout& L, const double* MinvK, const double* M, const double* x, double* y,
                     const double* r0, bool overwrite, bool fuse_constraints, bool diag, cudaStream_t s);
// fem.cu: point diagonal of the conforming Jacobian (diagonal A, b = 0), constrained rows = 1
void launch_fem_diagonal(FemPlan* plan, const DevParams& P, double* d, cudaStream_t s);
// dg_blockjac.cu: point diagonal of the QkDG Jacobian (same support as the block-Jacobi preconditioner)
int launch_dg_diagonal(const DevParams& P, const Kron1D& K, double* d, cudaStream_t s);
const QkLayout& fem_plan_layout(const FemPlan*);
const uint64_t* fem_plan_constrained(const FemPlan*, long long* n);  // device list of constrained DOFs

// matrix.cu: sparsity pattern (CSR / block CSR), assembled Jacobian, SpMV
struct MatrixPlan;
MatrixPlan* matrix_plan_create(const DevParams& P, FemPlan* fem, cudaStream_t s);
void matrix_plan_destroy(MatrixPlan*);
void matrix_pattern_size(MatrixPlan*, int layout, uint64_t* nrows, uint64_t* nnz);
int matrix_pattern_write(MatrixPlan*, int layout, void* rowptr, bool rowptr_dev, void* colidx, bool colidx_dev,
                         bool col32, cudaStream_t s);
int matrix_assemble(MatrixPlan*, int layout, double* values, bool values_dev, bool fresh, int* errflag,
                    cudaStream_t s);
int matrix_mv(MatrixPlan*, int layout, const double* values, const double* x, double* y, cudaStream_t s);

}  // namespace pdb

// dg_blockjac.cu — exact block-Jacobi preconditioner  z = D^-1 r  for the QkDG / SIPG operator, matrix-free.
//
// What it replaces in the reference (paths relative to /root/reference/dune/pdelab/):
//   AssembledBlockJacobiPreconditionerLocalOperator      backend/istl/matrixfree/assembledblockjacobipreconditioner.hh:96-230
//     (assembles every diagonal block with BlockDiagonalLocalOperatorWrapper, localoperator/blockdiagonalwrapper.hh,
//      LU-factorises it with Eigen and applies the inverse in jacobian_apply_volume)
//   GridOperatorPreconditioner::apply                    backend/istl/matrixfree/gridoperatorpreconditioner.hh:81-87
//   as used by ISTLBackend_SEQ_MatrixFree_Base           backend/istl/matrixfree/backends.hh:62-143
//
// The reference stores an n x n LU per cell (27 x 27 doubles = 5.8 KB per Q2 cell: 12 GB at 128^3
// cells).  On an axis-aligned grid with a cell-wise constant diagonal tensor the diagonal block is a
// Kronecker sum
//     D_e = |K| (M (x) M (x) M) [ T_x (+) T_y (+) T_z + c_e I ],   T_d = M^-1 S_d,
// with S_d the symmetric (SIPG, theta = -1) own-cell part of the 1-D operator in direction d.  Each
// T_d is diagonalised once per coefficient set (generalised symmetric eigenproblem S v = lambda M v,
// n1 <= 3: Cholesky factor of M on the host, cyclic Jacobi rotations on L^-1 S L^-T in the set-up
// kernel):  T_d = G_d^T Lambda_d G_d^-T  with  G_d = Q_d^T L^-1.  Then
//     D_e^-1 = (G_x (x) G_y (x) G_z)^T  diag( 1 / (|K| (lambda_i + mu_j + nu_k + c_e)) )  (G_x (x) G_y (x) G_z),
// i.e. 2 dim small sweeps and one scaling per cell — about the cost of the operator itself — from
// dim (n1 + n1^2) stored doubles per cell (36 for Q2 in 3-D: 10.7 B/DOF instead of 216 B/DOF).
// The result is the exact inverse of the diagonal blocks (checked against a dense inverse of the
// assembled blocks, tests/test_gpu_solver.py); cells of a ghost layer (constrained rows) get z = 0.

#include "common.cuh"
#include "dg_face.cuh"

namespace pdb {

struct BlockJacPlan {
  double* data = nullptr;  // [dim * (n1 + n1^2)][ncells]
  bool valid = false;
  double Linv[MAX_N1 * MAX_N1] = {};  // inverse Cholesky factor of the 1-D mass matrix, row-major n1 x n1
};

namespace {

using namespace dgface;

template <int K>
struct BJConst {
  double Linv[(K + 1) * (K + 1)];
};

template <int DIM, int K>
__global__ void __launch_bounds__(128) blockjac_setup_kernel(const DevParams P, const SmallConst<K> C, const BJConst<K> B,
                                                             double* __restrict__ data) {
  constexpr int N1 = K + 1, PER = N1 + N1 * N1;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncells) return;
  const int Nx = P.N[0], Ny = P.N[1];
  int g[3];
  g[0] = (int)(cell % Nx);
  g[1] = (int)((cell / Nx) % Ny);
  g[2] = DIM == 3 ? (int)(cell / ((long long)Nx * Ny)) : 0;
  const long long stride[3] = {1, Nx, (long long)Nx * Ny};
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double A0, cs[2], co[2], cg[2];
    bool onb[2];
    direction_coefs<K>(P, C, cell, g, d, stride, A0, cs, co, cg, onb);
    double T[N1 * N1], eL[N1], eR[N1];
    own_matrix<K>(C, A0, cs[0], cg[0], cs[1], cg[1], T, eL, eR);
    // S = M T (symmetric for SIPG), A = L^-1 S L^-T
    double S[N1 * N1], X[N1 * N1], A[N1 * N1];
#pragma unroll
    for (int i = 0; i < N1; i++)
#pragma unroll
      for (int j = 0; j < N1; j++) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < N1; k++) v = fma(C.M[i * N1 + k], T[k * N1 + j], v);
        S[i * N1 + j] = v;
      }
#pragma unroll
    for (int i = 0; i < N1; i++)
#pragma unroll
      for (int j = 0; j < N1; j++) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < N1; k++) v = fma(B.Linv[i * N1 + k], S[k * N1 + j], v);
        X[i * N1 + j] = v;
      }
#pragma unroll
    for (int i = 0; i < N1; i++)
#pragma unroll
      for (int j = 0; j < N1; j++) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < N1; k++) v = fma(X[i * N1 + k], B.Linv[j * N1 + k], v);
        A[i * N1 + j] = v;
      }
#pragma unroll
    for (int i = 0; i < N1; i++)
#pragma unroll
      for (int j = i + 1; j < N1; j++) A[i * N1 + j] = A[j * N1 + i] = 0.5 * (A[i * N1 + j] + A[j * N1 + i]);
    // cyclic Jacobi: A = Q diag(lambda) Q^T, columns of Q are the eigenvectors
    double Q[N1 * N1];
#pragma unroll
    for (int i = 0; i < N1 * N1; i++) Q[i] = (i / N1 == i % N1) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < (N1 == 2 ? 2 : 12); sweep++) {
#pragma unroll
      for (int p = 0; p < N1; p++)
#pragma unroll
        for (int q = p + 1; q < N1; q++) {
          const double apq = A[p * N1 + q];
          const double app = A[p * N1 + p], aqq = A[q * N1 + q];
          if (fabs(apq) > 1e-300 + 1e-18 * (fabs(app) + fabs(aqq))) {
            const double th = (aqq - app) / (2.0 * apq);
            const double tt = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
            for (int k = 0; k < N1; k++) {  // A <- A R
              const double akp = A[k * N1 + p], akq = A[k * N1 + q];
              A[k * N1 + p] = c * akp - s * akq;
              A[k * N1 + q] = s * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < N1; k++) {  // A <- R^T A
              const double apk = A[p * N1 + k], aqk = A[q * N1 + k];
              A[p * N1 + k] = c * apk - s * aqk;
              A[q * N1 + k] = s * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < N1; k++) {  // Q <- Q R
              const double qkp = Q[k * N1 + p], qkq = Q[k * N1 + q];
              Q[k * N1 + p] = c * qkp - s * qkq;
              Q[k * N1 + q] = s * qkp + c * qkq;
            }
          }
        }
    }
    // G = Q^T L^-1; store lambda (n1) then G (n1 x n1, row-major), item-major / cell-minor (coalesced)
    double* out = data + (long long)d * PER * P.ncells + cell;
#pragma unroll
    for (int i = 0; i < N1; i++) out[(long long)i * P.ncells] = A[i * N1 + i];
#pragma unroll
    for (int i = 0; i < N1; i++)
#pragma unroll
      for (int j = 0; j < N1; j++) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < N1; k++) v = fma(Q[k * N1 + i], B.Linv[k * N1 + j], v);
        out[(long long)(N1 + i * N1 + j) * P.ncells] = v;
      }
  }
}

// t <- (Mat along AXIS) t  or its transpose
template <int DIM, int K, int AXIS, bool TRANS>
__device__ __forceinline__ void bj_sweep(const double (&G)[(K + 1) * (K + 1)], double (&t)[SL<DIM, K>::N]) {
  constexpr int N1 = K + 1, N = SL<DIM, K>::N;
  constexpr int S = AXIS == 0 ? 1 : (AXIS == 1 ? N1 : N1 * N1);
#pragma unroll
  for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
    for (int lo = 0; lo < S; lo++) {
      const int base = hi * S * N1 + lo;
      double in[N1];
#pragma unroll
      for (int j = 0; j < N1; j++) in[j] = t[base + j * S];
#pragma unroll
      for (int i = 0; i < N1; i++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < N1; j++) acc = fma(TRANS ? G[j * N1 + i] : G[i * N1 + j], in[j], acc);
        t[base + i * S] = acc;
      }
    }
}

// MODE 0: z = D^-1 r;  1: z = D r;  2: z -= D r  (D = G^-1 diag(|K| den) G^-T with G^-1 = M G^T, G^-T = G M per direction)
template <int DIM, int K, int MODE = 0>
__global__ void __launch_bounds__(128) blockjac_apply_kernel(const DevParams P, const double* __restrict__ data,
                                                             const double* __restrict__ r, double* __restrict__ z,
                                                             const SmallConst<K> C = SmallConst<K>()) {
  constexpr int N1 = K + 1, N = SL<DIM, K>::N, PER = N1 + N1 * N1;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncells) return;
  const int Nx = P.N[0], Ny = P.N[1];
  const int g[3] = {(int)(cell % Nx), (int)((cell / Nx) % Ny), DIM == 3 ? (int)(cell / ((long long)Nx * Ny)) : 0};
  bool constrained = false;
#pragma unroll
  for (int d = 0; d < DIM; d++)
    constrained |= (g[d] == 0 && P.side_kind[d][0] == PDB200_SIDE_PROCESSOR) ||
                   (g[d] == P.N[d] - 1 && P.side_kind[d][1] == PDB200_SIDE_PROCESSOR);
  double t[N];
  if (constrained) {
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = 0.0;
    store_cell<N>(z + cell * N, t);
    return;
  }
  load_cell<N>(r + cell * N, t);
  double lam[DIM][N1], G[DIM][N1 * N1];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const double* src = data + (long long)d * PER * P.ncells + cell;
#pragma unroll
    for (int i = 0; i < N1; i++) lam[d][i] = __ldg(src + (long long)i * P.ncells);
#pragma unroll
    for (int i = 0; i < N1 * N1; i++) G[d][i] = __ldg(src + (long long)(N1 + i) * P.ncells);
  }
  if (MODE != 0) {
    small_mass<DIM, K, 0>(C, 1.0, t);
    small_mass<DIM, K, 1>(C, 1.0, t);
    if (DIM == 3) small_mass<DIM, K, 2>(C, 1.0, t);
  }
  bj_sweep<DIM, K, 0, false>(G[0], t);
  bj_sweep<DIM, K, 1, false>(G[1], t);
  if (DIM == 3) bj_sweep<DIM, K, 2, false>(G[DIM - 1], t);
  const double cc = P.c ? __ldg(P.c + cell) : 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int i0 = i % N1, i1 = (i / N1) % N1, i2 = i / (N1 * N1);
    const double den = lam[0][i0] + lam[1][i1] + (DIM == 3 ? lam[DIM - 1][i2] : 0.0) + cc;
    t[i] = MODE == 0 ? t[i] / (P.vol * den) : t[i] * (P.vol * den);
  }
  bj_sweep<DIM, K, 0, true>(G[0], t);
  bj_sweep<DIM, K, 1, true>(G[1], t);
  if (DIM == 3) bj_sweep<DIM, K, 2, true>(G[DIM - 1], t);
  if (MODE != 0) {
    small_mass<DIM, K, 0>(C, 1.0, t);
    small_mass<DIM, K, 1>(C, 1.0, t);
    if (DIM == 3) small_mass<DIM, K, 2>(C, 1.0, t);
  }
  if (MODE == 2) {
    double old[N];
    load_cell<N>(z + cell * N, old);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = old[i] - t[i];
  }
  store_cell<N>(z + cell * N, t);
}

// One wavefront of a block SOR sweep (BlockSORPreconditionerLocalOperator, backend/istl/matrixfree/
// blocksorpreconditioner.hh:36-60:  for T_i in index order:  a_i = d_i - sum_{j != i} A_ij v_j,  D_i b_i = a_i,
// v_i = (1 - omega) v_i + omega b_i,  in place).  Cells couple through faces only, so cell (i, j, k) depends on the
// already updated (i-1, j, k), (i, j-1, k), (i, j, k-1) and on the old values of its upper neighbours: all cells of the
// hyperplane i + j + k = s are independent and the sweep over s = 0 .. sum(N) - dim reproduces the lexicographic
// (index-set) order of the reference exactly.  One thread per cell of the hyperplane: the cell's row of J v by the
// Kronecker sweeps of dg_small.cu from the current v, then the exact block inverse by fast diagonalisation:
//     v_i += omega D_i^-1 (d_i - (J v)_i)        (the same update, written with the full row).
template <int DIM, int K>
__global__ void __launch_bounds__(128) blocksor_wave_kernel(const DevParams P, const SmallConst<K> C,
                                                            const double* __restrict__ data, const double* __restrict__ dvec,
                                                            double* v, double omega, int wave) {
  constexpr int N1 = K + 1, N = SL<DIM, K>::N, PER = N1 + N1 * N1;
  const int Nx = P.N[0], Ny = P.N[1];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ntan = DIM == 3 ? (long long)Ny * P.N[2] : Ny;
  if (tid >= ntan) return;
  int g[3];
  g[1] = (int)(tid % Ny);
  g[2] = DIM == 3 ? (int)(tid / Ny) : 0;
  g[0] = wave - g[1] - g[2];
  if (g[0] < 0 || g[0] >= Nx) return;
  const long long stride[3] = {1, Nx, (long long)Nx * Ny};
  const long long cell = g[0] + stride[1] * g[1] + stride[2] * g[2];
  double o[N], t[N];
  {
    const double* p = v + cell * N;  // plain loads: v is updated in place by earlier wavefronts
#pragma unroll
    for (int i = 0; i < N; i++) o[i] = p[i];
  }
  const double creact = P.c ? __ldg(P.c + cell) : 0.0;
  bool constrained = false;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double A0, cs[2], co[2], cg[2];
    bool onb[2];
    constrained |= direction_coefs<K>(P, C, cell, g, d, stride, A0, cs, co, cg, onb);
    double nb[2][N];
#pragma unroll
    for (int side = 0; side < 2; side++) {
      if (!onb[side]) {
        const double* p = v + (cell + (side ? stride[d] : -stride[d])) * N;
#pragma unroll
        for (int i = 0; i < N; i++) nb[side][i] = p[i];
      } else {
#pragma unroll
        for (int i = 0; i < N; i++) nb[side][i] = 0.0;
      }
    }
    if (d == 0)
      small_sweep<DIM, K, 0, true>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], creact, t);
    else if (d == 1)
      small_sweep<DIM, K, 1, false>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], 0.0, t);
    else
      small_sweep<DIM, K, 2, false>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], 0.0, t);
  }
  if (constrained) {  // ghost cells of an overlapping partition: the correction vanishes there
#pragma unroll
    for (int i = 0; i < N; i++) v[cell * N + i] = 0.0;
    return;
  }
  small_mass<DIM, K, 0>(C, C.vol, t);
  small_mass<DIM, K, 1>(C, 1.0, t);
  if (DIM == 3) small_mass<DIM, K, 2>(C, 1.0, t);
  {
    double dd[N];
    load_cell<N>(dvec + cell * N, dd);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = dd[i] - t[i];
  }
  double lam[DIM][N1], G[DIM][N1 * N1];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const double* src = data + (long long)d * PER * P.ncells + cell;
#pragma unroll
    for (int i = 0; i < N1; i++) lam[d][i] = __ldg(src + (long long)i * P.ncells);
#pragma unroll
    for (int i = 0; i < N1 * N1; i++) G[d][i] = __ldg(src + (long long)(N1 + i) * P.ncells);
  }
  bj_sweep<DIM, K, 0, false>(G[0], t);
  bj_sweep<DIM, K, 1, false>(G[1], t);
  if (DIM == 3) bj_sweep<DIM, K, 2, false>(G[DIM - 1], t);
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int i0 = i % N1, i1 = (i / N1) % N1, i2 = i / (N1 * N1);
    const double den = lam[0][i0] + lam[1][i1] + (DIM == 3 ? lam[DIM - 1][i2] : 0.0) + creact;
    t[i] = t[i] / (P.vol * den);
  }
  bj_sweep<DIM, K, 0, true>(G[0], t);
  bj_sweep<DIM, K, 1, true>(G[1], t);
  if (DIM == 3) bj_sweep<DIM, K, 2, true>(G[DIM - 1], t);
#pragma unroll
  for (int i = 0; i < N; i++) v[cell * N + i] = fma(omega, t[i], o[i]);
}

// point diagonal of the Jacobian: diag(D_e) = |K| [ sum_d (M T_d)_{i_d i_d} prod_{e != d} M_{i_e i_e} + c prod_d M_{i_d i_d} ]
// (PointDiagonalLocalOperatorWrapper, localoperator/pointdiagonalwrapper.hh); 1 on constrained (ghost) rows
template <int DIM, int K>
__global__ void __launch_bounds__(128) dg_diagonal_kernel(const DevParams P, const SmallConst<K> C, double* __restrict__ dg) {
  constexpr int N1 = K + 1, N = SL<DIM, K>::N;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncells) return;
  const int Nx = P.N[0], Ny = P.N[1];
  int g[3];
  g[0] = (int)(cell % Nx);
  g[1] = (int)((cell / Nx) % Ny);
  g[2] = DIM == 3 ? (int)(cell / ((long long)Nx * Ny)) : 0;
  const long long stride[3] = {1, Nx, (long long)Nx * Ny};
  double Sd[DIM][N1];
  bool constrained = false;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double A0, cs[2], co[2], cg[2];
    bool onb[2];
    constrained |= direction_coefs<K>(P, C, cell, g, d, stride, A0, cs, co, cg, onb);
    double T[N1 * N1], eL[N1], eR[N1];
    own_matrix<K>(C, A0, cs[0], cg[0], cs[1], cg[1], T, eL, eR);
#pragma unroll
    for (int i = 0; i < N1; i++) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < N1; k++) v = fma(C.M[i * N1 + k], T[k * N1 + i], v);
      Sd[d][i] = v;
    }
  }
  const double cc = P.c ? __ldg(P.c + cell) : 0.0;
  double t[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int id[3] = {i % N1, (i / N1) % N1, i / (N1 * N1)};
    double m = cc, v = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      double pd = Sd[d][id[d]];
#pragma unroll
      for (int e = 0; e < DIM; e++)
        if (e != d) pd *= C.M[id[e] * N1 + id[e]];
      v += pd;
      m *= C.M[id[d] * N1 + id[d]];
    }
    t[i] = constrained ? 1.0 : P.vol * (v + m);
  }
  store_cell<N>(dg + cell * N, t);
}

template <int DIM, int K>
void setup_variant(BlockJacPlan* plan, const DevParams& P, const Kron1D& K1, cudaStream_t s) {
  constexpr int N1 = K + 1, PER = N1 + N1 * N1;
  SmallConst<K> C;
  fill_small_const<K>(C, P, K1);
  BJConst<K> B;
  for (int i = 0; i < N1 * N1; i++) B.Linv[i] = plan->Linv[i];
  if (!plan->data) PDB_CUDA(cudaMalloc(&plan->data, (size_t)DIM * PER * P.ncells * sizeof(double)));
  blockjac_setup_kernel<DIM, K><<<(unsigned)((P.ncells + 127) / 128), 128, 0, s>>>(P, C, B, plan->data);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace

bool dg_blockjac_supported(const DevParams& P) {
  return P.dg && P.basis == PDB200_BASIS_LAGRANGE && kron_coefficients(P) && P.m >= P.k + 1 && P.theta == -1.0 &&
         (P.dim == 2 || P.dim == 3) && (P.k == 1 || P.k == 2);
}

BlockJacPlan* dg_blockjac_create(const DevParams& P, const Kron1D& K1) {
  BlockJacPlan* plan = new BlockJacPlan;
  // Cholesky factor of the 1-D mass matrix and its inverse (lower triangular), in long double
  const int n1 = P.k + 1;
  long double L[MAX_N1][MAX_N1] = {}, Li[MAX_N1][MAX_N1] = {};
  for (int j = 0; j < n1; j++) {
    long double d = K1.M[j * MAX_N1 + j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
    L[j][j] = sqrtl(d);
    for (int i = j + 1; i < n1; i++) {
      long double v = K1.M[i * MAX_N1 + j];
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k];
      L[i][j] = v / L[j][j];
    }
  }
  for (int c = 0; c < n1; c++)  // forward substitution L x = e_c
    for (int i = 0; i < n1; i++) {
      long double v = i == c ? 1.0L : 0.0L;
      for (int k = 0; k < i; k++) v -= L[i][k] * Li[k][c];
      Li[i][c] = v / L[i][i];
    }
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n1; j++) plan->Linv[i * n1 + j] = (double)Li[i][j];
  return plan;
}

void dg_blockjac_destroy(BlockJacPlan* plan) {
  if (!plan) return;
  if (plan->data) cudaFree(plan->data);
  delete plan;
}
void dg_blockjac_invalidate(BlockJacPlan* plan) {
  if (plan) plan->valid = false;
}

int launch_dg_diagonal(const DevParams& P, const Kron1D& K1, double* d, cudaStream_t s) {
  if (!(P.dg && P.basis == PDB200_BASIS_LAGRANGE && kron_coefficients(P) && P.m >= P.k + 1 &&
        (P.dim == 2 || P.dim == 3) && (P.k == 1 || P.k == 2)))
    throw Error("matrix-free point diagonal: needs QkDG in the Lagrange basis (k = 1, 2; dim = 2, 3), diagonal A, b = 0");
  const unsigned blocks = (unsigned)((P.ncells + 127) / 128);
#define PDB_DD(DD, KK)                                              \
  if (P.dim == DD && P.k == KK) {                                   \
    SmallConst<KK> C;                                               \
    fill_small_const<KK>(C, P, K1);                                 \
    dg_diagonal_kernel<DD, KK><<<blocks, 128, 0, s>>>(P, C, d);     \
  }
  PDB_DD(2, 1) PDB_DD(2, 2) PDB_DD(3, 1) PDB_DD(3, 2)
#undef PDB_DD
  PDB_CUDA(cudaGetLastError());
  return 1;
}

// z = D^-1 r; (re)builds the per-cell eigen-decompositions if the coefficients changed.  Returns launches.
int launch_dg_blockjac(BlockJacPlan* plan, const DevParams& P, const Kron1D& K1, const double* r, double* z,
                       cudaStream_t s) {
  if (!dg_blockjac_supported(P))
    throw Error("block Jacobi: needs QkDG (k = 1, 2; dim = 2, 3), SIPG, diagonal A, b = 0");
  int launches = 0;
  const unsigned blocks = (unsigned)((P.ncells + 127) / 128);
#define PDB_BJ(DD, KK)                                                                        \
  if (P.dim == DD && P.k == KK) {                                                             \
    if (!plan->valid) {                                                                       \
      setup_variant<DD, KK>(plan, P, K1, s);                                                  \
      launches++;                                                                             \
    }                                                                                         \
    blockjac_apply_kernel<DD, KK><<<blocks, 128, 0, s>>>(P, plan->data, r, z);                \
    launches++;                                                                               \
  }
  PDB_BJ(2, 1) PDB_BJ(2, 2) PDB_BJ(3, 1) PDB_BJ(3, 2)
#undef PDB_BJ
  PDB_CUDA(cudaGetLastError());
  plan->valid = true;
  return launches;
}

// mode 1: y = D z (BlockDiagonalLocalOperatorWrapper, localoperator/blockdiagonalwrapper.hh);  mode 2: y -= D z
// (with y = J z before: the block off-diagonal part, localoperator/blockoffdiagonalwrapper.hh)
int launch_dg_blockdiag(BlockJacPlan* plan, const DevParams& P, const Kron1D& K1, const double* z, double* y, int mode,
                        cudaStream_t s) {
  if (!dg_blockjac_supported(P))
    throw Error("block diagonal: needs QkDG (k = 1, 2; dim = 2, 3), SIPG, diagonal A, b = 0");
  if (mode != 1 && mode != 2) throw Error("launch_dg_blockdiag: mode");
  int launches = 0;
  const unsigned blocks = (unsigned)((P.ncells + 127) / 128);
#define PDB_BD(DD, KK)                                                                              \
  if (P.dim == DD && P.k == KK) {                                                                   \
    if (!plan->valid) {                                                                             \
      setup_variant<DD, KK>(plan, P, K1, s);                                                        \
      launches++;                                                                                   \
    }                                                                                               \
    SmallConst<KK> C;                                                                               \
    fill_small_const<KK>(C, P, K1);                                                                 \
    if (mode == 1) blockjac_apply_kernel<DD, KK, 1><<<blocks, 128, 0, s>>>(P, plan->data, z, y, C); \
    else blockjac_apply_kernel<DD, KK, 2><<<blocks, 128, 0, s>>>(P, plan->data, z, y, C);           \
    launches++;                                                                                     \
  }
  PDB_BD(2, 1) PDB_BD(2, 2) PDB_BD(3, 1) PDB_BD(3, 2)
#undef PDB_BD
  PDB_CUDA(cudaGetLastError());
  plan->valid = true;
  return launches;
}

// one block SOR sweep in place: forward (index order) or backward (reverse order); zero_start: v := 0 first
// (what a Krylov solver hands to a preconditioner).  Returns the number of launches.
int launch_dg_blocksor(BlockJacPlan* plan, const DevParams& P, const Kron1D& K1, const double* d, double* v, double omega,
                       bool backward, bool zero_start, cudaStream_t s) {
  if (!dg_blockjac_supported(P))
    throw Error("block SOR: needs QkDG (k = 1, 2; dim = 2, 3), SIPG, diagonal A, b = 0");
  int launches = 0;
  if (zero_start) PDB_CUDA(cudaMemsetAsync(v, 0, (size_t)P.ndofs * sizeof(double), s));
  const long long ntan = P.dim == 3 ? (long long)P.N[1] * P.N[2] : P.N[1];
  const unsigned blocks = (unsigned)((ntan + 127) / 128);
  int nwaves = 1;  // hyperplanes i + j + k = 0 .. sum_d (N_d - 1)
  for (int dd = 0; dd < P.dim; dd++) nwaves += P.N[dd] - 1;
#define PDB_SOR(DD, KK)                                                                                      \
  if (P.dim == DD && P.k == KK) {                                                                            \
    if (!plan->valid) {                                                                                      \
      setup_variant<DD, KK>(plan, P, K1, s);                                                                 \
      launches++;                                                                                            \
    }                                                                                                        \
    SmallConst<KK> C;                                                                                        \
    fill_small_const<KK>(C, P, K1);                                                                          \
    for (int w = 0; w < nwaves; w++)                                                                         \
      blocksor_wave_kernel<DD, KK><<<blocks, 128, 0, s>>>(P, C, plan->data, d, v, omega, backward ? nwaves - 1 - w : w); \
    launches += nwaves;                                                                                      \
  }
  PDB_SOR(2, 1) PDB_SOR(2, 2) PDB_SOR(3, 1) PDB_SOR(3, 2)
#undef PDB_SOR
  PDB_CUDA(cudaGetLastError());
  plan->valid = true;
  return launches;
}

}  // namespace pdb

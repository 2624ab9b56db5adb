// fem.cu — conforming Qk (k = 1, 2; dim = 2, 3) residual and exact jacobian_apply for
// ConvectionDiffusionFEM.
//
// What it computes (paths relative to /root/reference/dune/pdelab/):
//   GridOperator::residual / jacobian_apply        gridoperator/gridoperator.hh:176-197
//   ConvectionDiffusionFEM::alpha_volume           localoperator/convectiondiffusionfem.hh:63-136
//   ConvectionDiffusionFEM::alpha_boundary         :207-275 (bctype at the face centre)
//   gather x[ci(i)] / scatter r[ci(i)] += rl[i]    gridoperator/default/residualengine.hh:131-233,
//                                                  gridfunctionspace/lfsindexcache.hh:603-633
//   postAssembly -> constrain_residual             constraints/common/constraints.hh:904-915
// jacobian_apply is the exact derivative J z of the (affine) residual, not the reference's
// finite-difference mixin (numericaljacobianapply.hh:54-85) — documented deviation, DESIGN.md §2.
//
// Mapping to the machine.  The reference scatters cell-local results into the global vector cell by
// cell.  Here a CTA owns a box of lattice points (k*T_d per direction) and evaluates the
// (T_d + 1)-cell box that touches them:
//   1. the lattice values of the cell box are gathered once into shared memory through the
//      closed-form container index (LFSIndexCache replaced by arithmetic, host_tables.h);
//   2. one thread per cell evaluates alpha_volume sum-factorised in registers (1-D tables B, D from
//      the constant bank: 9 forward + 9 backward 1-D sweeps in 3-D instead of the reference's dense
//      O(n q) loops) and stores its n local results in shared memory;
//   3. one thread per owned lattice point adds the <= 2^dim cell contributions in ascending cell
//      order — the reference's accumulation order — and writes the row exactly once.
// No atomics, deterministic, every global value read and written once per CTA that needs it.

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "host_tables.h"

namespace pdb {

// exactly integrated 1-D matrices for the Kronecker form of alpha_volume (diagonal A, b = 0)
struct FemKron {
  double MinvK[MAX_N1 * MAX_N1];  // M^-1 K, row-major with leading dimension n1
  double M[MAX_N1 * MAX_N1];
};

struct FemPlan {
  QkLayout L;
  uint64_t* con = nullptr;  // constrained DOFs (device)
  long long ncon = 0;
  FemKron kron;
  double* r0 = nullptr;     // R(0) of the affine residual (Kronecker path), cached per coefficient set
  bool r0_valid = false;
  bool fused_constraints = false;  // the last launch already zeroed the constrained rows
};

namespace {

template <int DIM>
struct Tile;
template <>
struct Tile<2> {
  static constexpr int T0 = 16, T1 = 8, T2 = 1;
};
template <>
struct Tile<3> {
  static constexpr int T0 = 8, T1 = 4, T2 = 4;
};

constexpr int FEM_THREADS = 256;

template <int DIM, int N1>
struct LocalSize {
  static constexpr int N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
};

// One 1-D contraction along AXIS of a tensor with N1 entries per direction.
//   TRANS = false: out[.., q, ..] (+)= sum_i Mat[q*N1 + i] in[.., i, ..]
//   TRANS = true : out[.., i, ..] (+)= sum_q Mat[q*N1 + i] in[.., q, ..]
template <int DIM, int N1, int AXIS, bool TRANS, bool ACC>
__device__ __forceinline__ void sweep1d(const double* __restrict__ Mat, const double (&in)[LocalSize<DIM, N1>::N],
                                        double (&out)[LocalSize<DIM, N1>::N]) {
  constexpr int N = LocalSize<DIM, N1>::N;
  constexpr int S = AXIS == 0 ? 1 : (AXIS == 1 ? N1 : N1 * N1);
#pragma unroll
  for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
    for (int lo = 0; lo < S; lo++) {
      const int base = hi * S * N1 + lo;
#pragma unroll
      for (int o = 0; o < N1; o++) {
        double acc = ACC ? out[base + o * S] : 0.0;
#pragma unroll
        for (int i = 0; i < N1; i++) acc = fma(TRANS ? Mat[i * N1 + o] : Mat[o * N1 + i], in[base + i * S], acc);
        out[base + o * S] = acc;
      }
    }
}

// alpha_volume of one cell, sum-factorised (convectiondiffusionfem.hh:94-135)
template <int DIM, int K, bool RESIDUAL>
__device__ __forceinline__ void fem_cell_volume(const DevParams& P, long long cell,
                                                const double (&x)[LocalSize<DIM, K + 1>::N],
                                                double (&r)[LocalSize<DIM, K + 1>::N]) {
  constexpr int N1 = K + 1, N = LocalSize<DIM, N1>::N;
  const double* B = P.P;   // B[q*N1 + i] = p_i(x_q)
  const double* D = P.DP;  // D[q*N1 + i] = p_i'(x_q)
  double A[3][3];
  load_A_cell(P, cell, A);
  const bool pwA = pw_A(P);
  double bv[3];
  double u[N], gx[N], gy[N], gz[N];
  if (DIM == 3) {
    double t1[N], d1[N], t2[N], t2y[N], d2[N];
    sweep1d<DIM, N1, 0, false, false>(B, x, t1);
    sweep1d<DIM, N1, 0, false, false>(D, x, d1);
    sweep1d<DIM, N1, 1, false, false>(B, t1, t2);
    sweep1d<DIM, N1, 1, false, false>(D, t1, t2y);
    sweep1d<DIM, N1, 1, false, false>(B, d1, d2);
    sweep1d<DIM, N1, 2, false, false>(B, t2, u);
    sweep1d<DIM, N1, 2, false, false>(D, t2, gz);
    sweep1d<DIM, N1, 2, false, false>(B, t2y, gy);
    sweep1d<DIM, N1, 2, false, false>(B, d2, gx);
  } else {
    double t1[N], d1[N];
    sweep1d<DIM, N1, 0, false, false>(B, x, t1);
    sweep1d<DIM, N1, 0, false, false>(D, x, d1);
    sweep1d<DIM, N1, 1, false, false>(B, t1, u);
    sweep1d<DIM, N1, 1, false, false>(D, t1, gy);
    sweep1d<DIM, N1, 1, false, false>(B, d1, gx);
  }
  // quadrature-point work: flux = A grad u - u b, source = c u - f   (:110-134)
#pragma unroll
  for (int q = 0; q < N; q++) {
    const int q0 = q % N1, q1 = (q / N1) % N1, q2 = q / (N1 * N1);
    double w = P.wq[q0] * P.wq[q1];
    if (DIM == 3) w *= P.wq[q2];
    const double factor = w * P.vol;
    if (pwA) load_A_at(P, cell, q, A);      // !permeabilityIsConstantPerCell, convectiondiffusionfem.hh:97-100
    load_b(P, cell, q, bv);                 // param.b / param.c at the quadrature point, :127-128
    const double cc = load_c(P, cell, q);
    double g[3] = {gx[q] * P.ih[0], gy[q] * P.ih[1], DIM == 3 ? gz[q] * P.ih[2] : 0.0};
    double s = cc * u[q];
    if (RESIDUAL && P.f) s -= __ldg(P.f + cell * N + q);
    double F[3];
#pragma unroll
    for (int a = 0; a < 3; a++) F[a] = (A[a][0] * g[0] + A[a][1] * g[1] + A[a][2] * g[2] - u[q] * bv[a]) * factor * P.ih[a];
    u[q] = s * factor;
    gx[q] = F[0];
    gy[q] = F[1];
    if (DIM == 3) gz[q] = F[2];
  }
  if (DIM == 3) {
    double a1[N], a2[N], a3[N], b1[N], b2[N];
    sweep1d<DIM, N1, 2, true, false>(B, u, a1);
    sweep1d<DIM, N1, 2, true, true>(D, gz, a1);
    sweep1d<DIM, N1, 2, true, false>(B, gy, a2);
    sweep1d<DIM, N1, 2, true, false>(B, gx, a3);
    sweep1d<DIM, N1, 1, true, false>(B, a1, b1);
    sweep1d<DIM, N1, 1, true, true>(D, a2, b1);
    sweep1d<DIM, N1, 1, true, false>(B, a3, b2);
    sweep1d<DIM, N1, 0, true, false>(B, b1, r);
    sweep1d<DIM, N1, 0, true, true>(D, b2, r);
  } else {
    double a1[N], a2[N];
    sweep1d<DIM, N1, 1, true, false>(B, u, a1);
    sweep1d<DIM, N1, 1, true, true>(D, gy, a1);
    sweep1d<DIM, N1, 1, true, false>(B, gx, a2);
    sweep1d<DIM, N1, 0, true, false>(B, a1, r);
    sweep1d<DIM, N1, 0, true, true>(D, a2, r);
  }
}

// The same cell integral for a cell-wise constant DIAGONAL tensor, b = 0, without the source term:
//   r = |K| (M (x) M (x) M) [ sum_d (A_dd / h_d^2) (M^-1 K)_d x + c x ]
// (the (k+1)-point Gauss rule of convectiondiffusionfem.hh:93-94 integrates these polynomials exactly,
// so this equals alpha_volume to rounding): one sweep per direction plus three mass sweeps instead
// of 18 quadrature sweeps.
template <int DIM, int K>
__device__ __forceinline__ void fem_cell_volume_kron(const DevParams& P, const FemKron& Q, long long cell,
                                                     const double (&x)[LocalSize<DIM, K + 1>::N],
                                                     double (&r)[LocalSize<DIM, K + 1>::N]) {
  constexpr int N1 = K + 1, N = LocalSize<DIM, N1>::N;
  double a[3] = {1.0, 1.0, 1.0};
  if (P.a_mode == PDB200_A_SCALAR) {
    a[0] = a[1] = a[2] = __ldg(P.A + cell);
  } else if (P.a_mode == PDB200_A_DIAGONAL) {
#pragma unroll
    for (int d = 0; d < DIM; d++) a[d] = __ldg(P.A + cell * DIM + d);
  }
  const double cc = P.c ? __ldg(P.c + cell) : 0.0;
  double t[N];
#pragma unroll
  for (int i = 0; i < N; i++) t[i] = cc * x[i];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const int S = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
    const double al = a[d] * P.ih[d] * P.ih[d];
#pragma unroll
    for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
      for (int lo = 0; lo < S; lo++) {
        const int base = hi * S * N1 + lo;
#pragma unroll
        for (int o = 0; o < N1; o++) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < N1; i++) acc = fma(Q.MinvK[o * N1 + i], x[base + i * S], acc);
          t[base + o * S] = fma(al, acc, t[base + o * S]);
        }
      }
  }
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const int S = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
    const double sc = d == 0 ? P.vol : 1.0;
#pragma unroll
    for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
      for (int lo = 0; lo < S; lo++) {
        const int base = hi * S * N1 + lo;
        double in[N1];
#pragma unroll
        for (int i = 0; i < N1; i++) in[i] = t[base + i * S];
#pragma unroll
        for (int o = 0; o < N1; o++) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < N1; i++) acc = fma(Q.M[o * N1 + i] * sc, in[i], acc);
          t[base + o * S] = acc;
        }
      }
  }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = t[i];
}

// alpha_boundary of one boundary face (convectiondiffusionfem.hh:207-275); x and r are the cell's
// local vectors in shared memory.  Rare path (non-Dirichlet boundary cells only): plain loops.
template <int DIM, int K, bool RESIDUAL>
__device__ __noinline__ void fem_cell_boundary(const DevParams& P, long long cell, const int c[3], int dir, int side,
                                               const double* x, double* r) {
  constexpr int N1 = K + 1, N = LocalSize<DIM, N1>::N;
  const long long bf = bface_index(P, c, dir, side);
  const int bctype = P.bctype ? (int)P.bctype[bf] : (int)PDB200_BC_DIRICHLET;  // face centre (:226-229)
  if (bctype == PDB200_BC_DIRICHLET || bctype == PDB200_BC_NONE) return;
  if (bctype == PDB200_BC_NEUMANN && !(RESIDUAL && P.j)) return;
  const int m = P.m;
  const double area = P.area[dir];
  for (int q = 0; q < P.nfq; q++) {
    double bv[3];
    load_b(P, cell, face_pt(P, dir, side, q), bv);  // param.b(cell_inside, local), :254
    const double bn = bv[dir] * (side ? 1.0 : -1.0);
    int pt[3] = {0, 0, 0};
    double weight = 1.0;
    {
      int qq = q;
      for (int d = 0; d < DIM; d++)
        if (d != dir) {
          pt[d] = qq % m;
          qq /= m;
          weight *= P.wq[pt[d]];
        }
    }
    pt[dir] = side ? m + 1 : m;
    const double factor = weight * area;
    double val;
    if (bctype == PDB200_BC_NEUMANN) {
      val = P.j[bf * P.nfq + q] * factor;  // :243-249
    } else {                               // Outflow :252-272
      double u = 0.0;
      for (int i = 0; i < N; i++) {
        const int i0 = i % N1, i1 = (i / N1) % N1, i2 = i / (N1 * N1);
        double phi = P.P[pt[0] * N1 + i0] * P.P[pt[1] * N1 + i1];
        if (DIM == 3) phi *= P.P[pt[2] * N1 + i2];
        u += x[i] * phi;
      }
      const double o = (RESIDUAL && P.o) ? P.o[bf * P.nfq + q] : 0.0;
      val = (bn * u + o) * factor;
    }
    for (int i = 0; i < N; i++) {
      const int i0 = i % N1, i1 = (i / N1) % N1, i2 = i / (N1 * N1);
      double phi = P.P[pt[0] * N1 + i0] * P.P[pt[1] * N1 + i1];
      if (DIM == 3) phi *= P.P[pt[2] * N1 + i2];
      r[i] += val * phi;
    }
  }
}

// KRON: Kronecker form of the cell integral (diagonal A, b = 0); the u-independent part of the
// residual then comes from r0 = R(0) (the operator is affine), which is added in step 3.
template <int DIM, int K, bool RESIDUAL, bool KRON>
__global__ void __launch_bounds__(FEM_THREADS)
    fem_vector_kernel(const DevParams P, const QkLayout L, const FemKron Q, const double* __restrict__ x,
                      double* __restrict__ y, const double* __restrict__ r0, int overwrite) {
  constexpr int N1 = K + 1, N = LocalSize<DIM, N1>::N;
  constexpr int T0 = Tile<DIM>::T0, T1 = Tile<DIM>::T1, T2 = Tile<DIM>::T2;
  constexpr int C0 = T0 + 1, C1 = T1 + 1, C2 = DIM == 3 ? T2 + 1 : 1;          // cells of the box
  constexpr int Q0 = K * C0 + 1, Q1 = K * C1 + 1, Q2 = DIM == 3 ? K * C2 + 1 : 1;  // lattice points of the box
  constexpr int NCELL = C0 * C1 * C2, NPT = Q0 * Q1 * Q2;
  extern __shared__ double smem[];
  double* xs = smem;        // [Q2][Q1][Q0]
  double* rs = smem + NPT;  // [NCELL][N]
  const int tid = threadIdx.x;
  // first cell of the box: one cell below the first owned lattice point
  const int cb[3] = {(int)blockIdx.x * T0 - 1, (int)blockIdx.y * T1 - 1, DIM == 3 ? (int)blockIdx.z * T2 - 1 : 0};
  const int Np[3] = {K * P.N[0], K * P.N[1], DIM == 3 ? K * P.N[2] : 0};  // last lattice index

  // 1. gather the lattice values of the box (loadCoefficientsLFSUInside for all its cells at once)
  for (int i = tid; i < NPT; i += FEM_THREADS) {
    const int l0 = i % Q0, l1 = (i / Q0) % Q1, l2 = i / (Q0 * Q1);
    const int p[3] = {K * cb[0] + l0, K * cb[1] + l1, DIM == 3 ? K * cb[2] + l2 : 0};
    const bool valid = p[0] >= 0 && p[0] <= Np[0] && p[1] >= 0 && p[1] <= Np[1] && p[2] >= 0 && p[2] <= Np[2];
    xs[i] = valid ? __ldg(x + qk_lattice_index(L, p)) : 0.0;
  }
  __syncthreads();

  // 2. cell-local residuals
  for (int ci = tid; ci < NCELL; ci += FEM_THREADS) {
    const int lc0 = ci % C0, lc1 = (ci / C0) % C1, lc2 = ci / (C0 * C1);
    const int c[3] = {cb[0] + lc0, cb[1] + lc1, DIM == 3 ? cb[2] + lc2 : 0};
    const bool valid = c[0] >= 0 && c[0] < P.N[0] && c[1] >= 0 && c[1] < P.N[1] && c[2] >= 0 && c[2] < P.N[2];
    if (!valid) continue;
    const long long cell = cell_index(P.N, c[0], c[1], c[2]);
    double xl[N], rl[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int i0 = i % N1, i1 = (i / N1) % N1, i2 = i / (N1 * N1);
      xl[i] = xs[(K * lc0 + i0) + Q0 * ((K * lc1 + i1) + Q1 * (K * lc2 + i2))];
    }
    if (KRON)
      fem_cell_volume_kron<DIM, K>(P, Q, cell, xl, rl);
    else
      fem_cell_volume<DIM, K, RESIDUAL>(P, cell, xl, rl);
#pragma unroll
    for (int i = 0; i < N; i++) rs[ci * N + i] = rl[i];
    // boundary faces in intersection order (default/assembler.hh:156-236); needs b, j or o data
    if (!KRON && P.bctype) {
      bool onb = false;
      for (int d = 0; d < DIM; d++) onb |= c[d] == 0 || c[d] == P.N[d] - 1;
      if (onb) {
        double xcopy[N];  // the boundary routine takes pointers: keep xl itself in registers
#pragma unroll
        for (int i = 0; i < N; i++) xcopy[i] = xl[i];
        for (int d = 0; d < DIM; d++)
          for (int side = 0; side < 2; side++) {
            const bool on = side ? c[d] == P.N[d] - 1 : c[d] == 0;
            if (!on || P.side_kind[d][side] == PDB200_SIDE_PROCESSOR) continue;
            fem_cell_boundary<DIM, K, RESIDUAL>(P, cell, c, d, side, xcopy, rs + ci * N);
          }
      }
    }
  }
  __syncthreads();

  // 3. owned lattice points: r[ci] += rl[i] over the adjacent cells in ascending cell order
  constexpr int O0 = K * T0, O1 = K * T1, O2 = DIM == 3 ? K * T2 : 1;  // the grid covers all K*N_d + 1 points
  for (int i = tid; i < O0 * O1 * O2; i += FEM_THREADS) {
    const int o[3] = {i % O0, (i / O0) % O1, i / (O0 * O1)};
    int p[3] = {0, 0, 0};
    bool own = true;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      p[d] = K * (cb[d] + 1) + o[d];
      own = own && p[d] <= Np[d];
    }
    if (!own) continue;
    // adjacent cells per direction: local cell index range and local DOF index of p in each
    int clo[3] = {0, 0, 0}, cnt[3] = {1, 1, 1}, li[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const int rem = p[d] % K, cd = p[d] / K;  // K = 1: rem == 0 always
      if (rem != 0) {                           // interior node of cell cd
        clo[d] = cd;
        li[d][0] = rem;
      } else {
        const bool lower = cd - 1 >= 0, upper = cd < P.N[d];
        clo[d] = lower ? cd - 1 : cd;
        cnt[d] = (lower ? 1 : 0) + (upper ? 1 : 0);
        li[d][0] = lower ? K : 0;
        li[d][1] = 0;
      }
    }
    const long long gi = qk_lattice_index(L, p);
    double v = overwrite ? 0.0 : y[gi];
    if (KRON && r0) v += r0[gi];
    for (int a2 = 0; a2 < cnt[2]; a2++)
      for (int a1 = 0; a1 < cnt[1]; a1++)
        for (int a0 = 0; a0 < cnt[0]; a0++) {
          const int lc0 = clo[0] + a0 - cb[0], lc1 = clo[1] + a1 - cb[1], lc2 = DIM == 3 ? clo[2] + a2 - cb[2] : 0;
          const int ci = lc0 + C0 * (lc1 + C1 * lc2);
          const int il = li[0][a0] + N1 * (li[1][a1] + N1 * (DIM == 3 ? li[2][a2] : 0));
          v += rs[ci * N + il];
        }
    y[gi] = v;
  }
}

__global__ void constrain_kernel(double* __restrict__ y, const uint64_t* __restrict__ idx, long long n, double value) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[idx[i]] = value;
}

template <int DIM, int K, bool RESIDUAL, bool KRON>
void launch_fem_variant(const FemPlan* plan, const DevParams& P, const double* x, double* y, const double* r0,
                        bool overwrite, cudaStream_t s) {
  constexpr int N1 = K + 1, N = LocalSize<DIM, N1>::N;
  constexpr int T0 = Tile<DIM>::T0, T1 = Tile<DIM>::T1, T2 = Tile<DIM>::T2;
  constexpr int C0 = T0 + 1, C1 = T1 + 1, C2 = DIM == 3 ? T2 + 1 : 1;
  constexpr int NPT = (K * C0 + 1) * (K * C1 + 1) * (DIM == 3 ? K * C2 + 1 : 1);
  constexpr size_t smem = (size_t)(NPT + C0 * C1 * C2 * N) * sizeof(double);
  // lattice points per direction: K*N_d + 1, K*T_d owned per tile
  dim3 grid((K * P.N[0] + 1 + K * T0 - 1) / (K * T0), (K * P.N[1] + 1 + K * T1 - 1) / (K * T1),
            DIM == 3 ? (K * P.N[2] + 1 + K * T2 - 1) / (K * T2) : 1);
  PDB_CUDA(cudaFuncSetAttribute(fem_vector_kernel<DIM, K, RESIDUAL, KRON>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  fem_vector_kernel<DIM, K, RESIDUAL, KRON><<<grid, FEM_THREADS, smem, s>>>(P, plan->L, plan->kron, x, y, r0,
                                                                             overwrite ? 1 : 0);
  PDB_CUDA(cudaGetLastError());
}

template <int DIM, int K>
void launch_fem(FemPlan* plan, const DevParams& P, const double* x, double* y, bool residual, bool overwrite,
                cudaStream_t s) {
  const bool kron = kron_coefficients(P);
  plan->fused_constraints = false;
  if (!kron) {
    if (residual)
      launch_fem_variant<DIM, K, true, false>(plan, P, x, y, nullptr, overwrite, s);
    else
      launch_fem_variant<DIM, K, false, false>(plan, P, x, y, nullptr, overwrite, s);
    return;
  }
  const double* r0 = nullptr;
  if (residual) {
    // R(x) = J x + R(0): the source term and the Neumann / outflow data (convectiondiffusionfem.hh:134,
    // 243-272) do not depend on x; evaluate them once with the quadrature kernel and cache
    if (!plan->r0_valid) {
      if (!plan->r0) PDB_CUDA(cudaMalloc(&plan->r0, (size_t)P.ndofs * sizeof(double)));
      double* zero = nullptr;
      PDB_CUDA(cudaMalloc(&zero, (size_t)P.ndofs * sizeof(double)));
      PDB_CUDA(cudaMemsetAsync(zero, 0, (size_t)P.ndofs * sizeof(double), s));
      launch_fem_variant<DIM, K, true, false>(plan, P, zero, plan->r0, nullptr, true, s);
      PDB_CUDA(cudaStreamSynchronize(s));
      PDB_CUDA(cudaFree(zero));
      plan->r0_valid = true;
    }
    r0 = plan->r0;
  }
  // all boundary lattice points constrained (no bctype array): the kernel writes the zero rows itself
  if (P.ndofs >= (1ll << 31)) {  // fem_kron.cu keeps container indices in 32 bits
    launch_fem_variant<DIM, K, false, true>(plan, P, x, y, r0, overwrite, s);
    return;
  }
  plan->fused_constraints = P.bctype == nullptr;
  launch_fem_kron(P, plan->L, plan->kron.MinvK, plan->kron.M, x, y, r0, overwrite, plan->fused_constraints, false, s);
}

}  // namespace

FemPlan* fem_plan_create(const DevParams& P, const int8_t* bctype_dev, const Kron1D& K1) {
  FemPlan* plan = new FemPlan;
  plan->L = make_qk_layout(P);
  for (int i = 0; i < P.n1; i++)
    for (int j = 0; j < P.n1; j++) {
      plan->kron.MinvK[i * P.n1 + j] = K1.MinvK[i * MAX_N1 + j];
      plan->kron.M[i * P.n1 + j] = K1.M[i * MAX_N1 + j];
    }
  std::vector<int8_t> bct;
  if (bctype_dev) {
    long long nbf = 0;
    for (int d = 0; d < P.dim; d++) nbf += 2 * (P.ncells / P.N[d]);
    bct.resize(nbf);
    PDB_CUDA(cudaMemcpy(bct.data(), bctype_dev, bct.size(), cudaMemcpyDeviceToHost));
  }
  std::vector<uint64_t> list = host_constrained_dofs(P, bct.empty() ? nullptr : bct.data());
  plan->ncon = (long long)list.size();
  if (plan->ncon) {
    PDB_CUDA(cudaMalloc(&plan->con, list.size() * sizeof(uint64_t)));
    PDB_CUDA(cudaMemcpy(plan->con, list.data(), list.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  }
  return plan;
}

void fem_plan_destroy(FemPlan* p) {
  if (!p) return;
  if (p->con) cudaFree(p->con);
  if (p->r0) cudaFree(p->r0);
  delete p;
}

const QkLayout& fem_plan_layout(const FemPlan* p) { return p->L; }
void fem_plan_invalidate(FemPlan* p) {
  if (p) p->r0_valid = false;
}
const uint64_t* fem_plan_constrained(const FemPlan* p, long long* n) {
  *n = p->ncon;
  return p->con;
}

// d = point diagonal of the Jacobian (PointDiagonalLocalOperatorWrapper, localoperator/pointdiagonalwrapper.hh),
// with 1 on the constrained rows (the unit rows of set_trivial_rows, assemblerutilities.hh:666-684)
void launch_fem_diagonal(FemPlan* plan, const DevParams& P, double* d, cudaStream_t s) {
  if (!kron_coefficients(P) || P.ndofs >= (1ll << 31))
    throw Error("matrix-free point diagonal: needs a diagonal tensor and b = 0");
  if (P.m != P.k + 1) throw Error("conforming Qk kernel: intorderadd must be 0 or 1 (k+1 Gauss points)");
  const bool fused = P.bctype == nullptr;
  launch_fem_kron(P, plan->L, plan->kron.MinvK, plan->kron.M, d, d, nullptr, true, fused, true, s);
  if (plan->ncon && !fused) {
    constrain_kernel<<<(unsigned)((plan->ncon + 255) / 256), 256, 0, s>>>(d, plan->con, plan->ncon, 1.0);
    PDB_CUDA(cudaGetLastError());
  }
}

void launch_fem_vector(FemPlan* plan, const DevParams& P, const double* x, double* y, bool residual, bool overwrite,
                       cudaStream_t s) {
  if (P.m != P.k + 1) throw Error("conforming Qk kernel: intorderadd must be 0 or 1 (k+1 Gauss points)");
  if (P.dim == 2 && P.k == 1) launch_fem<2, 1>(plan, P, x, y, residual, overwrite, s);
  else if (P.dim == 2 && P.k == 2) launch_fem<2, 2>(plan, P, x, y, residual, overwrite, s);
  else if (P.dim == 3 && P.k == 1) launch_fem<3, 1>(plan, P, x, y, residual, overwrite, s);
  else if (P.dim == 3 && P.k == 2) launch_fem<3, 2>(plan, P, x, y, residual, overwrite, s);
  else throw Error("conforming Qk kernel: unsupported (dim, degree)");
  // postAssembly: constrain_residual (residualengine.hh:228-233, jacobianapplyengine.hh:249-254)
  if (plan->ncon && !plan->fused_constraints) {
    constrain_kernel<<<(unsigned)((plan->ncon + 255) / 256), 256, 0, s>>>(y, plan->con, plan->ncon, 0.0);
    PDB_CUDA(cudaGetLastError());
  }
}

}  // namespace pdb

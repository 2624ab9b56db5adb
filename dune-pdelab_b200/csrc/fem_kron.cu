// fem_kron.cu — conforming Qk (k = 1, 2; dim = 2, 3) matrix-free  y (+)= J x (+ R(0))  for
// ConvectionDiffusionFEM with a cell-wise constant DIAGONAL tensor and b = 0: the fast path behind
// GridOperator::residual / jacobian_apply (gridoperator/gridoperator.hh:176-197).
//
// What it replaces in the reference (paths relative to /root/reference/dune/pdelab/):
//   DefaultAssembler::assemble cell loop                 gridoperator/default/assembler.hh:85-279
//   LFSIndexCache gather / scatter                        gridfunctionspace/lfsindexcache.hh:603-633,
//                                                         gridoperator/default/residualengine.hh:131-233
//   ConvectionDiffusionFEM::alpha_volume                  localoperator/convectiondiffusionfem.hh:63-136
//   constrain_residual                                    constraints/common/constraints.hh:904-915
//
// Cell integral.  The (k+1)-point Gauss rule of convectiondiffusionfem.hh:93-94 integrates the
// products of 1-D polynomials exactly, so alpha_volume of cell e equals
//     r_e = |K| (M (x) M (x) M) [ sum_d (A_dd / h_d^2) (M^-1 K)_d x_e + c_e x_e ]
// (M, K the exact 1-D mass / stiffness matrices): one 1-D sweep per direction and three mass
// sweeps (integer matrices 30 M resp. 6 M, scale folded into the coefficients).
//
// Mapping to the machine.  One thread per cell; the scatter r[ci(i)] += rl[i] of the reference is
// turned into a race-free, atomic-free assembly along the three grid directions:
//   x: the 32 lanes of a warp are 32 consecutive cells of an x-row; the face  i_0 = k  of a cell is
//      handed to the right neighbour with warp shuffles;
//   y: the warps of a CTA are consecutive x-rows; face  i_1 = k  goes to the next warp through a
//      double-buffered shared-memory slot (one __syncthreads per step);
//   z: a thread marches along z; face  i_2 = k  is carried in registers to the next step, and so is
//      the plane of input values it shares with the next cell.
// After the three hand-overs a thread holds the finished rows of the k^dim lattice points its cell
// owns (local indices < k) and writes each exactly once; along x consecutive lanes write
// consecutive addresses of one sub-entity group of the container (host_tables.h: QkLayout), so all
// global traffic is coalesced.  Lane 0 / warp 0 / step 0 of a tile are the overlap cell layer
// (their own rows are written by the neighbouring tile).  Deterministic: fixed summation order.
// In 2-D the march runs along y and there is no shared-memory stage.

#include <algorithm>

#include "common.cuh"
#include "host_tables.h"

namespace pdb {

namespace {

struct FemKronParams {
  double s[3];     // |K| / D^dim / h_d^2   (D = 30 for k = 2, 6 for k = 1)
  double sc;       // |K| / D^dim
  double MinvK[MAX_N1 * MAX_N1];  // row-major, leading dimension n1
  long long goff[8];              // first container index of sub-entity group s (k = 2); goff[0] = 0 for k = 1
  int gd0[8], gd1[8];             // entities per x-row / y-column of group s
  int d2[8];                      // gd0 * gd1 (3-D): index stride of the group along z
  int N[3];
  int chunk;       // cells per march chunk
  int tiles_x;     // 2-D: warps of a CTA are independent x-tiles
  int fuse_constraints;  // every boundary lattice point is constrained (no bctype array): write 0 there
  // diag != 0: write the POINT DIAGONAL of the Jacobian instead of J x (PointDiagonalLocalOperatorWrapper,
  // localoperator/pointdiagonalwrapper.hh): the cell contribution is the closed form
  //   t_i = sum_d al_d (D K)_{i_d i_d} prod_{d' != d} (D M)_{i_d' i_d'} + cc prod_d (D M)_{i_d i_d},
  // assembled over the cells by the same hand-overs; constrained rows get 1 (unit rows of the matrix)
  int diag;
  double DK[MAX_N1], DM[MAX_N1];  // diagonals of D K and D M (1-D stiffness / mass, D = 30 or 6)
};

template <int DIM, int K>
struct KL {
  static constexpr int N1 = K + 1;
  static constexpr int N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  static constexpr int MD = DIM - 1;                        // march direction
  static constexpr int SM = DIM == 3 ? N1 * N1 : N1;        // local stride of the march direction
};

// Container indices (closed form of LFSIndexCache, host_tables.h) in 32-bit arithmetic: a thread
// keeps one running index per sub-entity group (the index of the group entity anchored at its
// cell); a local DOF (i0, i1, i2) is that index plus a warp-uniform delta.
template <int DIM, int K>
struct Idx {
  static constexpr int NG = K == 1 ? 1 : (1 << DIM);
  int base[NG];
  __device__ __forceinline__ void init(const FemKronParams& F, int c0, int c1, int c2) {
#pragma unroll
    for (int g = 0; g < NG; g++)
      base[g] = (int)F.goff[g] + c0 + F.gd0[g] * c1 + (DIM == 3 ? F.d2[g] * c2 : 0);
  }
  __device__ __forceinline__ void advance(const FemKronParams& F) {
#pragma unroll
    for (int g = 0; g < NG; g++) base[g] += DIM == 3 ? F.d2[g] : F.gd0[g];
  }
  __device__ __forceinline__ long long at(const FemKronParams& F, int i0, int i1, int i2) const {
    if (K == 1) return base[0] + i0 + F.gd0[0] * i1 + (DIM == 3 ? F.d2[0] * i2 : 0);
    const int g = (i0 & 1) | ((i1 & 1) << 1) | (DIM == 3 ? (i2 & 1) << 2 : 0);
    return base[g] + (i0 >> 1) + F.gd0[g] * (i1 >> 1) + (DIM == 3 ? F.d2[g] * (i2 >> 1) : 0);
  }
};

// v <- (D M along stride S) v for every line: 30 M = [[4,2,-1],[2,16,2],[-1,2,4]], 6 M = [[2,1],[1,2]]
template <int DIM, int K, int AXIS>
__device__ __forceinline__ void mass_sweep(double (&t)[KL<DIM, K>::N]) {
  constexpr int N1 = K + 1, N = KL<DIM, K>::N;
  constexpr int S = AXIS == 0 ? 1 : (AXIS == 1 ? N1 : N1 * N1);
#pragma unroll
  for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
    for (int lo = 0; lo < S; lo++) {
      const int base = hi * S * N1 + lo;
      if (K == 2) {
        const double v0 = t[base], v1 = t[base + S], v2 = t[base + 2 * S];
        const double e = v0 + v2;
        const double w = fma(2.0, v1, -e);
        t[base] = fma(5.0, v0, w);
        t[base + S] = fma(16.0, v1, e + e);
        t[base + 2 * S] = fma(5.0, v2, w);
      } else {
        const double e = t[base] + t[base + S];
        t[base] += e;
        t[base + S] += e;
      }
    }
}

// t (+)= al (M^-1 K along AXIS) x
template <int DIM, int K, int AXIS, bool ACC>
__device__ __forceinline__ void stiff_sweep(const FemKronParams& F, double al, const double (&x)[KL<DIM, K>::N],
                                            double (&t)[KL<DIM, K>::N]) {
  constexpr int N1 = K + 1, N = KL<DIM, K>::N;
  constexpr int S = AXIS == 0 ? 1 : (AXIS == 1 ? N1 : N1 * N1);
  double W[N1 * N1];
#pragma unroll
  for (int i = 0; i < N1 * N1; i++) W[i] = al * F.MinvK[i];
#pragma unroll
  for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
    for (int lo = 0; lo < S; lo++) {
      const int base = hi * S * N1 + lo;
#pragma unroll
      for (int o = 0; o < N1; o++) {
        double acc = ACC ? t[base + o * S] : 0.0;
#pragma unroll
        for (int i = 0; i < N1; i++) acc = fma(W[o * N1 + i], x[base + i * S], acc);
        t[base + o * S] = acc;
      }
    }
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

template <int DIM, int K>
struct KS {  // per-thread staging slots in shared memory (values per thread)
  static constexpr int NLD = KL<DIM, K>::N - KL<DIM, K>::SM;  // input planes i_m = 1..k
  static constexpr int KO = DIM == 3 ? K * K * K : K * K;      // owned lattice points
  static constexpr int NCO = 4;                                // A_dd (<= 3) and c
  static constexpr int PER_THREAD = NLD + NCO + 2 * KO;
};

// The loads of step s+1 (input planes, coefficients: group X; old y and R(0) of the rows to be
// written: group Y) are issued with cp.async into per-thread shared-memory slots while step s is
// computed, so no global-memory latency is exposed inside the march.
template <int DIM, int K, int WY, int MINB>
__global__ void __launch_bounds__(32 * WY, MINB)
    fem_kron_kernel(const DevParams P, const FemKronParams F, const double* __restrict__ xg, double* __restrict__ yg,
                    const double* __restrict__ r0, int overwrite) {
  using L = KL<DIM, K>;
  using S = KS<DIM, K>;
  constexpr int N1 = L::N1, N = L::N, SM = L::SM, NT = 32 * WY;
  constexpr int NYX = DIM == 3 ? K * N1 : 1;  // values handed to the next warp per cell
  constexpr int NSLOT = DIM == 3 ? 2 * WY * NYX * 32 : 0;
  extern __shared__ double dyn[];
  double* slot = dyn;
  const int tid = threadIdx.x;
  double* xs = dyn + NSLOT + tid;        // [NLD][NT]
  double* co = xs + S::NLD * NT;         // [NCO][NT]
  double* ys = co + S::NCO * NT;         // [KO][NT]   old y
  double* rs = ys + S::KO * NT;          // [KO][NT]   R(0)
  const int lane = tid & 31, warp = tid >> 5;
  // cell coordinates: lane 0 / warp 0 / step 0 are the overlap layer of the tile
  int c0, c1 = 0, cm0;  // cm0: first march coordinate (overlap layer)
  if (DIM == 3) {
    c0 = (int)blockIdx.x * 31 - 1 + lane;
    c1 = (int)blockIdx.y * (WY - 1) - 1 + warp;
    cm0 = (int)blockIdx.z * F.chunk - 1;
  } else {
    const int tile = (int)blockIdx.x * WY + warp;
    c0 = tile * 31 - 1 + lane;
    cm0 = (int)blockIdx.y * F.chunk - 1;
    if (tile >= F.tiles_x) return;  // whole warp; no block-wide barrier in 2-D
  }
  const int Nm = F.N[L::MD];
  const bool vx = c0 >= 0 && c0 < F.N[0];
  const bool vy = DIM == 3 ? (c1 >= 0 && c1 < F.N[1]) : true;
  // rows of this thread that exist: the cell layer c_d == N_d owns only the lattice layer i_d == 0
  const bool ownx = lane > 0 && c0 <= F.N[0];
  const bool owny = DIM == 3 ? (warp > 0 && c1 <= F.N[1]) : true;
  const bool lastx = c0 == F.N[0], lasty = DIM == 3 && c1 == F.N[1];
  const bool need_y = !overwrite, need_r = r0 != nullptr;
  const int steps = min(F.chunk, Nm + 1 - (cm0 + 1)) + 1;
  Idx<DIM, K> cur, nxt;  // container indices anchored at the cell of this step / the next step
  cur.init(F, c0, DIM == 3 ? c1 : cm0, DIM == 3 ? cm0 : 0);
  nxt = cur;
  int cell_nxt = c0 + F.N[0] * ((DIM == 3 ? c1 : cm0) + (DIM == 3 ? F.N[1] * cm0 : 0));
  const int cell_stride = DIM == 3 ? F.N[0] * F.N[1] : F.N[0];

  // group X of a step: input planes 1..k and the coefficients of the cell
  // (nxt / cell_nxt address the cell of `step` when these are called)
  auto issue_x = [&](int step) {
    const int cm = cm0 + step;
    if (step < steps && vx && vy && cm >= 0 && cm < Nm) {
      if (!F.diag) {
#pragma unroll
        for (int i = SM; i < N; i++) {
          const int i0 = i % N1, i1 = (i / N1) % N1, i2 = i / (N1 * N1);
          cp_async8(xs + (i - SM) * NT, xg + nxt.at(F, i0, i1, i2));
        }
      }
      const long long cell = cell_nxt;
      if (P.a_mode == PDB200_A_SCALAR) {
        cp_async8(co, P.A + cell);
      } else if (P.a_mode == PDB200_A_DIAGONAL) {
#pragma unroll
        for (int d = 0; d < DIM; d++) cp_async8(co + d * NT, P.A + cell * DIM + d);
      }
      if (P.c) cp_async8(co + 3 * NT, P.c + cell);
    }
    cp_async_commit();
  };
  // group Y of a step: old y / R(0) of the rows the step will write
  auto issue_y = [&](int step) {
    const int cm = cm0 + step;
    if ((need_y || need_r) && step > 0 && step < steps && ownx && owny) {
      const bool lastm = cm == Nm;
#pragma unroll
      for (int j = 0; j < S::KO; j++) {
        const int i0 = j % K, i1 = (j / K) % K, i2 = DIM == 3 ? j / (K * K) : 0;
        const int im = DIM == 3 ? i2 : i1;
        if ((lastx && i0) || (lasty && i1) || (lastm && im)) continue;
        const long long gi = nxt.at(F, i0, i1, i2);
        if (need_y) cp_async8(ys + j * NT, yg + gi);
        if (need_r) cp_async8(rs + j * NT, r0 + gi);
      }
    }
    cp_async_commit();
  };

  double x[N], t[N];
  double carry[DIM == 3 ? K * K : K];
#pragma unroll
  for (int i = 0; i < (DIM == 3 ? K * K : K); i++) carry[i] = 0.0;
  bool have_plane = false;  // x[.., i_m = k] of the previous step is this step's plane i_m = 0
  issue_x(0);
  issue_y(0);

  for (int step = 0; step < steps; step++) {
    const int cm = cm0 + step;
    const bool valid = vx && vy && cm >= 0 && cm < Nm;
    cp_async_wait<1>();  // group X of this step has landed (group Y may still be in flight)
    double a[3] = {1.0, 1.0, 1.0}, cc = 0.0;
    if (valid) {
      // ---- gather (loadCoefficientsLFSUInside); the plane shared with the previous cell is carried
#pragma unroll
      for (int i = 0; i < SM; i++) {
        const int i0 = i % N1, i1 = (i / N1) % N1;
        if (have_plane)
          x[i] = x[i + K * SM];
        else
          x[i] = F.diag ? 0.0 : __ldg(xg + cur.at(F, i0, DIM == 3 ? i1 : 0, 0));
      }
#pragma unroll
      for (int i = SM; i < N; i++) x[i] = xs[(i - SM) * NT];
      if (P.a_mode == PDB200_A_SCALAR) {
        a[0] = a[1] = a[2] = co[0];
      } else if (P.a_mode == PDB200_A_DIAGONAL) {
#pragma unroll
        for (int d = 0; d < DIM; d++) a[d] = co[d * NT];
      }
      if (P.c) cc = co[3 * NT] * F.sc;
    }
    nxt.advance(F);
    cell_nxt += cell_stride;
    issue_x(step + 1);  // the slots were just read by their only user: refill them behind the compute
    if (valid && F.diag) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int id[3] = {i % N1, (i / N1) % N1, i / (N1 * N1)};
        double m = cc, v = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          double pd = a[d] * F.s[d] * F.DK[id[d]];
#pragma unroll
          for (int e = 0; e < DIM; e++)
            if (e != d) pd *= F.DM[id[e]];
          v += pd;
          m *= F.DM[id[d]];
        }
        t[i] = v + m;
      }
    } else if (valid) {
      stiff_sweep<DIM, K, 0, false>(F, a[0] * F.s[0], x, t);
      stiff_sweep<DIM, K, 1, true>(F, a[1] * F.s[1], x, t);
      if (DIM == 3) stiff_sweep<DIM, K, 2, true>(F, a[2] * F.s[2], x, t);
      if (P.c) {
#pragma unroll
        for (int i = 0; i < N; i++) t[i] = fma(cc, x[i], t[i]);
      }
      mass_sweep<DIM, K, 0>(t);
      mass_sweep<DIM, K, 1>(t);
      if (DIM == 3) mass_sweep<DIM, K, 2>(t);
    } else {
#pragma unroll
      for (int i = 0; i < N; i++) t[i] = 0.0;
    }
    have_plane = valid;  // uniform along the march for a thread with vx && vy; others never load

    // ---- x: face i_0 = k -> right neighbour's i_0 = 0 ------------------------------------------
#pragma unroll
    for (int j = 0; j < N / N1; j++) {
      const double up = __shfl_up_sync(0xffffffffu, t[j * N1 + K], 1);
      if (lane > 0) t[j * N1] += up;
    }
    // ---- y: face i_1 = k (i_0 < k) -> next warp's i_1 = 0 ---------------------------------------
    if (DIM == 3) {
      double* buf = slot + (step & 1) * (WY * NYX * 32);
#pragma unroll
      for (int i2 = 0; i2 < N1; i2++)
#pragma unroll
        for (int i0 = 0; i0 < K; i0++) buf[(warp * NYX + i2 * K + i0) * 32 + lane] = t[i0 + N1 * (K + N1 * i2)];
      __syncthreads();
      if (warp > 0) {
#pragma unroll
        for (int i2 = 0; i2 < N1; i2++)
#pragma unroll
          for (int i0 = 0; i0 < K; i0++) t[i0 + N1 * N1 * i2] += buf[((warp - 1) * NYX + i2 * K + i0) * 32 + lane];
      }
    }
    // ---- march direction: face i_m = k carried to the next step ---------------------------------
#pragma unroll
    for (int j = 0; j < (DIM == 3 ? K * K : K); j++) {
      const int i0 = j % K, i1 = DIM == 3 ? j / K : 0;
      const int lo = DIM == 3 ? i0 + N1 * i1 : i0;
      t[lo] += carry[j];
      carry[j] = t[lo + K * SM];
    }
    // ---- the k^dim lattice points this cell owns are complete: write each row once --------------
    cp_async_wait<1>();  // group Y of this step (issued one step ago) has landed
    if (step > 0 && ownx && owny) {
      const bool lastm = cm == Nm;
#pragma unroll
      for (int j = 0; j < S::KO; j++) {
        const int i0 = j % K, i1 = (j / K) % K, i2 = DIM == 3 ? j / (K * K) : 0;
        const int im = DIM == 3 ? i2 : i1;
        if ((lastx && i0) || (lasty && i1) || (lastm && im)) continue;
        const long long gi = cur.at(F, i0, i1, i2);
        double v = t[i0 + N1 * (i1 + N1 * i2)];
        if (F.fuse_constraints) {
          // constrain_residual: all boundary lattice points are Dirichlet- or processor-constrained
          const bool onb = (i0 == 0 && (c0 == 0 || lastx)) || (DIM == 3 && i1 == 0 && (c1 == 0 || lasty)) ||
                           (im == 0 && (cm == 0 || lastm));
          if (onb) {
            yg[gi] = F.diag ? 1.0 : 0.0;
            continue;
          }
        }
        if (need_r) v += rs[j * NT];
        if (need_y) v += ys[j * NT];
        yg[gi] = v;
      }
    }
    issue_y(step + 1);
    cur = nxt;
  }
  cp_async_wait<0>();
}

template <int DIM, int K, int WY, int MINB>
void launch_variant(const DevParams& P, const QkLayout& Lq, const double* MinvK, const double* M1, const double* x,
                    double* y, const double* r0, bool overwrite, bool fuse_constraints, bool diag, cudaStream_t s) {
  FemKronParams F;
  const double D = K == 2 ? 30.0 : 6.0;
  F.diag = diag ? 1 : 0;
  for (int i = 0; i < MAX_N1; i++) F.DK[i] = F.DM[i] = 0.0;
  for (int i = 0; i <= K; i++) {  // (M M^-1 K)_ii = K_ii
    double kii = 0.0;
    for (int j = 0; j <= K; j++) kii += M1[i * (K + 1) + j] * MinvK[j * (K + 1) + i];
    F.DK[i] = D * kii;
    F.DM[i] = D * M1[i * (K + 1) + i];
  }
  double sc = P.vol;
  for (int d = 0; d < DIM; d++) sc /= D;
  F.sc = sc;
  for (int d = 0; d < 3; d++) {
    F.s[d] = d < DIM ? sc / (P.h[d] * P.h[d]) : 0.0;
    F.N[d] = d < DIM ? P.N[d] : 1;
  }
  for (int i = 0; i < MAX_N1 * MAX_N1; i++) F.MinvK[i] = i < (K + 1) * (K + 1) ? MinvK[i] : 0.0;
  for (int g = 0; g < 8; g++) {
    int edim = 0;
    for (int d = 0; d < DIM; d++) edim += (g >> d) & 1;
    F.goff[g] = K == 1 ? 0 : (g < (1 << DIM) ? Lq.block_off[edim] + Lq.group_off[g] : 0);
    F.gd0[g] = K == 1 ? P.N[0] + 1 : ((g & 1) ? P.N[0] : P.N[0] + 1);
    F.gd1[g] = K == 1 ? P.N[1] + 1 : ((g & 2) ? P.N[1] : P.N[1] + 1);
    F.d2[g] = F.gd0[g] * F.gd1[g];
  }
  F.fuse_constraints = fuse_constraints ? 1 : 0;
  const int tiles_x = (P.N[0] + 1 + 30) / 31;
  F.tiles_x = tiles_x;
  const int Nm = P.N[DIM - 1] + 1;  // cell layers along the march direction incl. the closing one
  const long long columns = DIM == 3 ? (long long)tiles_x * ((P.N[1] + 1 + WY - 2) / (WY - 1)) : (tiles_x + WY - 1) / WY;
  // chunks along the march direction: every chunk repeats one overlap layer, and the CTAs should fill
  // whole waves of the 148 SMs; pick the count with the least estimated time (waves x steps)
  int best = 1;
  double best_cost = 1e300;
  const int max_chunks = std::max(1, Nm / 8);
  for (int ch = 1; ch <= max_chunks; ch++) {
    const int len = (Nm + ch - 1) / ch;
    const long long ctas = columns * ((Nm + len - 1) / len);
    const double waves = (double)((ctas + 147) / 148);
    const double cost = waves * (len + 1);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = ch;
    }
  }
  F.chunk = (Nm + best - 1) / best;
  const int chunks = (Nm + F.chunk - 1) / F.chunk;
  dim3 grid;
  if (DIM == 3)
    grid = dim3(tiles_x, (P.N[1] + 1 + WY - 2) / (WY - 1), chunks);
  else
    grid = dim3((tiles_x + WY - 1) / WY, chunks, 1);
  constexpr size_t smem = ((DIM == 3 ? 2 * WY * K * (K + 1) * 32 : 0) + (size_t)KS<DIM, K>::PER_THREAD * 32 * WY) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    PDB_CUDA(cudaFuncSetAttribute(fem_kron_kernel<DIM, K, WY, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  fem_kron_kernel<DIM, K, WY, MINB><<<grid, 32 * WY, smem, s>>>(P, F, x, y, r0, overwrite ? 1 : 0);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace

// MinvK: M^-1 K of the 1-D Lagrange basis, row-major with leading dimension k+1
void launch_fem_kron(const DevParams& P, const QkLayout& L, const double* K1, const double* M1, const double* x, double* y,
                     const double* r0, bool overwrite, bool fuse_constraints, bool diag, cudaStream_t s) {
  if (P.dim == 2 && P.k == 1) launch_variant<2, 1, 4, 8>(P, L, K1, M1, x, y, r0, overwrite, fuse_constraints, diag, s);
  else if (P.dim == 2 && P.k == 2) launch_variant<2, 2, 4, 4>(P, L, K1, M1, x, y, r0, overwrite, fuse_constraints, diag, s);
  else if (P.dim == 3 && P.k == 1) launch_variant<3, 1, 16, 2>(P, L, K1, M1, x, y, r0, overwrite, fuse_constraints, diag, s);
  else if (P.dim == 3 && P.k == 2) launch_variant<3, 2, 12, 1>(P, L, K1, M1, x, y, r0, overwrite, fuse_constraints, diag, s);
  else throw Error("conforming Qk Kronecker kernel: unsupported (dim, degree)");
}

}  // namespace pdb

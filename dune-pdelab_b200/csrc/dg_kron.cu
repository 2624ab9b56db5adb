// dg_kron.cu — Kronecker-factorised jacobian_apply (and, with a cached R(0), residual) for QkDG of
// higher degree in 3-D: the kernel of BASELINE.json config 3 (k = 4, 64^3 cells).
//
// Same operator identity as dg_fast.cu, valid for cell-wise constant DIAGONAL
// diffusion tensors, b = 0, any k:
//     y_e = |K| (M (x) M (x) M) [ sum_d M^-1 L_d(z_{e-d}, z_e, z_{e+d}) / h_d + c_e z_e ]
// i.e. exactly GridOperator::jacobian_apply for ConvectionDiffusionDG
// (localoperator/convectiondiffusiondg.hh:106-188, 271-471, 684-879) because the reference's
// (k+1)-point Gauss rule integrates every integrand exactly.  Per direction and per line of
// n1 = k+1 nodes the 1-D operator is
//     t_i += sum_j T_ij o_j + PL1_i (d1.l) + PL2_i l_k + PR1_i (d0.r) + PR2_i r_0
// with o, l, r the line of the cell and of its lower / upper neighbour and
//     T   = A0 M^-1 K + m0 (csL d0^T + cgL e_0^T) + ctL q0 e_0^T - mk (csR d1^T - cgR e_k^T) + ctR q1 e_k^T
//     PL1 = m0 coL, PL2 = -(m0 cgL + q0 ctL), PR1 = -mk coR, PR2 = -(mk cgR + q1 ctR)
// (m0, mk = first / last column of M^-1, q0 = M^-1 d0, q1 = M^-1 d1, d0/d1 = basis derivatives at
// 0/1; cs, co, cg, ct from the harmonic weights and the penalty, convectiondiffusiondg.hh:326-346).
//
// Mapping to the machine.  A cell has n = n1^3 DOFs (125 for k = 4), too many accumulators for one
// thread, so n1 threads share a cell, each owning one PLANE of n1 x n1 nodes:
//   * x- and y-sweeps run in "z-plane layout" (thread s owns nodes (.,.,iz=s): x- and y-lines local),
//   * the z-sweep runs in "y-plane layout" (thread s owns (.,iy=s,.): z-lines local),
//   * the two partial sums are combined, and M (x) M (x) M applied, by passing the accumulators
//     through shared memory once the input tile is no longer needed.
// 32 / n1 cells per warp; CTA = TX x TY x TZ cells; tile + face halo are brought in by TMA exactly as
// in dg_fast.cu ([2 cells][Nx/2][Ny][Nz] view, out-of-range cells zero-filled), the result leaves
// through one TMA store.  Per-cell 1-D matrices T^x, T^y, T^z are built cooperatively (thread s
// builds row s) into a shared-memory coefficient block.

#include <cuda.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace pdb {

namespace {

template <int K>
struct KronCfg;
template <>
struct KronCfg<4> {
  // 24 cells = 4 warps of 6 cells, 105 KB of shared memory: TWO CTAs per SM, so that one CTA's tile load / result
  // store runs under the other's sweeps (the 4 x 4 x 3 tile of round 1 filled the SM with ONE CTA of 152 KB: a third of
  // the tile time was load and store latency with nothing to overlap it)
  // A warp (6 cells x 5 planes) is ONE x-row of the tile: six cells stored 125 doubles apart put the 15 + 14 lanes of
  // the two half-warps on distinct bank pairs in the x- and y-sweeps (tools/smem_bank_sim.py: 0 % replays instead of
  // 33 % with the 4 x 2 x 3 tile, whose warps straddle two rows; the z-sweep and the y-plane transposes stay at 50 %).
  // Same 24 cells, 88 cells of shared memory and two CTAs per SM as 4 x 2 x 3.
#if defined(PDB200_KRON4_BIG_TILE)
  static constexpr int TX = 4, TY = 4, TZ = 3;  // 48 cells = 8 warps of 6 cells
#elif defined(PDB200_KRON4_TILE_423)
  static constexpr int TX = 4, TY = 2, TZ = 3;
#else
  static constexpr int TX = 6, TY = 2, TZ = 2;
#endif
};
template <>
struct KronCfg<3> {
  static constexpr int TX = 8, TY = 2, TZ = 2;  // 32 cells = 4 warps of 8 cells, 73 KB: three CTAs per SM
};

template <int K>
struct KronDims {
  static constexpr int N1 = K + 1, NLOC = N1 * N1 * N1, NPL = N1 * N1;
  static constexpr int TX = KronCfg<K>::TX, TY = KronCfg<K>::TY, TZ = KronCfg<K>::TZ;
  static constexpr int CELLS = TX * TY * TZ;
  static constexpr int CPW = 32 / N1;  // cells per warp
  static_assert(CELLS % CPW == 0, "tile must be a whole number of warps");
  static constexpr int WARPS = CELLS / CPW, THREADS = WARPS * 32;
  static constexpr int ROWX = TX + 4;
  static constexpr int al(int doubles) { return (doubles + 15) / 16 * 16; }  // TMA destinations: 128-byte aligned
  static constexpr int R0 = 0;
  static constexpr int R1 = al(R0 + TZ * TY * ROWX * NLOC);
  static constexpr int R2 = al(R1 + TZ * TX * NLOC);
  static constexpr int R3 = al(R2 + TZ * TX * NLOC);
  static constexpr int R4 = al(R3 + TY * TX * NLOC);
  static constexpr int TILE_DOUBLES = al(R4 + TY * TX * NLOC);
  // bytes the five TMA loads deliver (without the alignment padding)
  static constexpr int TX_BYTES = 8 * NLOC * (TZ * TY * ROWX + 2 * TZ * TX + 2 * TY * TX);
  // per-cell coefficient block: T^x, T^y, T^z (3 n1^2), A0[3], cs[6], co[6], creact
  static constexpr int COEF = 3 * NPL + 16;
  static constexpr int COEF0 = (TILE_DOUBLES + 15) / 16 * 16;
  static constexpr int SMEM_DOUBLES = COEF0 + CELLS * COEF;
  static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;
  static_assert((R1 * 8) % 128 == 0 && (R2 * 8) % 128 == 0 && (R3 * 8) % 128 == 0 && (R4 * 8) % 128 == 0,
                "TMA destinations must be 128-byte aligned");
  static_assert(3 * CELLS * NLOC <= TILE_DOUBLES, "staging + scratch + R(0) tile must fit into the input tile");
  static_assert((2 * CELLS * NLOC * 8) % 128 == 0, "R(0) tile is a TMA destination");
};

template <int K>
struct KronConst {
  static constexpr int N1 = K + 1;
  double MinvK[N1 * N1], M[N1 * N1], m0[N1], mk[N1], q0[N1], q1[N1], d0[N1], d1[N1];
  double ih2[3];
  double alpha_pen, theta, vol;
};

__device__ __forceinline__ uint32_t k_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void k_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void k_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void k_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "KWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra KDONE;\n\t"
      "bra KWAIT_LOOP;\n\t"
      "KDONE:\n\t"
      "}" ::"r"(k_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void k_tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          k_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(k_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void k_tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2,
                                                    int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   map),
               "r"(k_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void k_tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(k_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void k_tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ double k_fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, fma(e, e, e), y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

__device__ __forceinline__ double k_load_adiag(const DevParams& P, int cell, int d) {
  if (P.a_mode == PDB200_A_IDENTITY) return 1.0;
  if (P.a_mode == PDB200_A_SCALAR) return __ldg(P.A + cell);
  if (P.a_mode == PDB200_A_DIAGONAL) return __ldg(P.A + (long long)cell * 3 + d);
  return __ldg(P.A + (long long)cell * 9 + d * 4);
}

// One sweep of the 1-D operator over the N1 lines of a plane.
//   o, l, r: base addresses (doubles) of the plane in the own / lower / upper cell
//   SL: stride between consecutive nodes of a line, SP: stride between the lines of the plane
//   t[line][i] (+)= ...
template <int K, int SL, int SP, bool FIRST>
__device__ __forceinline__ void kron_sweep(const double* __restrict__ o, const double* __restrict__ l,
                                           const double* __restrict__ r, const double* __restrict__ Tm,
                                           const KronConst<K>& C, double coL, double cgL, double ctL, double coR,
                                           double cgR, double ctR, double creact, double (&t)[(K + 1) * (K + 1)]) {
  constexpr int N1 = K + 1;
  double T[N1 * N1], PL1[N1], PL2[N1], PR1[N1], PR2[N1];
#pragma unroll
  for (int i = 0; i < N1 * N1; i++) T[i] = Tm[i];
#pragma unroll
  for (int i = 0; i < N1; i++) {
    PL1[i] = C.m0[i] * coL;
    PL2[i] = -fma(C.m0[i], cgL, C.q0[i] * ctL);
    PR1[i] = -C.mk[i] * coR;
    PR2[i] = -fma(C.mk[i], cgR, C.q1[i] * ctR);
  }
#pragma unroll
  for (int ln = 0; ln < N1; ln++) {
    double ov[N1], lv[N1], rv[N1];
#pragma unroll
    for (int j = 0; j < N1; j++) {
      ov[j] = o[ln * SP + j * SL];
      lv[j] = l[ln * SP + j * SL];
      rv[j] = r[ln * SP + j * SL];
    }
    double dlo = 0.0, dro = 0.0;
#pragma unroll
    for (int j = 0; j < N1; j++) {
      dlo = fma(C.d1[j], lv[j], dlo);
      dro = fma(C.d0[j], rv[j], dro);
    }
#pragma unroll
    for (int i = 0; i < N1; i++) {
      double acc = FIRST ? creact * ov[i] : t[ln * N1 + i];
#pragma unroll
      for (int j = 0; j < N1; j++) acc = fma(T[i * N1 + j], ov[j], acc);
      acc = fma(PL1[i], dlo, acc);
      acc = fma(PL2[i], lv[K], acc);
      acc = fma(PR1[i], dro, acc);
      acc = fma(PR2[i], rv[0], acc);
      t[ln * N1 + i] = acc;
    }
  }
}

// v[line][i] <- sum_j (s M)_ij v[line][j]
template <int K>
__device__ __forceinline__ void kron_mass_lines(const KronConst<K>& C, double s, double (&v)[(K + 1) * (K + 1)]) {
  constexpr int N1 = K + 1;
#pragma unroll
  for (int ln = 0; ln < N1; ln++) {
    double in[N1];
#pragma unroll
    for (int j = 0; j < N1; j++) in[j] = v[ln * N1 + j];
#pragma unroll
    for (int i = 0; i < N1; i++) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < N1; j++) acc = fma(C.M[i * N1 + j] * s, in[j], acc);
      v[ln * N1 + i] = acc;
    }
  }
}
// the same along the other index of the plane: v[j][col] <- sum_j M_ij v[j][col]
template <int K>
__device__ __forceinline__ void kron_mass_cols(const KronConst<K>& C, double (&v)[(K + 1) * (K + 1)]) {
  constexpr int N1 = K + 1;
#pragma unroll
  for (int col = 0; col < N1; col++) {
    double in[N1];
#pragma unroll
    for (int j = 0; j < N1; j++) in[j] = v[j * N1 + col];
#pragma unroll
    for (int i = 0; i < N1; i++) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < N1; j++) acc = fma(C.M[i * N1 + j], in[j], acc);
      v[i * N1 + col] = acc;
    }
  }
}

template <int K>
__global__ void __launch_bounds__(KronDims<K>::THREADS, 1)
    dg_kron_3d_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_yh,
                      const __grid_constant__ CUtensorMap tm_zh, const __grid_constant__ CUtensorMap tm_out,
                      const __grid_constant__ CUtensorMap tm_r0, const DevParams P, const KronConst<K> C,
                      const int accumulate, const int has_r0, double* __restrict__ yout) {
  using D = KronDims<K>;
  constexpr int N1 = D::N1, NLOC = D::NLOC, NPL = D::NPL, TX = D::TX, TY = D::TY, TZ = D::TZ, ROWX = D::ROWX;
  extern __shared__ __align__(128) double tile[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = blockIdx.z * TZ;

  if (tid == 0) {
    k_mbar_init(&bar, 1);
    k_fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    k_mbar_expect_tx(&bar, D::TX_BYTES);
    k_tma_load_4d(tile + D::R0, &tm_rows, 0, x0 / 2 - 1, y0, z0, &bar);
    k_tma_load_4d(tile + D::R1, &tm_yh, 0, x0 / 2, y0 - 1, z0, &bar);
    k_tma_load_4d(tile + D::R2, &tm_yh, 0, x0 / 2, y0 + TY, z0, &bar);
    k_tma_load_4d(tile + D::R3, &tm_zh, 0, x0 / 2, y0, z0 - 1, &bar);
    k_tma_load_4d(tile + D::R4, &tm_zh, 0, x0 / 2, y0, z0 + TZ, &bar);
  }

  // lane -> (cell of the warp, plane index)
  const int lane = tid & 31, warp = tid >> 5;
  const int cw = lane / N1, s = lane - cw * N1;
  const bool lane_on = cw < D::CPW;
  const int ci = warp * D::CPW + (lane_on ? cw : 0);
  const int cx = ci % TX, cy = (ci / TX) % TY, cz = ci / (TX * TY);
  const int gx = x0 + cx, gy = y0 + cy, gz = z0 + cz;
  const int Nx = P.N[0], Ny = P.N[1], Nz = P.N[2];
  const bool active = lane_on && gx < Nx && gy < Ny && gz < Nz;
  double* coef = tile + D::COEF0 + ci * D::COEF;  // [T^x | T^y | T^z | A0[3] | cs[6] | co[6] | creact]
  double* cA0 = coef + 3 * NPL;
  double* ccs = cA0 + 3;
  double* cco = ccs + 6;

  // ---- per-cell coefficients: thread s sets up face s (thread N1-1 also the remaining ones) ------
  bool constrained = false;
  if (active) {
    const int cell = gx + Nx * (gy + Ny * gz);
    const int g[3] = {gx, gy, gz};
    const int N[3] = {Nx, Ny, Nz};
    const int stride[3] = {1, Nx, Nx * Ny};
    for (int f = s; f < 6; f += N1) {
      const int d = f >> 1, side = f & 1;
      const bool onb = side ? g[d] == N[d] - 1 : g[d] == 0;
      const double a = k_load_adiag(P, cell, d);
      const double ao = k_load_adiag(P, onb ? cell : cell + (side ? stride[d] : -stride[d]), d);
      int kind = onb ? 1 : 0;
      if (onb) {
        if (P.side_kind[d][side] == PDB200_SIDE_PROCESSOR) {
          kind = 2;
        } else if (P.bctype) {
          const long long bf = d == 0 ? gy + (long long)Ny * gz : (d == 1 ? gx + (long long)Nx * gz : gx + (long long)Nx * gy);
          kind = P.bctype[P.bf_off[d][side] + bf] == PDB200_BC_DIRICHLET ? 1 : 2;
        }
      }
      const double aih = a * C.ih2[d];
      double csi, coi;
      if (P.weights_on) {
        csi = coi = aih * ao * k_fast_rcp(a + ao + 1e-20);
      } else {
        csi = 0.5 * aih;
        coi = 0.5 * ao * C.ih2[d];
      }
      ccs[f] = kind == 0 ? csi : (kind == 1 ? aih : -0.0);  // -0.0: no u-dependent term (face_has_penalty)
      cco[f] = kind == 0 ? coi : 0.0;
      if (side == 0) cA0[d] = aih;
    }
    if (s == 0) cA0[15] = P.c ? __ldg(P.c + cell) : 0.0;  // creact (last slot of the block)
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
      for (int side = 0; side < 2; side++)
        if ((side ? g[d] == N[d] - 1 : g[d] == 0) && P.side_kind[d][side] == PDB200_SIDE_PROCESSOR) constrained = true;
  }
  __syncwarp();
  // ---- thread s builds row s of T^x, T^y, T^z -------------------------------------------------------
  if (active) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const double A0 = cA0[d], csL = ccs[2 * d], coL = cco[2 * d], csR = ccs[2 * d + 1], coR = cco[2 * d + 1];
      const double cgL = P.weights_on ? C.alpha_pen * (csL + coL) : (face_has_penalty(csL) ? C.alpha_pen * C.ih2[d] : 0.0);
      const double cgR = P.weights_on ? C.alpha_pen * (csR + coR) : (face_has_penalty(csR) ? C.alpha_pen * C.ih2[d] : 0.0);
      const double ctL = -C.theta * csL, ctR = C.theta * csR;
      const double m0c = C.m0[s] * csL, mkc = -C.mk[s] * csR;
#pragma unroll
      for (int j = 0; j < N1; j++) {
        double v = fma(A0, C.MinvK[s * N1 + j], fma(m0c, C.d0[j], mkc * C.d1[j]));
        if (j == 0) v += fma(C.m0[s], cgL, C.q0[s] * ctL);
        if (j == K) v += fma(C.mk[s], cgR, C.q1[s] * ctR);
        coef[d * NPL + s * N1 + j] = v;
      }
    }
  }
  __syncwarp();

  // ---- shared-memory addresses of the cell and its six face neighbours -------------------------
  const int so = D::R0 + ((cz * TY + cy) * ROWX + cx + 2) * NLOC;
  const int xl = so - NLOC, xr = so + NLOC;
  const int yl = cy > 0 ? so - ROWX * NLOC : D::R1 + (cz * TX + cx) * NLOC;
  const int yr = cy < TY - 1 ? so + ROWX * NLOC : D::R2 + (cz * TX + cx) * NLOC;
  const int zl = cz > 0 ? so - TY * ROWX * NLOC : D::R3 + (cy * TX + cx) * NLOC;
  const int zr = cz < TZ - 1 ? so + TY * ROWX * NLOC : D::R4 + (cy * TX + cx) * NLOC;

  k_mbar_wait(&bar, 0);

  double t[NPL], tz[NPL];
  if (active) {
    const double creact = cA0[15];
    // penalty / symmetry coefficients per direction
    double cg[6], ct[6];
#pragma unroll
    for (int f = 0; f < 6; f++) {
      cg[f] = P.weights_on ? C.alpha_pen * (ccs[f] + cco[f]) : (face_has_penalty(ccs[f]) ? C.alpha_pen * C.ih2[f >> 1] : 0.0);
      ct[f] = (f & 1) ? C.theta * ccs[f] : -C.theta * ccs[f];
    }
    // z-plane layout: plane iz = s; index in plane = iy * N1 + ix
    //   x-sweep: lines along x (node stride 1), one line per iy (stride N1):  t[iy][ix]
    kron_sweep<K, 1, N1, true>(tile + so + s * NPL, tile + xl + s * NPL, tile + xr + s * NPL, coef, C, cco[0], cg[0], ct[0],
                               cco[1], cg[1], ct[1], creact, t);
    //   y-sweep: lines along y (node stride N1), one line per ix (stride 1): result indexed [ix][iy]
    double ty[NPL];
    kron_sweep<K, N1, 1, true>(tile + so + s * NPL, tile + yl + s * NPL, tile + yr + s * NPL, coef + NPL, C, cco[2], cg[2],
                               ct[2], cco[3], cg[3], ct[3], 0.0, ty);
#pragma unroll
    for (int iy = 0; iy < N1; iy++)
#pragma unroll
      for (int ix = 0; ix < N1; ix++) t[iy * N1 + ix] += ty[ix * N1 + iy];
    // y-plane layout: plane iy = s; z-sweep: lines along z (node stride NPL), one line per ix: tz[ix][iz]
    kron_sweep<K, NPL, 1, true>(tile + so + s * N1, tile + zl + s * N1, tile + zr + s * N1, coef + 2 * NPL, C, cco[4], cg[4],
                                ct[4], cco[5], cg[5], ct[5], 0.0, tz);
  }
  __syncthreads();  // every thread is done reading the input tile: it becomes staging + scratch
  if (has_r0 && tid == 0) {
    // residual form R(x) = J x + R(0): the tile of the cached R(0) travels into the third region of
    // the dead input tile and is added to y by a second TMA reduce (no thread touches it)
    k_mbar_expect_tx(&bar, D::CELLS * NLOC * 8);
    k_tma_load_4d(tile + 2 * D::CELLS * NLOC, &tm_r0, 0, x0 / 2, y0, z0, &bar);
  }

  double* stage = tile;                         // [CELLS][NLOC], final layout of the output tile
  double* scratch = tile + D::CELLS * NLOC;     // [CELLS][NLOC], accumulators in transit
  double* mys = scratch + ci * NLOC;
  if (active) {
    // tz (y-plane layout, [ix][iz]) -> shared, node-addressed
#pragma unroll
    for (int ix = 0; ix < N1; ix++)
#pragma unroll
      for (int iz = 0; iz < N1; iz++) mys[iz * NPL + s * N1 + ix] = tz[ix * N1 + iz];
  }
  __syncwarp();
  if (active) {
#pragma unroll
    for (int i = 0; i < NPL; i++) t[i] += mys[s * NPL + i];
    kron_mass_lines<K>(C, 1.0, t);  // M along x
    kron_mass_cols<K>(C, t);        // M along y
  }
  __syncwarp();
  if (active) {
#pragma unroll
    for (int i = 0; i < NPL; i++) mys[s * NPL + i] = t[i];
  }
  __syncwarp();
  {
    double* dst = stage + ((cz * TY + cy) * TX + cx) * NLOC;
    if (active) {
      // y-plane layout again: u[ix][iz] = value at node (ix, s, iz); M along z, scaled by |K|
      double u[NPL];
#pragma unroll
      for (int ix = 0; ix < N1; ix++)
#pragma unroll
        for (int iz = 0; iz < N1; iz++) u[ix * N1 + iz] = mys[iz * NPL + s * N1 + ix];
      kron_mass_lines<K>(C, C.vol, u);
#pragma unroll
      for (int ix = 0; ix < N1; ix++)
#pragma unroll
        for (int iz = 0; iz < N1; iz++) dst[iz * NPL + s * N1 + ix] = constrained ? 0.0 : u[ix * N1 + iz];
      if (accumulate && constrained) {  // constrained rows are SET to zero, not incremented
        double* __restrict__ row = yout + (long long)(gx + Nx * (gy + Ny * gz)) * NLOC;
#pragma unroll
        for (int ix = 0; ix < N1; ix++)
#pragma unroll
          for (int iz = 0; iz < N1; iz++) row[iz * NPL + s * N1 + ix] = 0.0;
      }
    } else if (lane_on) {
      // cells of the box outside the grid: clipped by the TMA store, nothing to write
    }
  }
  k_fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    if (accumulate)  // y += tile: read-modify-write in L2, every element touched once per launch
      k_tma_reduce_add_4d(&tm_out, stage, 0, x0 / 2, y0, z0);
    else
      k_tma_store_4d(&tm_out, stage, 0, x0 / 2, y0, z0);
    if (has_r0) {
      k_mbar_wait(&bar, 1);
      k_tma_reduce_add_4d(&tm_out, tile + 2 * D::CELLS * NLOC, 0, x0 / 2, y0, z0);
    }
    k_tma_store_commit_and_wait();
  }
}


typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct KronPlan {
  int k = 0;
  KronConst<4> C4;
  KronConst<3> C3;
  EncodeFn encode = nullptr;
  struct Maps {
    const void* ptr = nullptr;
    CUtensorMap rows, yh, zh, core;
  };
  std::vector<Maps> cache;
  double* scratch = nullptr;
  long long scratch_n = 0;
};

bool dg_kron_supported(const DevParams& P) {
  return P.dg && P.basis == PDB200_BASIS_LAGRANGE && P.dim == 3 && (P.k == 4 || P.k == 3) && P.m >= P.k + 1 && kron_coefficients(P) &&
         P.N[0] % 2 == 0;
}

template <int K>
static void fill_const(KronConst<K>& C, const DevParams& P, const Kron1D& K1) {
  constexpr int N1 = K + 1;
  for (int i = 0; i < N1; i++) {
    for (int j = 0; j < N1; j++) {
      C.MinvK[i * N1 + j] = K1.MinvK[i * MAX_N1 + j];
      C.M[i * N1 + j] = K1.M[i * MAX_N1 + j];
    }
    C.m0[i] = K1.m0[i];
    C.mk[i] = K1.mk[i];
    C.q0[i] = K1.q0[i];
    C.q1[i] = K1.q1[i];
    C.d0[i] = K1.d0[i];
    C.d1[i] = K1.d1[i];
  }
  for (int d = 0; d < 3; d++) C.ih2[d] = 1.0 / (P.h[d] * P.h[d]);
  C.alpha_pen = P.alpha * P.k * (P.k + P.dim - 1);
  C.theta = P.theta;
  C.vol = P.vol;
}

KronPlan* dg_kron_plan_create(const DevParams& P, const Kron1D& K1) {
  KronPlan* plan = new KronPlan;
  plan->k = P.k;
  fill_const<4>(plan->C4, P, K1);
  fill_const<3>(plan->C3, P, K1);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PDB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) throw Error("cuTensorMapEncodeTiled is not available in this driver");
  plan->encode = (EncodeFn)fn;
  PDB_CUDA(cudaFuncSetAttribute(dg_kron_3d_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, KronDims<4>::SMEM_BYTES));
  PDB_CUDA(cudaFuncSetAttribute(dg_kron_3d_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, KronDims<3>::SMEM_BYTES));
  return plan;
}

void dg_kron_plan_destroy(KronPlan* plan) {
  if (!plan) return;
  if (plan->scratch) cudaFree(plan->scratch);
  delete plan;
}

template <int K>
static void kron_encode(KronPlan* plan, CUtensorMap* m, const void* ptr, const DevParams& P, int bx, int by, int bz) {
  constexpr int NLOC = KronDims<K>::NLOC;
  cuuint64_t gdim[4] = {2 * NLOC, (cuuint64_t)P.N[0] / 2, (cuuint64_t)P.N[1], (cuuint64_t)P.N[2]};
  cuuint64_t gstr[3] = {16ull * NLOC, 8ull * NLOC * P.N[0], 8ull * NLOC * P.N[0] * P.N[1]};
  cuuint32_t box[4] = {2 * NLOC, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = plan->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

template <int K>
static KronPlan::Maps& kron_maps(KronPlan* plan, const void* ptr, const DevParams& P) {
  using D = KronDims<K>;
  for (auto& m : plan->cache)
    if (m.ptr == ptr) return m;
  if ((uintptr_t)ptr % 16 != 0) throw Error("Kronecker DG kernel: vectors must be 16-byte aligned");
  if (plan->cache.size() >= 16) plan->cache.erase(plan->cache.begin());
  KronPlan::Maps m;
  m.ptr = ptr;
  kron_encode<K>(plan, &m.rows, ptr, P, D::ROWX / 2, D::TY, D::TZ);
  kron_encode<K>(plan, &m.yh, ptr, P, D::TX / 2, 1, D::TZ);
  kron_encode<K>(plan, &m.zh, ptr, P, D::TX / 2, D::TY, 1);
  kron_encode<K>(plan, &m.core, ptr, P, D::TX / 2, D::TY, D::TZ);
  plan->cache.push_back(m);
  return plan->cache.back();
}

template <int K>
static void kron_launch(KronPlan* plan, const KronConst<K>& C, const DevParams& P, const double* x, double* out,
                        const double* r0, bool accumulate, cudaStream_t s) {
  using D = KronDims<K>;
  const KronPlan::Maps mx = kron_maps<K>(plan, x, P);
  const KronPlan::Maps my = kron_maps<K>(plan, out, P);
  dim3 grid((P.N[0] + D::TX - 1) / D::TX, (P.N[1] + D::TY - 1) / D::TY, (P.N[2] + D::TZ - 1) / D::TZ);
  const KronPlan::Maps mr = r0 ? kron_maps<K>(plan, r0, P) : my;
  dg_kron_3d_kernel<K><<<grid, D::THREADS, D::SMEM_BYTES, s>>>(mx.rows, mx.yh, mx.zh, my.core, mr.core, P, C,
                                                               accumulate ? 1 : 0, r0 ? 1 : 0, out);
  PDB_CUDA(cudaGetLastError());
}

int launch_dg_kron(KronPlan* plan, const DevParams& P, const double* x, double* y, const double* r0, bool overwrite,
                   cudaStream_t s) {
  if (r0 && overwrite) throw Error("the residual form accumulates (r += J x + R(0))");
  // accumulate semantics (y += J x [+ R(0)]) through TMA reduce-add stores
  if (P.k == 4)
    kron_launch<4>(plan, plan->C4, P, x, y, r0, !overwrite, s);
  else
    kron_launch<3>(plan, plan->C3, P, x, y, r0, !overwrite, s);
  int launches = 1;
  return launches;
}

}  // namespace pdb

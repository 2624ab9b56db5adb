// dg_small.cu — Kronecker-factorised jacobian_apply / residual for the SMALL QkDG cells: dim = 2 with
// k = 1, 2 (n = 4, 9 DOFs per cell: the configurations of the reference's own DG tests,
// test/testconvectiondiffusiondg.cc, test/testfastdgassembler.cc, test/matrixfree/matrix_free_linear.cc)
// and dim = 3 with k = 1 (n = 8).
//
// Same operator identity as dg_fast.cu / dg_kron.cu, valid for cell-wise constant
// DIAGONAL diffusion tensors and b = 0:
//     y_e = |K| (M (x) ... (x) M) [ sum_d M^-1 L_d(z_{e-d}, z_e, z_{e+d}) / h_d^2 + c_e z_e ]
// i.e. exactly GridOperator::jacobian_apply for ConvectionDiffusionDG
// (localoperator/convectiondiffusiondg.hh:106-188, 271-471, 684-879); per line of n1 = k+1 nodes
//     t_i += sum_j T_ij o_j + PL1_i (d1.l) + PL2_i l_k + PR1_i (d0.r) + PR2_i r_0
// with the per-cell, per-direction matrix T and vectors PL*, PR* of dg_kron.cu.
//
// Mapping to the machine.  A cell is only 32..72 bytes, so a shared-memory tile with halo would be
// mostly halo.  One thread per cell, cells consecutive along x: a warp reads 32 neighbouring cells
// (1..2.3 KB contiguous) with vector loads; the 2*dim face neighbours are read the same way and hit
// L1/L2 (every cell is fetched from DRAM once).  Everything stays in registers; each thread writes
// its own n results (accumulate forms read-modify-write the same addresses).

#include "common.cuh"

namespace pdb {

namespace {

template <int K>
struct SmallConst {
  static constexpr int N1 = K + 1;
  double MinvK[N1 * N1], M[N1 * N1], m0[N1], mk[N1], q0[N1], q1[N1], d0[N1], d1[N1];
  double ih2[3];
  double alpha_pen, theta, vol;
};

template <int DIM, int K>
struct SL {
  static constexpr int N1 = K + 1, N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
};

__device__ __forceinline__ double s_fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, fma(e, e, e), y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

__device__ __forceinline__ double s_load_adiag(const DevParams& P, long long cell, int d) {
  if (P.a_mode == PDB200_A_IDENTITY) return 1.0;
  if (P.a_mode == PDB200_A_SCALAR) return __ldg(P.A + cell);
  if (P.a_mode == PDB200_A_DIAGONAL) return __ldg(P.A + cell * P.dim + d);
  return __ldg(P.A + cell * P.dim * P.dim + d * (P.dim + 1));
}

template <int N>
__device__ __forceinline__ void load_cell(const double* __restrict__ p, double (&v)[N]) {
  if (N % 2 == 0) {  // n = 4, 8: the cell is 16-byte aligned
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(p) + i);
      v[2 * i] = t.x;
      v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = __ldg(p + i);
  }
}

// t (+)= M^-1 L_d / h_d^2 along direction AXIS for all lines of the cell
template <int DIM, int K, int AXIS, bool FIRST>
__device__ __forceinline__ void small_sweep(const SmallConst<K>& C, const double (&o)[SL<DIM, K>::N],
                                            const double (&l)[SL<DIM, K>::N], const double (&r)[SL<DIM, K>::N], double A0,
                                            double csL, double coL, double cgL, double csR, double coR, double cgR,
                                            double creact, double (&t)[SL<DIM, K>::N]) {
  constexpr int N1 = K + 1, N = SL<DIM, K>::N;
  constexpr int S = AXIS == 0 ? 1 : (AXIS == 1 ? N1 : N1 * N1);
  const double ctL = -C.theta * csL, ctR = C.theta * csR;
  double T[N1 * N1], PL1[N1], PL2[N1], PR1[N1], PR2[N1];
#pragma unroll
  for (int i = 0; i < N1; i++) {
    const double m0c = C.m0[i] * csL, mkc = -C.mk[i] * csR;
    const double eL = fma(C.m0[i], cgL, C.q0[i] * ctL), eR = fma(C.mk[i], cgR, C.q1[i] * ctR);
#pragma unroll
    for (int j = 0; j < N1; j++) {
      double v = fma(A0, C.MinvK[i * N1 + j], fma(m0c, C.d0[j], mkc * C.d1[j]));
      if (j == 0) v += eL;
      if (j == K) v += eR;
      T[i * N1 + j] = v;
    }
    PL1[i] = C.m0[i] * coL;
    PL2[i] = -eL;
    PR1[i] = -C.mk[i] * coR;
    PR2[i] = -eR;
  }
#pragma unroll
  for (int hi = 0; hi < N / (S * N1); hi++)
#pragma unroll
    for (int lo = 0; lo < S; lo++) {
      const int base = hi * S * N1 + lo;
      double dlo = 0.0, dro = 0.0;
#pragma unroll
      for (int j = 0; j < N1; j++) {
        dlo = fma(C.d1[j], l[base + j * S], dlo);
        dro = fma(C.d0[j], r[base + j * S], dro);
      }
#pragma unroll
      for (int i = 0; i < N1; i++) {
        double acc = FIRST ? creact * o[base + i * S] : t[base + i * S];
#pragma unroll
        for (int j = 0; j < N1; j++) acc = fma(T[i * N1 + j], o[base + j * S], acc);
        acc = fma(PL1[i], dlo, acc);
        acc = fma(PL2[i], l[base + K * S], acc);
        acc = fma(PR1[i], dro, acc);
        acc = fma(PR2[i], r[base], acc);
        t[base + i * S] = acc;
      }
    }
}

template <int DIM, int K, int AXIS>
__device__ __forceinline__ void small_mass(const SmallConst<K>& C, double s, double (&t)[SL<DIM, K>::N]) {
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
        for (int j = 0; j < N1; j++) acc = fma(C.M[i * N1 + j] * s, in[j], acc);
        t[base + i * S] = acc;
      }
    }
}

template <int DIM, int K>
__global__ void __launch_bounds__(128) dg_small_kernel(const DevParams P, const SmallConst<K> C, const double* __restrict__ z,
                                                       double* __restrict__ y, const double* __restrict__ r0,
                                                       int accumulate) {
  constexpr int N = SL<DIM, K>::N;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncells) return;
  const int Nx = P.N[0], Ny = P.N[1];
  int g[3];
  g[0] = (int)(cell % Nx);
  g[1] = (int)((cell / Nx) % Ny);
  g[2] = DIM == 3 ? (int)(cell / ((long long)Nx * Ny)) : 0;
  const long long stride[3] = {1, Nx, (long long)Nx * Ny};

  double o[N], t[N];
  load_cell<N>(z + cell * N, o);
  const double creact = P.c ? __ldg(P.c + cell) : 0.0;
  bool constrained = false;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const bool onb[2] = {g[d] == 0, g[d] == P.N[d] - 1};
    const double a = s_load_adiag(P, cell, d);
    double cs[2], co[2], cg[2];
    double nb[2][N];
#pragma unroll
    for (int side = 0; side < 2; side++) {
      int kind = onb[side] ? 1 : 0;
      if (onb[side]) {
        if (P.side_kind[d][side] == PDB200_SIDE_PROCESSOR) {
          kind = 2;
          constrained = true;
        } else if (P.bctype) {
          kind = P.bctype[bface_index(P, g, d, side)] == PDB200_BC_DIRICHLET ? 1 : 2;
        }
      }
      const long long other = onb[side] ? cell : cell + (side ? stride[d] : -stride[d]);
      const double ao = s_load_adiag(P, other, d);
      if (!onb[side]) {
        load_cell<N>(z + other * N, nb[side]);
      } else {
#pragma unroll
        for (int i = 0; i < N; i++) nb[side][i] = 0.0;
      }
      // harmonic weights and penalty, convectiondiffusiondg.hh:326-346 (interior), :717-734 (boundary)
      const double aih = a * C.ih2[d];
      double csi, coi;
      if (P.weights_on) {
        csi = coi = aih * ao * s_fast_rcp(a + ao + 1e-20);
      } else {
        csi = 0.5 * aih;
        coi = 0.5 * ao * C.ih2[d];
      }
      cs[side] = kind == 0 ? csi : (kind == 1 ? aih : 0.0);
      co[side] = kind == 0 ? coi : 0.0;
      cg[side] = P.weights_on ? C.alpha_pen * (cs[side] + co[side]) : (cs[side] != 0.0 ? C.alpha_pen * C.ih2[d] : 0.0);
    }
    const double A0 = a * C.ih2[d];
    if (d == 0)
      small_sweep<DIM, K, 0, true>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], creact, t);
    else if (d == 1)
      small_sweep<DIM, K, 1, false>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], 0.0, t);
    else
      small_sweep<DIM, K, 2, false>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], 0.0, t);
  }
  small_mass<DIM, K, 0>(C, C.vol, t);
  small_mass<DIM, K, 1>(C, 1.0, t);
  if (DIM == 3) small_mass<DIM, K, 2>(C, 1.0, t);
  double* __restrict__ out = y + cell * N;
  if (constrained) {  // constraints/p0.hh:31-41 + constrain_residual: the rows are SET to zero
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = 0.0;
    return;
  }
  if (r0) {
    double rr[N];
    load_cell<N>(r0 + cell * N, rr);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] += rr[i];
  }
  if (accumulate) {
#pragma unroll
    for (int i = 0; i < N; i++) t[i] += out[i];
  }
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<double2*>(out)[i] = make_double2(t[2 * i], t[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = t[i];
  }
}

template <int DIM, int K>
void launch_variant(const DevParams& P, const Kron1D& K1, const double* z, double* y, const double* r0, bool overwrite,
                    cudaStream_t s) {
  SmallConst<K> C;
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
  for (int d = 0; d < 3; d++) C.ih2[d] = d < DIM ? 1.0 / (P.h[d] * P.h[d]) : 0.0;
  C.alpha_pen = P.alpha * P.k * (P.k + P.dim - 1);
  C.theta = P.theta;
  C.vol = P.vol;
  const unsigned blocks = (unsigned)((P.ncells + 127) / 128);
  dg_small_kernel<DIM, K><<<blocks, 128, 0, s>>>(P, C, z, y, r0, overwrite ? 0 : 1);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace

bool dg_small_supported(const DevParams& P) {
  if (!(P.dg && P.b == nullptr && P.a_mode != PDB200_A_FULL && P.m >= P.k + 1)) return false;
  return (P.dim == 2 && (P.k == 1 || P.k == 2)) || (P.dim == 3 && P.k == 1);
}

int launch_dg_small(const DevParams& P, const Kron1D& K1, const double* z, double* y, const double* r0, bool overwrite,
                    cudaStream_t s) {
  if (r0 && overwrite) throw Error("the residual form accumulates (r += J x + R(0))");
  if (P.dim == 2 && P.k == 1) launch_variant<2, 1>(P, K1, z, y, r0, overwrite, s);
  else if (P.dim == 2 && P.k == 2) launch_variant<2, 2>(P, K1, z, y, r0, overwrite, s);
  else if (P.dim == 3 && P.k == 1) launch_variant<3, 1>(P, K1, z, y, r0, overwrite, s);
  else throw Error("dg_small: unsupported (dim, degree)");
  return 1;
}

}  // namespace pdb

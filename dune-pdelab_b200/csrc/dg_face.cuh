// dg_face.cuh — what the thread-per-cell QkDG kernels (dg_small.cu, dg_blockjac.cu) share: the 1-D
// constants of the Kronecker form and the face coefficients of one cell in one direction
// (harmonic weights and penalty, localoperator/convectiondiffusiondg.hh:326-346 interior faces,
// :717-734 Dirichlet boundary faces).
#pragma once

#include "common.cuh"

namespace pdb {
namespace dgface {

template <int K>
struct SmallConst {
  static constexpr int N1 = K + 1;
  double MinvK[N1 * N1], M[N1 * N1], m0[N1], mk[N1], q0[N1], q1[N1], d0[N1], d1[N1];
  double Dn[N1 * N1];  // nodal derivative matrix Dn[i][j] = p_j'(i / k): u'(x_i) = sum_j Dn_ij u_j (convection)
  double ih2[3], ih[3];
  double alpha_pen, theta, vol;
};

// Convection with a cell-wise constant velocity along one direction (convectiondiffusiondg.hh:178-187, 426-448, 797-822,
// 860): cb = b_d / h_d of the cell (volume term, integrated by parts: cb [u'(x_i) + (M^-1 e_0)_i u(0) - (M^-1 e_k)_i u(1)]),
// cu?s / cu?o = beta / h_d on the own / the neighbour's trace at the lower (L) and upper (R) face, whichever is upwind;
// beta comes from the velocity of the larger-index cell.  All zero without a velocity field.
struct Conv1 {
  double cb = 0.0, cuLs = 0.0, cuLo = 0.0, cuRs = 0.0, cuRo = 0.0;
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

// Face coefficients of cell `cell` (coordinates g) in direction d for both sides:
//   cs = w_self a / h^2, co = w_other a_other / h^2, cg = penalty coefficient; A0 = a / h^2.
// kind per side: 0 interior, 1 Dirichlet boundary, 2 no u-dependent term (None / Neumann / Outflow with
// b = 0, processor boundary).  Returns true if the cell touches a processor side (constrained rows).
template <int K>
__device__ __forceinline__ bool direction_coefs(const DevParams& P, const SmallConst<K>& C, long long cell, const int (&g)[3],
                                                int d, const long long (&stride)[3], double& A0, double (&cs)[2],
                                                double (&co)[2], double (&cg)[2], bool (&onb)[2], Conv1* V = nullptr,
                                                int* errflag = nullptr) {
  bool constrained = false;
  onb[0] = g[d] == 0;
  onb[1] = g[d] == P.N[d] - 1;
  const double a = s_load_adiag(P, cell, d);
  const bool conv = V != nullptr && P.b != nullptr;
  const double bself = conv ? __ldg(P.b + cell * P.dim + d) : 0.0;
  if (conv) V->cb = bself * C.ih[d];
#pragma unroll
  for (int side = 0; side < 2; side++) {
    int kind = onb[side] ? 1 : 0;
    bool outflow = false;
    if (onb[side]) {
      if (P.side_kind[d][side] == PDB200_SIDE_PROCESSOR) {
        kind = 2;
        constrained = true;
      } else if (P.bctype) {
        const int bt = P.bctype[bface_index(P, g, d, side)];
        kind = bt == PDB200_BC_DIRICHLET ? 1 : 2;
        outflow = bt == PDB200_BC_OUTFLOW;
      }
    }
    if (conv) {
      // lower face: this cell is the inside (larger-index) cell, n = -e_d; upper interior face: the neighbour is, its
      // outer normal is -e_d and ITS velocity counts (:426-438); boundary faces: own velocity, n = +-e_d (:797)
      double beta, cself = 0.0, cother = 0.0;
      bool self;
      if (side == 0) {
        beta = -bself;
        self = beta >= 0.0;
      } else {
        beta = kind == 0 ? __ldg(P.b + (cell + stride[d]) * P.dim + d) : bself;
        self = kind == 0 ? !(-beta >= 0.0) : beta >= 0.0;
      }
      if (outflow) {
        if (beta < -1e-30 && errflag) *errflag = 1;  // "Outflow boundary condition on inflow!" :802-806
        cself = beta * C.ih[d];
      } else if (kind != 2) {
        if (self) cself = beta * C.ih[d];
        else if (kind == 0) cother = beta * C.ih[d];
      }
      if (side == 0) {
        V->cuLs = cself;
        V->cuLo = cother;
      } else {
        V->cuRs = cself;
        V->cuRo = cother;
      }
    }
    const long long other = onb[side] ? cell : cell + (side ? stride[d] : -stride[d]);
    const double ao = s_load_adiag(P, other, d);
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
    // weightsOff: the penalty does not depend on A (harmonic_average = 1): selected by the kind of face
    cg[side] = P.weights_on ? C.alpha_pen * (cs[side] + co[side]) : (kind != 2 ? C.alpha_pen * C.ih2[d] : 0.0);
  }
  A0 = a * C.ih2[d];
  return constrained;
}

// own-cell 1-D matrix T = M^-1 L_own / h^2 of direction d (row-major N1 x N1), see dg_kron.cu
template <int K>
__device__ __forceinline__ void own_matrix(const SmallConst<K>& C, double A0, double csL, double cgL, double csR,
                                           double cgR, double (&T)[(K + 1) * (K + 1)], double (&eL)[K + 1],
                                           double (&eR)[K + 1], const Conv1& V = Conv1()) {
  constexpr int N1 = K + 1;
  const double ctL = -C.theta * csL, ctR = C.theta * csR;
#pragma unroll
  for (int i = 0; i < N1; i++) {
    const double m0c = C.m0[i] * csL, mkc = -C.mk[i] * csR;
    eL[i] = fma(C.m0[i], cgL, C.q0[i] * ctL);
    eR[i] = fma(C.mk[i], cgR, C.q1[i] * ctR);
#pragma unroll
    for (int j = 0; j < N1; j++) {
      double v = fma(A0, C.MinvK[i * N1 + j], fma(m0c, C.d0[j], mkc * C.d1[j]));
      v = fma(V.cb, C.Dn[i * N1 + j], v);                       // convection, volume part: cb u'(x_i)
      if (j == 0) v += fma(C.m0[i], V.cuLs + V.cb, eL[i]);      // own trace at the lower face
      if (j == K) v += fma(C.mk[i], V.cuRs - V.cb, eR[i]);      // own trace at the upper face
      T[i * N1 + j] = v;
    }
  }
}

template <int K>
inline void fill_small_const(SmallConst<K>& C, const DevParams& P, const Kron1D& K1) {
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
    for (int j = 0; j < N1; j++) C.Dn[i * N1 + j] = K1.Dn[i * MAX_N1 + j];
  }
  for (int d = 0; d < 3; d++) {
    C.ih2[d] = d < P.dim ? 1.0 / (P.h[d] * P.h[d]) : 0.0;
    C.ih[d] = d < P.dim ? 1.0 / P.h[d] : 0.0;
  }
  C.alpha_pen = P.alpha * P.k * (P.k + P.dim - 1);
  C.theta = P.theta;
  C.vol = P.vol;
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

template <int N>
__device__ __forceinline__ void store_cell(double* __restrict__ p, const double (&v)[N]) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<double2*>(p)[i] = make_double2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = v[i];
  }
}

// ---- per-cell sweeps of the thread-per-cell kernels (dg_small.cu, block SOR in dg_blockjac.cu) -------------

// t (+)= M^-1 L_d / h_d^2 along direction AXIS for all lines of the cell
template <int DIM, int K, int AXIS, bool FIRST>
__device__ __forceinline__ void small_sweep(const SmallConst<K>& C, const double (&o)[SL<DIM, K>::N],
                                            const double (&l)[SL<DIM, K>::N], const double (&r)[SL<DIM, K>::N], double A0,
                                            double csL, double coL, double cgL, double csR, double coR, double cgR,
                                            double creact, double (&t)[SL<DIM, K>::N], const Conv1& V = Conv1()) {
  constexpr int N1 = K + 1, N = SL<DIM, K>::N;
  constexpr int S = AXIS == 0 ? 1 : (AXIS == 1 ? N1 : N1 * N1);
  double T[N1 * N1], eL[N1], eR[N1], PL1[N1], PL2[N1], PR1[N1], PR2[N1];
  own_matrix<K>(C, A0, csL, cgL, csR, cgR, T, eL, eR, V);
#pragma unroll
  for (int i = 0; i < N1; i++) {
    PL1[i] = C.m0[i] * coL;
    PL2[i] = fma(C.m0[i], V.cuLo, -eL[i]);  // the neighbour's trace: jump terms and, if it is upwind, the convective flux
    PR1[i] = -C.mk[i] * coR;
    PR2[i] = fma(C.mk[i], V.cuRo, -eR[i]);
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

}  // namespace dgface
}  // namespace pdb

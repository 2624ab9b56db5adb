// dg_generic.cu — reference-order QkDG kernels (any dim in {2,3}, k in 1..4, any coefficient mode).
//
// One thread per cell, gather formulation: a cell evaluates its volume integral and ITS side of
// all 2*dim faces, so every row of the result is written exactly once, without atomics and in a
// fixed order.  Per quadrature point the loops are the dense loops of the reference
//   ConvectionDiffusionDG::alpha_volume               localoperator/convectiondiffusiondg.hh:106-188
//   ConvectionDiffusionDG::alpha_skeleton             :271-471   (s- or n-side, see below)
//   ConvectionDiffusionDG::residual_boundary_integral :684-879
//   ConvectionDiffusionDG::lambda_volume              :1048-1075
// with basis values formed on the fly from the 1-D tables instead of LocalBasisCache look-ups.
// This is the parity path for configurations without a specialised kernel; dg_fast.cu holds the
// bandwidth-oriented kernel of the headline configuration.
//
// Face ownership.  The reference integrates an interior face once, from the cell with the larger
// index ("inside" s, default/assembler.hh:178-184) and scatters to both cells.  Written from the
// point of view of either cell ("self", outward normal n, neighbour "other") both updates read
//   r_self += [ up(u) (b_F.n) - (w_self An_self.grad u_self + w_other An_other.grad u_other)
//               + gamma (u_self - u_other) ] psi_self f
//           + theta (u_self - u_other) w_self (An_self . grad psi_self) f
// with b_F the velocity of the larger-index cell (:426) — derivation in DESIGN.md §4.

#include "common.cuh"

namespace pdb {
namespace {

template <int DIM, int K>
struct Loc {
  static constexpr int N1 = K + 1;
  static constexpr int N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
};

// 1-D basis values and physical derivatives at the point with table indices pt[]
template <int DIM, int K>
__device__ __forceinline__ void point_tables(const DevParams& P, const int pt[3], double pv[3][K + 1],
                                             double dv[3][K + 1]) {
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int i = 0; i <= K; i++) {
      if (d < DIM) {
        pv[d][i] = P.P[pt[d] * (K + 1) + i];
        dv[d][i] = P.ih[d] * P.DP[pt[d] * (K + 1) + i];
      } else {
        pv[d][i] = i == 0 ? 1.0 : 0.0;
        dv[d][i] = 0.0;
      }
    }
}

// u = sum x_i phi_i, grad u = sum x_i grad phi_i   (convectiondiffusiondg.hh:153-172)
template <int DIM, int K>
__device__ __forceinline__ void interp(const double* x, const double pv[3][K + 1], const double dv[3][K + 1],
                                       double& u, double gu[3]) {
  constexpr int N1 = K + 1, N2 = DIM == 3 ? N1 : 1;
  u = 0.0;
  gu[0] = gu[1] = gu[2] = 0.0;
  int idx = 0;
  for (int i2 = 0; i2 < N2; i2++)
    for (int i1 = 0; i1 < N1; i1++) {
      double a = pv[1][i1] * pv[2][i2];
      double b1 = dv[1][i1] * pv[2][i2];
      double b2 = pv[1][i1] * dv[2][i2];
      for (int i0 = 0; i0 < N1; i0++, idx++) {
        double xi = x[idx];
        double p0 = pv[0][i0];
        u += xi * (p0 * a);
        gu[0] += xi * (dv[0][i0] * a);
        gu[1] += xi * (p0 * b1);
        if (DIM == 3) gu[2] += xi * (p0 * b2);
      }
    }
}

// r_i += cphi * phi_i + cg . grad phi_i
template <int DIM, int K>
__device__ __forceinline__ void accum(double* r, const double pv[3][K + 1], const double dv[3][K + 1],
                                      double cphi, const double cg[3]) {
  constexpr int N1 = K + 1, N2 = DIM == 3 ? N1 : 1;
  int idx = 0;
  for (int i2 = 0; i2 < N2; i2++)
    for (int i1 = 0; i1 < N1; i1++) {
      double a = pv[1][i1] * pv[2][i2];
      double b1 = dv[1][i1] * pv[2][i2];
      double b2 = pv[1][i1] * dv[2][i2];
      for (int i0 = 0; i0 < N1; i0++, idx++) {
        double p0 = pv[0][i0];
        double v = cphi * (p0 * a) + cg[0] * (dv[0][i0] * a) + cg[1] * (p0 * b1);
        if (DIM == 3) v += cg[2] * (p0 * b2);
        r[idx] += v;
      }
    }
}

template <int DIM>
__device__ __forceinline__ double dotd(const double* a, const double* b) {
  double s = a[0] * b[0] + a[1] * b[1];
  if (DIM == 3) s += a[2] * b[2];
  return s;
}

template <int DIM, int K, bool RESIDUAL>
__global__ void __launch_bounds__(128) dg_generic_kernel(const DevParams P, const double* __restrict__ x,
                                                         double* __restrict__ y, int overwrite,
                                                         int* __restrict__ errflag) {
  constexpr int N = Loc<DIM, K>::N;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncells) return;
  int c[3];
  {
    long long e = cell;
    c[0] = (int)(e % P.N[0]);
    e /= P.N[0];
    c[1] = (int)(e % P.N[1]);
    c[2] = (int)(e / P.N[1]);
  }
  double xs[N], xo[N], r[N];
  for (int i = 0; i < N; i++) {
    xs[i] = x[cell * N + i];
    r[i] = 0.0;
  }
  double A_s[3][3], b_s[3];
  load_A_cell(P, cell, A_s);
  const bool pwA = pw_A(P);  // permeabilityIsConstantPerCell() == false: A re-evaluated at every point
  const int m = P.m;

  double pv[3][K + 1], dv[3][K + 1];
  // ---- volume: lambda_volume (:1048-1075) + alpha_volume (:106-188) -------------------------
  for (int q = 0; q < P.nq; q++) {
    int pt[3] = {0, 0, 0};
    double weight = 1.0;
    {
      int qq = q;
      for (int d = 0; d < DIM; d++) {
        pt[d] = qq % m;
        qq /= m;
        weight *= P.wq[pt[d]];
      }
    }
    point_tables<DIM, K>(P, pt, pv, dv);
    if (pwA) load_A_at(P, cell, q, A_s);  // :143-146
    load_b(P, cell, q, b_s);              // param.b(cell, ip.position()), :178
    const double c_s = load_c(P, cell, q);  // :181
    double u, gu[3];
    interp<DIM, K>(xs, pv, dv, u, gu);
    double Agu[3];
    for (int i = 0; i < 3; i++) Agu[i] = A_s[i][0] * gu[0] + A_s[i][1] * gu[1] + A_s[i][2] * gu[2];
    const double factor = weight * P.vol;
    double cphi = c_s * u * factor;
    if (RESIDUAL && P.f) cphi -= P.f[cell * P.nq + q] * factor;
    double cg[3];
    for (int d = 0; d < 3; d++) cg[d] = (Agu[d] - u * b_s[d]) * factor;
    accum<DIM, K>(r, pv, dv, cphi, cg);
  }

  // ---- faces in YaspGrid intersection order 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z --------------------
  bool constrained = false;
  const long long stride[3] = {1, (long long)P.N[0], (long long)P.N[0] * P.N[1]};
  const int degree = K;
  for (int dir = 0; dir < DIM; dir++)
    for (int side = 0; side < 2; side++) {
      const bool onb = side ? c[dir] == P.N[dir] - 1 : c[dir] == 0;
      const double nsign = side ? 1.0 : -1.0;
      double An_s[3];
      if (pwA) load_A_cell(P, cell, A_s);  // the volume loop left the last point's tensor behind
      for (int d = 0; d < 3; d++) An_s[d] = A_s[d][dir] * nsign;
      const double area = P.area[dir];
      if (!onb) {
        // interior face, this cell's side of alpha_skeleton
        const long long other = cell + (side ? stride[dir] : -stride[dir]);
        for (int i = 0; i < N; i++) xo[i] = x[other * N + i];
        double A_o[3][3], An_o[3], b_F[3];
        load_A_cell(P, other, A_o);
        for (int d = 0; d < 3; d++) An_o[d] = A_o[d][dir] * nsign;
        const long long bc = side ? other : cell;  // velocity of the larger-index cell (:426) ...
        const int bside = 0;                       // ... evaluated on ITS lower face (geo_in_inside)
        const double h_F = fmin(P.vol, P.vol) / area;  // :313
        double omega_s, omega_o, harmonic_average;
        if (P.weights_on) {  // :326-338
          const double delta_s = An_s[dir] * nsign, delta_o = An_o[dir] * nsign;
          omega_s = delta_o / (delta_s + delta_o + 1e-20);
          omega_o = delta_s / (delta_s + delta_o + 1e-20);
          harmonic_average = 2.0 * delta_s * delta_o / (delta_s + delta_o + 1e-20);
        } else {
          omega_s = omega_o = 0.5;
          harmonic_average = 1.0;
        }
        double penalty = (P.alpha / h_F) * harmonic_average * degree * (degree + DIM - 1);  // :346
        double po[3][K + 1], dvo[3][K + 1];
        for (int q = 0; q < P.nfq; q++) {
          if (pwA) {  // :367-382: both tensors at the face point, weights and penalty only with weightsOn
            load_A_at(P, cell, face_pt(P, dir, side, q), A_s);
            load_A_at(P, other, face_pt(P, dir, 1 - side, q), A_o);
            for (int d = 0; d < 3; d++) {
              An_s[d] = A_s[d][dir] * nsign;
              An_o[d] = A_o[d][dir] * nsign;
            }
            if (P.weights_on) {
              const double delta_s = An_s[dir] * nsign, delta_o = An_o[dir] * nsign;
              omega_s = delta_o / (delta_s + delta_o + 1e-20);
              omega_o = delta_s / (delta_s + delta_o + 1e-20);
              harmonic_average = 2.0 * delta_s * delta_o / (delta_s + delta_o + 1e-20);
              penalty = (P.alpha / h_F) * harmonic_average * degree * (degree + DIM - 1);
            }
          }
          load_b(P, bc, face_pt(P, dir, bside, q), b_F);  // param.b(cell_inside, iplocal_s), :426
          const double betan = b_F[dir] * nsign;
          // upwinding (:429-438): the reference tests (b.n_ref >= 0) with n_ref the normal of the
          // larger-index cell
          const bool take_self = side == 0 ? (betan >= 0.0) : !((-betan) >= 0.0);
          int pt_s[3] = {0, 0, 0}, pt_o[3];
          double weight = 1.0;
          {
            int qq = q;
            for (int d = 0; d < DIM; d++)
              if (d != dir) {
                pt_s[d] = qq % m;
                qq /= m;
                weight *= P.wq[pt_s[d]];
              }
          }
          for (int d = 0; d < 3; d++) pt_o[d] = pt_s[d];
          pt_s[dir] = side ? m + 1 : m;
          pt_o[dir] = side ? m : m + 1;
          point_tables<DIM, K>(P, pt_s, pv, dv);
          point_tables<DIM, K>(P, pt_o, po, dvo);
          double u_s, gu_s[3], u_o, gu_o[3];
          interp<DIM, K>(xs, pv, dv, u_s, gu_s);
          interp<DIM, K>(xo, po, dvo, u_o, gu_o);
          const double factor = weight * area;
          double val = (take_self ? u_s : u_o) * betan * factor;                                   // :444
          val += -(omega_s * dotd<DIM>(An_s, gu_s) + omega_o * dotd<DIM>(An_o, gu_o)) * factor;  // :451
          val += penalty * (u_s - u_o) * factor;                                                    // :465
          const double t3 = (u_s - u_o) * factor * P.theta * omega_s;                               // :458
          double cg[3] = {t3 * An_s[0], t3 * An_s[1], t3 * An_s[2]};
          accum<DIM, K>(r, pv, dv, val, cg);
        }
      } else if (P.side_kind[dir][side] == PDB200_SIDE_PROCESSOR) {
        constrained = true;  // constraints/p0.hh:31-41; nothing is integrated (assembler.hh:239-250)
      } else {
        // residual_boundary_integral, :684-879
        const long long bf = bface_index(P, c, dir, side);
        const double h_F = P.vol / area;                                                    // :717
        double harmonic_average = P.weights_on ? An_s[dir] * nsign : 1.0;                   // :724-727
        double penalty = (P.alpha / h_F) * harmonic_average * degree * (degree + DIM - 1);  // :734
        for (int q = 0; q < P.nfq; q++) {
          if (pwA) {  // :752-760
            load_A_at(P, cell, face_pt(P, dir, side, q), A_s);
            for (int d = 0; d < 3; d++) An_s[d] = A_s[d][dir] * nsign;
            if (P.weights_on) {
              harmonic_average = An_s[dir] * nsign;
              penalty = (P.alpha / h_F) * harmonic_average * degree * (degree + DIM - 1);
            }
          }
          const int bctype = load_bctype(P, bf, q);  // param.bctype(ig.intersection(), ip.position()), :763
          if (bctype == PDB200_BC_NONE) continue;
          load_b(P, cell, face_pt(P, dir, side, q), b_s);  // param.b(cell_inside, iplocal_s), :797
          const double betan = b_s[dir] * nsign;
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
          point_tables<DIM, K>(P, pt, pv, dv);
          const double factor = weight * area;
          const double zero3[3] = {0, 0, 0};
          if (bctype == PDB200_BC_NEUMANN) {
            if (RESIDUAL && P.j) accum<DIM, K>(r, pv, dv, P.j[bf * P.nfq + q] * factor, zero3);  // :778-789
            continue;
          }
          double u_s, gu_s[3];
          interp<DIM, K>(xs, pv, dv, u_s, gu_s);
          if (bctype == PDB200_BC_OUTFLOW) {  // :800-822
            if (betan < -1e-30) {
              *errflag = 1;
              continue;
            }
            double val = u_s * betan * factor;
            if (RESIDUAL && P.o) val += P.o[bf * P.nfq + q] * factor;
            accum<DIM, K>(r, pv, dv, val, zero3);
            continue;
          }
          double g = (RESIDUAL && P.g) ? P.g[bf * P.nfq + q] : 0.0;  // :840-844
          double val = ((betan >= 0.0) ? u_s : g) * betan * factor;   // :860
          val += -(dotd<DIM>(An_s, gu_s)) * factor;                   // :865-867
          val += penalty * (u_s - g) * factor;                        // :875
          const double t3 = (u_s - g) * factor * P.theta;             // :870-872
          double cg[3] = {t3 * An_s[0], t3 * An_s[1], t3 * An_s[2]};
          accum<DIM, K>(r, pv, dv, val, cg);
        }
      }
    }

  // onUnbindLFSV: y += r (jacobianapplyengine.hh:197-202); postAssembly: constrained rows := 0
  for (int i = 0; i < N; i++) {
    double v = overwrite ? r[i] : y[cell * N + i] + r[i];
    y[cell * N + i] = constrained ? 0.0 : v;
  }
}

template <int DIM, int K>
void launch_dk(const DevParams& P, const double* x, double* y, bool residual, bool overwrite, int* err,
               cudaStream_t s) {
  const int threads = 128;
  const unsigned blocks = (unsigned)((P.ncells + threads - 1) / threads);
  if (residual)
    dg_generic_kernel<DIM, K, true><<<blocks, threads, 0, s>>>(P, x, y, overwrite ? 1 : 0, err);
  else
    dg_generic_kernel<DIM, K, false><<<blocks, threads, 0, s>>>(P, x, y, overwrite ? 1 : 0, err);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace

void launch_dg_generic(const DevParams& P, const double* x, double* y, bool residual, bool overwrite,
                       int* errflag, cudaStream_t s) {
  if (P.m != P.k + 1) throw Error("generic DG kernel: intorderadd must be 0 or 1 (m = k+1 Gauss points)");
#define PDB_CASE(D, KK)                                              \
  if (P.dim == D && P.k == KK) {                                     \
    launch_dk<D, KK>(P, x, y, residual, overwrite, errflag, s);      \
    return;                                                          \
  }
  PDB_CASE(2, 1) PDB_CASE(2, 2) PDB_CASE(2, 3) PDB_CASE(2, 4)
  PDB_CASE(3, 1) PDB_CASE(3, 2) PDB_CASE(3, 3) PDB_CASE(3, 4)
#undef PDB_CASE
  throw Error("generic DG kernel: unsupported (dim, degree); compiled: dim 2..3, degree 1..4");
}

}  // namespace pdb

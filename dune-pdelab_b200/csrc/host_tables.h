// host_tables.h — host-side set-up of the operator handle: Gauss-Legendre rule, Lagrange tables,
// exactly integrated 1-D matrices, and the closed-form DOF numbering of structured YaspGrid spaces.
//
// Reference behaviour restated here (paths relative to /root/reference/dune/pdelab/):
//   * basis: finiteelement/qkdglagrange.hh:55-79 (p, dp), :118-128 (multi-index, x fastest)
//   * ordering: ordering/leafgridviewordering.hh:166-184 (geometry-type blocks in
//     GlobalGeometryTypeIndex order), ordering/leaforderingbase.hh:97-203 (index = offset +
//     entity index * dofs per entity), gridfunctionspace/localfunctionspace.hh:616-654
//   * constraints: constraints/conforming.hh:53-93,108-138, constraints/p0.hh:31-41
#pragma once

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace pdb {

inline double host_lagrange_p(int k, int i, double x) {
  double r = 1.0;
  for (int j = 0; j <= k; j++)
    if (j != i) r *= (k * x - j) / (i - j);
  return r;
}
inline double host_lagrange_dp(int k, int i, double x) {
  double r = 0.0;
  for (int j = 0; j <= k; j++)
    if (j != i) {
      double prod = (k * 1.0) / (i - j);
      for (int l = 0; l <= k; l++)
        if (l != i && l != j) prod *= (k * x - l) / (i - l);
      r += prod;
    }
  return r;
}
inline long double host_lagrange_p_ld(int k, int i, long double x) {
  long double r = 1.0L;
  for (int j = 0; j <= k; j++)
    if (j != i) r *= (k * x - j) / (long double)(i - j);
  return r;
}
inline long double host_lagrange_dp_ld(int k, int i, long double x) {
  long double r = 0.0L;
  for (int j = 0; j <= k; j++)
    if (j != i) {
      long double prod = (long double)k / (i - j);
      for (int l = 0; l <= k; l++)
        if (l != i && l != j) prod *= (k * x - l) / (long double)(i - l);
      r += prod;
    }
  return r;
}

// shifted Legendre polynomials P_n(2x - 1) and derivatives (finiteelement/qkdglegendre.hh:76-139)
inline void host_legendre_ld(int k, long double x, long double* v, long double* dv) {
  v[0] = 1;
  dv[0] = 0;
  if (k >= 1) {
    v[1] = 2 * x - 1;
    dv[1] = 2;
  }
  for (int n = 2; n <= k; n++) {
    v[n] = ((2 * n - 1) * (2 * x - 1) * v[n - 1] - (n - 1) * v[n - 2]) / n;
    dv[n] = (2 * x - 1) * dv[n - 1] + 2 * n * v[n - 1];
  }
}
// Gauss-Lobatto points on [0,1], ascending (finiteelement/qkdglobatto.hh:28-66 sorts dune-geometry's GaussLobatto rule
// so that the lower half lies below 1/2; closed forms of the roots of P_k' for k <= 4)
inline void host_lobatto_points(int k, long double* xi) {
  long double t[5] = {0, 0, 0, 0, 0};
  switch (k) {
    case 1: t[0] = -1, t[1] = 1; break;
    case 2: t[0] = -1, t[1] = 0, t[2] = 1; break;
    case 3: t[0] = -1, t[1] = -sqrtl(0.2L), t[2] = sqrtl(0.2L), t[3] = 1; break;
    default: t[0] = -1, t[1] = -sqrtl(3.0L / 7.0L), t[2] = 0, t[3] = sqrtl(3.0L / 7.0L), t[4] = 1; break;
  }
  for (int i = 0; i <= k; i++) xi[i] = (1 + t[i]) / 2;
}
inline long double host_nodal_p_ld(int k, const long double* xi, int i, long double x) {
  long double r = 1;
  for (int j = 0; j <= k; j++)
    if (j != i) r *= (x - xi[j]) / (xi[i] - xi[j]);
  return r;
}
inline long double host_nodal_dp_ld(int k, const long double* xi, int i, long double x) {
  long double r = 0;
  for (int j = 0; j <= k; j++)
    if (j != i) {
      long double prod = 1 / (xi[i] - xi[j]);
      for (int l = 0; l <= k; l++)
        if (l != i && l != j) prod *= (x - xi[l]) / (xi[i] - xi[l]);
      r += prod;
    }
  return r;
}
// value / derivative of 1-D basis function i of the QkDG space (basis != Lagrange)
inline void host_basis_ld(int basis, int k, int i, long double x, long double* p, long double* dp) {
  if (basis == PDB200_BASIS_LEGENDRE) {
    long double v[MAX_N1], dv[MAX_N1];
    host_legendre_ld(k, x, v, dv);
    *p = v[i];
    *dp = dv[i];
  } else {
    long double xi[MAX_N1];
    host_lobatto_points(k, xi);
    *p = host_nodal_p_ld(k, xi, i, x);
    *dp = host_nodal_dp_ld(k, xi, i, x);
  }
}

// m-point Gauss-Legendre rule on [0,1], ascending (dune-geometry QuadratureRules, GaussLegendre)
inline void host_gauss(int m, std::vector<long double>& x, std::vector<long double>& w) {
  x.assign(m, 0);
  w.assign(m, 0);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < m; i++) {
    long double t = cosl(pi * (i + 0.75L) / (m + 0.5L)), dp = 1;
    for (int it = 0; it < 100; it++) {
      long double p0 = 1, p1 = t;
      for (int j = 2; j <= m; j++) {
        long double p2 = ((2 * j - 1) * t * p1 - (j - 1) * p0) / j;
        p0 = p1;
        p1 = p2;
      }
      dp = m * (t * p1 - p0) / (t * t - 1);
      long double dt = p1 / dp;
      t -= dt;
      if (fabsl(dt) < 1e-19L) break;
    }
    long double p0 = 1, p1 = t;
    for (int j = 2; j <= m; j++) {
      long double p2 = ((2 * j - 1) * t * p1 - (j - 1) * p0) / j;
      p0 = p1;
      p1 = p2;
    }
    dp = m * (t * p1 - p0) / (t * t - 1);
    x[m - 1 - i] = (1 + t) / 2;
    w[m - 1 - i] = 1 / ((1 - t * t) * dp * dp);
  }
}

inline void host_fill_tables(DevParams& P, Kron1D& K, std::vector<double>& xq, std::vector<double>& wq) {
  const int k = P.k, n1 = P.n1, m = P.m;
  std::vector<long double> gx, gw;
  host_gauss(m, gx, gw);
  xq.resize(m);
  wq.resize(m);
  for (int q = 0; q < m; q++) {
    xq[q] = (double)gx[q];
    wq[q] = (double)gw[q];
    P.wq[q] = wq[q];
  }
  for (int pt = 0; pt < m + 2; pt++) {
    const double x = pt < m ? xq[pt] : (pt == m ? 0.0 : 1.0);
    for (int i = 0; i < n1; i++) {
      if (P.basis == PDB200_BASIS_LAGRANGE) {
        P.P[pt * n1 + i] = host_lagrange_p(k, i, x);
        P.DP[pt * n1 + i] = host_lagrange_dp(k, i, x);
      } else {  // the points are long double where they exist (Gauss points), exact 0 / 1 otherwise
        long double p, dp;
        host_basis_ld(P.basis, k, i, pt < m ? gx[pt] : (long double)x, &p, &dp);
        P.P[pt * n1 + i] = (double)p;
        P.DP[pt * n1 + i] = (double)dp;
      }
    }
  }
  // exact 1-D matrices in long double (Gauss rule with k+1 points is exact for degree 2k+1)
  std::vector<long double> ex, ew;
  host_gauss(k + 2, ex, ew);
  long double M[MAX_N1][MAX_N1] = {}, S[MAX_N1][MAX_N1] = {}, e0[MAX_N1] = {}, e1[MAX_N1] = {};
  for (int q = 0; q < k + 2; q++)
    for (int i = 0; i < n1; i++) {
      for (int j = 0; j < n1; j++) {
        M[i][j] += ew[q] * host_lagrange_p_ld(k, i, ex[q]) * host_lagrange_p_ld(k, j, ex[q]);
        S[i][j] += ew[q] * host_lagrange_dp_ld(k, i, ex[q]) * host_lagrange_dp_ld(k, j, ex[q]);
      }
      e0[i] += ew[q] * host_lagrange_dp_ld(k, i, ex[q]) * (1 - ex[q]);
      e1[i] += ew[q] * host_lagrange_dp_ld(k, i, ex[q]) * ex[q];
    }
  // inverse of M by Gauss-Jordan
  long double aug[MAX_N1][2 * MAX_N1] = {};
  for (int i = 0; i < n1; i++) {
    for (int j = 0; j < n1; j++) aug[i][j] = M[i][j];
    aug[i][n1 + i] = 1;
  }
  for (int c = 0; c < n1; c++) {
    int piv = c;
    for (int r = c + 1; r < n1; r++)
      if (fabsl(aug[r][c]) > fabsl(aug[piv][c])) piv = r;
    for (int j = 0; j < 2 * n1; j++) std::swap(aug[c][j], aug[piv][j]);
    long double d = aug[c][c];
    for (int j = 0; j < 2 * n1; j++) aug[c][j] /= d;
    for (int r = 0; r < n1; r++)
      if (r != c) {
        long double f = aug[r][c];
        for (int j = 0; j < 2 * n1; j++) aug[r][j] -= f * aug[c][j];
      }
  }
  long double Mi[MAX_N1][MAX_N1];
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n1; j++) Mi[i][j] = aug[i][n1 + j];
  long double d0[MAX_N1], d1[MAX_N1];
  for (int i = 0; i < n1; i++) {
    d0[i] = host_lagrange_dp_ld(k, i, 0.0L);
    d1[i] = host_lagrange_dp_ld(k, i, 1.0L);
  }
  for (int i = 0; i < n1; i++) {
    long double E0 = 0, E1 = 0, q0 = 0, q1 = 0;
    for (int j = 0; j < n1; j++) {
      K.M[i * MAX_N1 + j] = (double)M[i][j];
      K.Minv[i * MAX_N1 + j] = (double)Mi[i][j];
      long double mk = 0;
      for (int l = 0; l < n1; l++) mk += Mi[i][l] * S[l][j];
      K.MinvK[i * MAX_N1 + j] = (double)mk;
      E0 += Mi[i][j] * e0[j];
      E1 += Mi[i][j] * e1[j];
      q0 += Mi[i][j] * d0[j];
      q1 += Mi[i][j] * d1[j];
    }
    K.d0[i] = (double)d0[i];
    K.d1[i] = (double)d1[i];
    for (int j = 0; j < n1; j++) K.Dn[i * MAX_N1 + j] = (double)host_lagrange_dp_ld(k, j, (long double)i / k);
    K.E0[i] = (double)E0;  // valid for k = 2 (u' linear): K o = e0 u'(0) + e1 u'(1)
    K.E1[i] = (double)E1;
    K.m0[i] = (double)Mi[i][0];
    K.mk[i] = (double)Mi[i][k];
    K.q0[i] = (double)q0;
    K.q1[i] = (double)q1;
  }
}

// ---- DOF numbering ----------------------------------------------------------------------------

// Closed-form container indices of a conforming Q2 space: one DOF per sub-entity, blocks in
// entity-dimension order (vertices | edges | faces | cells); inside a block the entities are
// grouped by extension bitset s (bit d set = entity extends along d) in increasing integer value,
// lexicographic with x fastest inside a group (YaspGrid index set).
struct QkLayout {
  int dim, k;
  int N[3];
  long long block_off[4];  // by entity dimension
  long long group_off[8];  // by bitset, relative to its block
  long long ndofs;
};

__host__ __device__ inline long long qk_group_size(const QkLayout& L, int s) {
  long long sz = 1;
  for (int d = 0; d < L.dim; d++) sz *= ((s >> d) & 1) ? L.N[d] : L.N[d] + 1;
  return sz;
}

inline QkLayout make_qk_layout(const DevParams& P) {
  QkLayout L;
  L.dim = P.dim;
  L.k = P.k;
  for (int d = 0; d < 3; d++) L.N[d] = P.N[d];
  for (int i = 0; i < 4; i++) L.block_off[i] = 0;
  for (int i = 0; i < 8; i++) L.group_off[i] = 0;
  if (P.k == 1) {
    L.ndofs = 1;
    for (int d = 0; d < P.dim; d++) L.ndofs *= P.N[d] + 1;
    return L;
  }
  long long count[4] = {0, 0, 0, 0};
  for (int edim = 0; edim <= P.dim; edim++)
    for (int s = 0; s < (1 << P.dim); s++) {
      int pc = 0;
      for (int d = 0; d < P.dim; d++) pc += (s >> d) & 1;
      if (pc == edim) {
        L.group_off[s] = count[edim];
        count[edim] += qk_group_size(L, s);
      }
    }
  long long off = 0;
  for (int edim = 0; edim <= P.dim; edim++) {
    L.block_off[edim] = off;
    off += count[edim];
  }
  L.ndofs = off;
  return L;
}

// container index of the DOF at lattice point l (0 <= l_d <= k*N_d)
__host__ __device__ inline long long qk_lattice_index(const QkLayout& L, const int l[3]) {
  if (L.k == 1) {
    long long idx = 0, stride = 1;
    for (int d = 0; d < L.dim; d++) {
      idx += stride * l[d];
      stride *= L.N[d] + 1;
    }
    return idx;
  }
  int s = 0, edim = 0;
  long long idx = 0, stride = 1;
  for (int d = 0; d < L.dim; d++) {
    const int ext = l[d] & 1;
    s |= ext << d;
    edim += ext;
    idx += stride * (l[d] >> 1);
    stride *= ext ? L.N[d] : L.N[d] + 1;
  }
  return L.block_off[edim] + L.group_off[s] + idx;
}

inline long long host_num_dofs(const DevParams& P) {
  if (P.dg) return P.ncells * P.n;
  return make_qk_layout(P).ndofs;
}

inline void host_cell_dof_indices(const DevParams& P, long long cell, uint64_t* idx) {
  if (P.dg) {
    for (int i = 0; i < P.n; i++) idx[i] = (uint64_t)(cell * P.n + i);
    return;
  }
  QkLayout L = make_qk_layout(P);
  int c[3];
  c[0] = (int)(cell % P.N[0]);
  c[1] = (int)((cell / P.N[0]) % P.N[1]);
  c[2] = (int)(cell / ((long long)P.N[0] * P.N[1]));
  for (int i = 0; i < P.n; i++) {
    int l[3] = {0, 0, 0}, ii = i;
    for (int d = 0; d < P.dim; d++) {
      l[d] = P.k * c[d] + ii % P.n1;
      ii /= P.n1;
    }
    idx[i] = (uint64_t)qk_lattice_index(L, l);
  }
}

inline std::vector<uint64_t> host_constrained_dofs(const DevParams& P, const int8_t* bctype) {
  std::vector<uint64_t> out;
  if (P.dg) {
    // P0ParallelConstraints: all DOFs of cells with a processor intersection
    for (long long cell = 0; cell < P.ncells; cell++) {
      int c[3] = {(int)(cell % P.N[0]), (int)((cell / P.N[0]) % P.N[1]), (int)(cell / ((long long)P.N[0] * P.N[1]))};
      bool con = false;
      for (int d = 0; d < P.dim; d++) {
        if (c[d] == 0 && P.side_kind[d][0] == PDB200_SIDE_PROCESSOR) con = true;
        if (c[d] == P.N[d] - 1 && P.side_kind[d][1] == PDB200_SIDE_PROCESSOR) con = true;
      }
      if (con)
        for (int i = 0; i < P.n; i++) out.push_back((uint64_t)(cell * P.n + i));
    }
    return out;
  }
  // conforming: DOFs on boundary faces that are Dirichlet at the face centre, or on processor faces
  QkLayout L = make_qk_layout(P);
  std::vector<uint64_t> tmp;
  for (int dir = 0; dir < P.dim; dir++)
    for (int side = 0; side < 2; side++) {
      const bool processor = P.side_kind[dir][side] == PDB200_SIDE_PROCESSOR;
      const long long nfaces = P.ncells / P.N[dir];
      for (long long f = 0; f < nfaces; f++) {
        const int bct = bctype ? (int)bctype[P.bf_off[dir][side] + f] : (int)PDB200_BC_DIRICHLET;
        if (!processor && bct != PDB200_BC_DIRICHLET) continue;
        // tangential cell coordinates of the face
        int c[3] = {0, 0, 0};
        long long ff = f;
        for (int d = 0; d < P.dim; d++)
          if (d != dir) {
            c[d] = (int)(ff % P.N[d]);
            ff /= P.N[d];
          }
        c[dir] = side ? P.N[dir] - 1 : 0;
        int nt = 1;
        for (int d = 0; d < P.dim - 1; d++) nt *= P.n1;
        for (int t = 0; t < nt; t++) {
          int l[3] = {0, 0, 0}, tt = t;
          for (int d = 0; d < P.dim; d++)
            if (d != dir) {
              l[d] = P.k * c[d] + tt % P.n1;
              tt /= P.n1;
            }
          l[dir] = side ? P.k * P.N[dir] : 0;
          tmp.push_back((uint64_t)qk_lattice_index(L, l));
        }
      }
    }
  std::sort(tmp.begin(), tmp.end());
  tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
  return tmp;
}

}  // namespace pdb

// pdelab_oracle.cc — CPU restatement of the reference's operator-evaluation path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (dune-pdelab_b200/) may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / CPU baseline.
//
// PARITY PINNING: "parity unpinned" at the third-party boundaries.  The reference (dune-pdelab
// 2.10-git) cannot be compiled here (dune-common/-geometry/-grid/-istl/-localfunctions/-typetree/
// -functions are absent) and its tests hold no golden vectors for this path.  This
// file restates, loop for loop and in the reference's evaluation order, the code cited at each
// function; the pieces that live in un-vendored DUNE core modules (>= 2.10, unpinned master in
// .gitlab-ci.yml:24-33) are restated from their published definitions:
//   * dune-geometry QuadratureRules<ctype,dim>::rule(cube, order, GaussLegendre): tensor-product
//     Gauss-Legendre, m = order/2+1 points per direction (call sites
//     common/quadraturerules.hh:117-120, convectiondiffusiondg.hh:140,361,746,1062);
//   * dune-grid YaspGrid: lexicographic cells, intersections 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z,
//     sub-entity numbering grouped by extension bitset; axis-aligned geometry closed forms;
//   * dune-localfunctions LagrangeCubeLocalFiniteElement<k>: lexicographic tensor Lagrange basis,
//     one DOF per sub-entity for k<=2;
//   * dune-istl BCRSMatrix::setIndices: ascending column indices per row.
// What IS pinned: the reference's own invariants (tests/test_reference_invariants.py restates
// test/testconvectiondiffusiondg.cc, test/testmatrixfree.cc, test/testfastdgassembler.cc,
// test/matrixfree/matrix_free_linear.cc, test/test-blocked-istl-ordering.cc) and an independent
// numpy/scipy assembly of the same weak forms (tests/test_oracle_vs_numpy.py).
//
// All paths in comments are relative to /root/reference/dune/pdelab/.

#include "../include/pdelab_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------
// 1-D tables
// ---------------------------------------------------------------------------------------------

// finiteelement/qkdglagrange.hh:55-62  (p) and :65-79 (dp): Lagrange polynomials on nodes j/k.
double lagrange_p(int k, int i, double x) {
  double result = 1.0;
  for (int j = 0; j <= k; j++)
    if (j != i) result *= (k * x - j) / (i - j);
  return result;
}
double lagrange_dp(int k, int i, double x) {
  double result = 0.0;
  for (int j = 0; j <= k; j++)
    if (j != i) {
      double prod = (k * 1.0) / (i - j);
      for (int l = 0; l <= k; l++)
        if (l != i && l != j) prod *= (k * x - l) / (i - l);
      result += prod;
    }
  return result;
}

// finiteelement/qkdglegendre.hh:76-139: shifted Legendre polynomials by the three-term recurrence, values and
// derivatives of all polynomials up to degree k (LegendrePolynomials1d::pdp; k = 0, 1 are the specialisations :142-253)
void legendre_pdp(int k, double x, double* value, double* derivative) {
  value[0] = 1;
  derivative[0] = 0.0;
  if (k < 1) return;
  value[1] = 2 * x - 1;
  derivative[1] = 2.0;
  for (int n = 2; n <= k; n++) {
    value[n] = ((2 * n - 1) * (2 * x - 1) * value[n - 1] - (n - 1) * value[n - 2]) / n;
    derivative[n] = (2 * x - 1) * derivative[n - 1] + 2 * n * value[n - 1];
  }
}
// finiteelement/qkdglobatto.hh:28-66: the k+1 Gauss-Lobatto points of [0,1] (dune-geometry GaussLobatto rule, un-vendored:
// the mathematically unique point set, taken in ASCENDING order — the constructor re-sorts the rule's point pairs so
// that the lower half lies below 1/2; the order inside a half is the rule's, assumed outermost pair first.  Parity of the
// DOF order is unpinned for k >= 3).
void lobatto_points(int k, double* xi) {
  double t[5] = {0, 0, 0, 0, 0};
  switch (k) {
    case 1: t[0] = -1, t[1] = 1; break;
    case 2: t[0] = -1, t[1] = 0, t[2] = 1; break;
    case 3: t[0] = -1, t[1] = -std::sqrt(0.2), t[2] = std::sqrt(0.2), t[3] = 1; break;
    default: t[0] = -1, t[1] = -std::sqrt(3.0 / 7.0), t[2] = 0, t[3] = std::sqrt(3.0 / 7.0), t[4] = 1; break;
  }
  for (int i = 0; i <= k; i++) xi[i] = (1 + t[i]) / 2;
}
// qkdglobatto.hh:69-93: Lagrange polynomials through the Gauss-Lobatto points
double lobatto_p(int k, const double* xi, int i, double x) {
  double result = 1.0;
  for (int j = 0; j <= k; j++)
    if (j != i) result *= (x - xi[j]) / (xi[i] - xi[j]);
  return result;
}
double lobatto_dp(int k, const double* xi, int i, double x) {
  double result = 0.0;
  for (int j = 0; j <= k; j++)
    if (j != i) {
      double prod = 1.0 / (xi[i] - xi[j]);
      for (int l = 0; l <= k; l++)
        if (l != i && l != j) prod *= (x - xi[l]) / (xi[i] - xi[l]);
      result += prod;
    }
  return result;
}

// Gauss-Legendre rule with m points on [0,1], ascending abscissae (dune-geometry tabulates the
// same mathematically unique rule; restated with Newton iteration in long double).
void gauss_legendre(int m, std::vector<double>& x, std::vector<double>& w) {
  x.resize(m);
  w.resize(m);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < m; i++) {
    long double t = std::cos(pi * (i + 0.75L) / (m + 0.5L));  // root of P_m on [-1,1], descending
    long double dp = 0;
    for (int it = 0; it < 100; it++) {
      long double p0 = 1, p1 = t;
      for (int j = 2; j <= m; j++) {
        long double p2 = ((2 * j - 1) * t * p1 - (j - 1) * p0) / j;
        p0 = p1;
        p1 = p2;
      }
      if (m == 0) p1 = 1;
      dp = m * (t * p1 - p0) / (t * t - 1);
      long double dt = p1 / dp;
      t -= dt;
      if (std::fabs((double)dt) < 1e-19) break;
    }
    {  // recompute derivative at converged root
      long double p0 = 1, p1 = t;
      for (int j = 2; j <= m; j++) {
        long double p2 = ((2 * j - 1) * t * p1 - (j - 1) * p0) / j;
        p0 = p1;
        p1 = p2;
      }
      dp = m * (t * p1 - p0) / (t * t - 1);
    }
    long double wi = 2 / ((1 - t * t) * dp * dp);
    x[m - 1 - i] = (double)((1 + t) / 2);
    w[m - 1 - i] = (double)(wi / 2);
  }
}

struct Tables {
  int k, n1, m, npts;          // npts = m + 2 : Gauss points, then xi=0, then xi=1
  std::vector<double> xq, wq;  // Gauss
  std::vector<double> P, DP;   // [npts][n1]
  double p(int pt, int i) const { return P[pt * n1 + i]; }
  double dp(int pt, int i) const { return DP[pt * n1 + i]; }
};

Tables make_tables(int k, int intorder, int basis = PDB200_BASIS_LAGRANGE) {
  Tables T;
  T.k = k;
  T.n1 = k + 1;
  T.m = intorder / 2 + 1;
  gauss_legendre(T.m, T.xq, T.wq);
  T.npts = T.m + 2;
  T.P.resize(T.npts * T.n1);
  T.DP.resize(T.npts * T.n1);
  for (int pt = 0; pt < T.npts; pt++) {
    double x = pt < T.m ? T.xq[pt] : (pt == T.m ? 0.0 : 1.0);
    double lv[10], ld[10], xi[10];
    if (basis == PDB200_BASIS_LEGENDRE) legendre_pdp(k, x, lv, ld);
    if (basis == PDB200_BASIS_LOBATTO) lobatto_points(k, xi);
    for (int i = 0; i < T.n1; i++) {
      if (basis == PDB200_BASIS_LEGENDRE) {
        T.P[pt * T.n1 + i] = lv[i];
        T.DP[pt * T.n1 + i] = ld[i];
      } else if (basis == PDB200_BASIS_LOBATTO) {
        T.P[pt * T.n1 + i] = lobatto_p(k, xi, i, x);
        T.DP[pt * T.n1 + i] = lobatto_dp(k, xi, i, x);
      } else {
        T.P[pt * T.n1 + i] = lagrange_p(k, i, x);
        T.DP[pt * T.n1 + i] = lagrange_dp(k, i, x);
      }
    }
  }
  return T;
}

// ---------------------------------------------------------------------------------------------
// structured grid (YaspGrid restated) and problem view
// ---------------------------------------------------------------------------------------------

struct Ctx {
  const pdb200_problem* p;
  int dim, k, n1, n;  // n = (k+1)^dim local DOFs
  int N[3];
  double h[3], lo[3];
  long ncells;
  long nbf;              // boundary faces
  long bf_off[3][2];     // first boundary face of (dir, side)
  Tables T;
  int nq, nfq;           // volume / face quadrature points
  int np;                // sample points per cell of the point-wise coefficient layouts: nq + 2 dim nfq
  bool pwA, pwB, pwC, pwBC;  // which coefficient call-backs are sampled per quadrature point (pdelab_b200.h)
  double theta;
  double vol;            // |K|
  bool dg;

  explicit Ctx(const pdb200_problem* p_) : p(p_) {
    dim = p->dim;
    k = p->degree;
    if (dim != 2 && dim != 3) throw std::runtime_error("dim must be 2 or 3");
    if (k < 1 || k > 8) throw std::runtime_error("degree must be in 1..8");
    dg = p->space == PDB200_SPACE_QKDG;
    if (!dg && k > 2) throw std::runtime_error("conforming Qk supports k in {1,2} (finiteelementmap/qkfem.hh:17-78)");
    n1 = k + 1;
    n = 1;
    ncells = 1;
    vol = 1.0;
    for (int d = 0; d < 3; d++) {
      N[d] = d < dim ? p->cells[d] : 1;
      lo[d] = d < dim ? p->lower[d] : 0.0;
      h[d] = d < dim ? (p->upper[d] - p->lower[d]) / N[d] : 1.0;
      if (d < dim) {
        n *= n1;
        ncells *= N[d];
        vol *= h[d];
      }
    }
    // convectiondiffusiondg.hh:139 intorder = intorderadd + quadrature_factor*order (factor 2);
    // convectiondiffusionfem.hh:93 intorder = intorderadd + 2*order
    if (p->basis < PDB200_BASIS_LAGRANGE || p->basis > PDB200_BASIS_LOBATTO) throw std::runtime_error("unknown QkDG basis");
    if (!dg && p->basis != PDB200_BASIS_LAGRANGE) throw std::runtime_error("conforming Qk spaces use the Lagrange basis");
    if (p->basis == PDB200_BASIS_LOBATTO && k > 4) throw std::runtime_error("Gauss-Lobatto points are tabulated for k <= 4");
    T = make_tables(k, p->intorderadd + 2 * k, p->basis);
    nq = 1;
    nfq = 1;
    for (int d = 0; d < dim; d++) nq *= T.m;
    for (int d = 0; d < dim - 1; d++) nfq *= T.m;
    nbf = 0;
    for (int d = 0; d < dim; d++)
      for (int s = 0; s < 2; s++) {
        bf_off[d][s] = nbf;
        nbf += ncells / N[d];
      }
    np = nq + 2 * dim * nfq;
    pwA = (p->pointwise & PDB200_POINTWISE_A) && p->a_mode != PDB200_A_IDENTITY;
    pwB = (p->pointwise & PDB200_POINTWISE_B) && p->b;
    pwC = (p->pointwise & PDB200_POINTWISE_C) && p->c;
    pwBC = (p->pointwise & PDB200_POINTWISE_BCTYPE) && p->bctype;
    if (pwBC && !dg)
      throw std::runtime_error("ConvectionDiffusionFEM evaluates bctype at the face centre (convectiondiffusionfem.hh:226-229)");
    // convectiondiffusiondg.hh:99-101
    theta = 1.0;
    if (p->dg_method == PDB200_DG_SIPG) theta = -1.0;
    if (p->dg_method == PDB200_DG_IIPG) theta = 0.0;
  }

  long cell_index(const int c[3]) const { return c[0] + (long)N[0] * (c[1] + (long)N[1] * c[2]); }
  void cell_coord(long e, int c[3]) const {
    c[0] = (int)(e % N[0]);
    e /= N[0];
    c[1] = (int)(e % N[1]);
    c[2] = (int)(e / N[1]);
  }
  // boundary face number of the face (dir, side) of the cell with coordinates c
  long bface(const int c[3], int dir, int side) const {
    long idx = 0, stride = 1;
    for (int d = 0; d < dim; d++)
      if (d != dir) {
        idx += stride * c[d];
        stride *= N[d];
      }
    return bf_off[dir][side] + idx;
  }
  // sample-point number (inside its cell) of face quadrature point q of face (dir, side): pdelab_b200.h
  int face_pt(int dir, int side, int q) const { return nq + (2 * dir + side) * nfq + q; }
  void A_entry(long e, double out[3][3]) const {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) out[i][j] = 0.0;
    switch (p->a_mode) {
      case PDB200_A_IDENTITY:
        for (int i = 0; i < dim; i++) out[i][i] = 1.0;
        break;
      case PDB200_A_SCALAR:
        for (int i = 0; i < dim; i++) out[i][i] = p->A[e];
        break;
      case PDB200_A_DIAGONAL:
        for (int i = 0; i < dim; i++) out[i][i] = p->A[e * dim + i];
        break;
      default:
        for (int i = 0; i < dim; i++)
          for (int j = 0; j < dim; j++) out[i][j] = p->A[e * dim * dim + i * dim + j];
    }
  }
  // param.A(cell, localcenter) — with permeabilityIsConstantPerCell() == false the value is replaced at every
  // quadrature point before it is used (convectiondiffusiondg.hh:143-146, 367-382, 752-760), so the point-wise
  // layout holds no centre sample: sample point 0 stands in for it
  void A(long cell, double out[3][3]) const { A_entry(pwA ? cell * np : cell, out); }
  // param.A(cell, x_pt) for sample point pt of the cell (point-wise layout only)
  void A_at(long cell, int pt, double out[3][3]) const { A_entry(cell * np + pt, out); }
  // param.b(cell, x_pt): cell-wise constant field or the sample at point pt
  void b(long cell, int pt, double out[3]) const {
    const long e = pwB ? cell * np + pt : cell;
    for (int d = 0; d < 3; d++) out[d] = (p->b && d < dim) ? p->b[e * dim + d] : 0.0;
  }
  // param.c(cell, x_q) at volume point q
  double c(long cell, int q) const { return p->c ? p->c[pwC ? cell * nq + q : cell] : 0.0; }
  double f(long cell, int q) const { return p->f ? p->f[cell * nq + q] : 0.0; }
  // param.bctype(intersection, x_q): per face or per face quadrature point
  int bctype(long bf, int q = 0) const {
    return p->bctype ? (int)p->bctype[pwBC ? bf * nfq + q : bf] : (int)PDB200_BC_DIRICHLET;
  }
  double g(long bf, int q) const { return p->g ? p->g[bf * nfq + q] : 0.0; }
  double j(long bf, int q) const { return p->j ? p->j[bf * nfq + q] : 0.0; }
  double o(long bf, int q) const { return p->o ? p->o[bf * nfq + q] : 0.0; }

  // QkLocalBasis::evaluateFunction / evaluateJacobian, finiteelement/qkdglagrange.hh:156-200,
  // followed by jac.mv with jacobianInverseTransposed = diag(1/h) (convectiondiffusiondg.hh:162-167).
  // pt[d] indexes the 1-D point table (Gauss points, then 0, then 1).
  void eval_basis(const int pt[3], double* phi, double* grad /* [n][3] physical */) const {
    for (int i = 0; i < n; i++) {
      int alpha[3], ii = i;
      for (int d = 0; d < dim; d++) {
        alpha[d] = ii % n1;
        ii /= n1;
      }
      double v = 1.0;
      for (int d = 0; d < dim; d++) v *= T.p(pt[d], alpha[d]);
      phi[i] = v;
      for (int d = 0; d < dim; d++) {
        double gd = T.dp(pt[d], alpha[d]);
        for (int l = 0; l < dim; l++)
          if (l != d) gd *= T.p(pt[l], alpha[l]);
        grad[i * 3 + d] = (1.0 / h[d]) * gd;
      }
      for (int d = dim; d < 3; d++) grad[i * 3 + d] = 0.0;
    }
  }
  // volume quadrature point q -> 1-D point indices and weight
  double vol_point(int q, int pt[3]) const {
    double w = 1.0;
    for (int d = 0; d < 3; d++) pt[d] = 0;
    for (int d = 0; d < dim; d++) {
      pt[d] = q % T.m;
      q /= T.m;
      w *= T.wq[pt[d]];
    }
    return w;
  }
  // face quadrature point q of a face normal to dir -> 1-D point indices of the tangential
  // coordinates (entry dir is left to the caller) and weight
  double face_point(int q, int dir, int pt[3]) const {
    double w = 1.0;
    for (int d = 0; d < 3; d++) pt[d] = 0;
    for (int d = 0; d < dim; d++)
      if (d != dir) {
        pt[d] = q % T.m;
        q /= T.m;
        w *= T.wq[pt[d]];
      }
    return w;
  }
  double face_area(int dir) const {
    double a = 1.0;
    for (int d = 0; d < dim; d++)
      if (d != dir) a *= h[d];
    return a;
  }
};

inline double dot3(const double* a, const double* b, int dim) {
  double s = 0.0;
  for (int d = 0; d < dim; d++) s += a[d] * b[d];
  return s;
}
inline void matvec3(const double A[3][3], const double* x, double* y, int dim) {
  for (int i = 0; i < 3; i++) y[i] = 0.0;
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) y[i] += A[i][j] * x[j];
}

// column-major local matrix, gridoperator/common/localmatrix.hh:437  _container[j*rows+i]
struct LocalMatrix {
  int rows, cols;
  std::vector<double> a;
  void assign(int r, int c) {
    rows = r;
    cols = c;
    a.assign((size_t)r * c, 0.0);
  }
  void accumulate(int i, int j, double v) { a[(size_t)j * rows + i] += v; }
  double operator()(int i, int j) const { return a[(size_t)j * rows + i]; }
};

// ---------------------------------------------------------------------------------------------
// ConvectionDiffusionDG  (localoperator/convectiondiffusiondg.hh)
// ---------------------------------------------------------------------------------------------

struct Scratch {
  std::vector<double> phi_s, phi_n, grad_s, grad_n;
  explicit Scratch(int n) : phi_s(n), phi_n(n), grad_s(3 * n), grad_n(3 * n) {}
};

// alpha_volume, convectiondiffusiondg.hh:106-188 (with_f=false);
// ConvectionDiffusionFEM::alpha_volume, convectiondiffusionfem.hh:63-136 (with_f=true)
void alpha_volume(const Ctx& C, long cell, const double* x, double* r, Scratch& S, bool with_f) {
  double A[3][3], b[3];
  C.A(cell, A);
  double* phi = S.phi_s.data();
  double* gradphi = S.grad_s.data();
  for (int q = 0; q < C.nq; q++) {
    int pt[3];
    double weight = C.vol_point(q, pt);
    if (C.pwA) C.A_at(cell, q, A);  // !permeabilityIsConstantPerCell: :143-146 / convectiondiffusionfem.hh:97-100
    C.eval_basis(pt, phi, gradphi);
    double u = 0.0;
    for (int i = 0; i < C.n; i++) u += x[i] * phi[i];
    double gradu[3] = {0, 0, 0}, Agradu[3];
    for (int i = 0; i < C.n; i++)
      for (int d = 0; d < C.dim; d++) gradu[d] += x[i] * gradphi[i * 3 + d];
    matvec3(A, gradu, Agradu, C.dim);
    C.b(cell, q, b);          // param.b(cell, ip.position()), :178
    double c = C.c(cell, q);  // param.c(cell, ip.position()), :181
    double factor = weight * C.vol;
    if (with_f) {
      double f = C.f(cell, q);
      for (int i = 0; i < C.n; i++)
        r[i] += (dot3(Agradu, &gradphi[i * 3], C.dim) - u * dot3(b, &gradphi[i * 3], C.dim) +
                 (c * u - f) * phi[i]) * factor;
    } else {
      for (int i = 0; i < C.n; i++)
        r[i] += (dot3(Agradu, &gradphi[i * 3], C.dim) - u * dot3(b, &gradphi[i * 3], C.dim) +
                 c * u * phi[i]) * factor;
    }
  }
}

// jacobian_volume, convectiondiffusiondg.hh:199-266 / convectiondiffusionfem.hh:140-203
// only_row >= 0 restricts the accumulation to that test function (row) — every entry has its own accumulation
// sequence, so the kept row is bit-identical to the full local matrix's (used by the sampled-row Jacobian)
void jacobian_volume(const Ctx& C, long cell, LocalMatrix& mat, Scratch& S, int only_row = -1) {
  double A[3][3], b[3];
  C.A(cell, A);
  double* phi = S.phi_s.data();
  double* gradphi = S.grad_s.data();
  std::vector<double> Agradphi(3 * C.n);
  for (int q = 0; q < C.nq; q++) {
    int pt[3];
    double weight = C.vol_point(q, pt);
    if (C.pwA) C.A_at(cell, q, A);  // :234-237
    C.eval_basis(pt, phi, gradphi);
    for (int i = 0; i < C.n; i++) matvec3(A, &gradphi[i * 3], &Agradphi[i * 3], C.dim);
    C.b(cell, q, b);          // :255
    double c = C.c(cell, q);  // :258
    double factor = weight * C.vol;
    for (int j = 0; j < C.n; j++)
      for (int i = 0; i < C.n; i++) {
        if (only_row >= 0 && i != only_row) continue;
        mat.accumulate(i, j, (dot3(&Agradphi[j * 3], &gradphi[i * 3], C.dim) -
                              phi[j] * dot3(b, &gradphi[i * 3], C.dim) + c * phi[j] * phi[i]) * factor);
      }
  }
}

// lambda_volume, convectiondiffusiondg.hh:1048-1075
void dg_lambda_volume(const Ctx& C, long cell, double* r, Scratch& S) {
  double* phi = S.phi_s.data();
  double* gradphi = S.grad_s.data();
  for (int q = 0; q < C.nq; q++) {
    int pt[3];
    double weight = C.vol_point(q, pt);
    C.eval_basis(pt, phi, gradphi);
    double f = C.f(cell, q);
    double factor = weight * C.vol;
    for (int i = 0; i < C.n; i++) r[i] += -f * phi[i] * factor;
  }
}

struct FaceCoef {
  double An_s[3], An_n[3], omega_s, omega_n, penalty, nF[3];
};

// coefficient set-up of alpha_skeleton / jacobian_skeleton, convectiondiffusiondg.hh:303-346
FaceCoef skeleton_coef(const Ctx& C, long cell_s, long cell_n, int dir) {
  FaceCoef F;
  double A_s[3][3], A_n[3][3];
  C.A(cell_s, A_s);
  C.A(cell_n, A_n);
  double area = C.face_area(dir);
  double h_F = std::min(C.vol, C.vol) / area;
  for (int d = 0; d < 3; d++) F.nF[d] = 0.0;
  F.nF[dir] = -1.0;  // outer normal of the inside (= higher index) cell towards cell - e_dir
  matvec3(A_s, F.nF, F.An_s, C.dim);
  matvec3(A_n, F.nF, F.An_n, C.dim);
  double harmonic_average;
  if (C.p->dg_weights == PDB200_DG_WEIGHTS_ON) {
    double delta_s = dot3(F.An_s, F.nF, C.dim);
    double delta_n = dot3(F.An_n, F.nF, C.dim);
    F.omega_s = delta_n / (delta_s + delta_n + 1e-20);
    F.omega_n = delta_s / (delta_s + delta_n + 1e-20);
    harmonic_average = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20);
  } else {
    F.omega_s = F.omega_n = 0.5;
    harmonic_average = 1.0;
  }
  int degree = C.k;
  F.penalty = (C.p->dg_alpha / h_F) * harmonic_average * degree * (degree + C.dim - 1);
  return F;
}

// "update all variables dependent on A if A is not cell-wise constant", convectiondiffusiondg.hh:367-382 (alpha_skeleton)
// and :575-590 (jacobian_skeleton): A_s, A_n at face quadrature point q; the weights and the penalty factor only
// with weightsOn
void skeleton_coef_at_point(const Ctx& C, FaceCoef& F, long cell_s, long cell_n, int dir, int q) {
  double A_s[3][3], A_n[3][3];
  C.A_at(cell_s, C.face_pt(dir, 0, q), A_s);  // geo_in_inside.global(ip): lower face of the inside cell
  C.A_at(cell_n, C.face_pt(dir, 1, q), A_n);  // geo_in_outside.global(ip): upper face of the outside cell
  matvec3(A_s, F.nF, F.An_s, C.dim);
  matvec3(A_n, F.nF, F.An_n, C.dim);
  if (C.p->dg_weights == PDB200_DG_WEIGHTS_ON) {
    double area = C.face_area(dir);
    double h_F = std::min(C.vol, C.vol) / area;
    double delta_s = dot3(F.An_s, F.nF, C.dim);
    double delta_n = dot3(F.An_n, F.nF, C.dim);
    F.omega_s = delta_n / (delta_s + delta_n + 1e-20);
    F.omega_n = delta_s / (delta_s + delta_n + 1e-20);
    double harmonic_average = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20);
    int degree = C.k;
    F.penalty = (C.p->dg_alpha / h_F) * harmonic_average * degree * (degree + C.dim - 1);
  }
}

// alpha_skeleton, convectiondiffusiondg.hh:271-471.  Face between the inside cell cell_s and
// cell_n = cell_s - e_dir (the assembler visits a face from the cell with the larger index,
// gridoperator/default/assembler.hh:178-184).
void dg_alpha_skeleton(const Ctx& C, long cell_s, long cell_n, int dir, const double* x_s,
                       const double* x_n, double* r_s, double* r_n, Scratch& S) {
  FaceCoef F = skeleton_coef(C, cell_s, cell_n, dir);
  double area = C.face_area(dir);
  double b[3];
  double *phi_s = S.phi_s.data(), *phi_n = S.phi_n.data();
  double *tg_s = S.grad_s.data(), *tg_n = S.grad_n.data();
  for (int q = 0; q < C.nfq; q++) {
    int pt_s[3], pt_n[3];
    double weight = C.face_point(q, dir, pt_s);
    if (C.pwA) skeleton_coef_at_point(C, F, cell_s, cell_n, dir, q);  // :367-382
    for (int d = 0; d < 3; d++) pt_n[d] = pt_s[d];
    pt_s[dir] = C.T.m;      // xi_dir = 0 in the inside cell
    pt_n[dir] = C.T.m + 1;  // xi_dir = 1 in the outside cell
    C.eval_basis(pt_s, phi_s, tg_s);
    C.eval_basis(pt_n, phi_n, tg_n);
    double u_s = 0.0, u_n = 0.0;
    for (int i = 0; i < C.n; i++) u_s += x_s[i] * phi_s[i];
    for (int i = 0; i < C.n; i++) u_n += x_n[i] * phi_n[i];
    double gradu_s[3] = {0, 0, 0}, gradu_n[3] = {0, 0, 0};
    for (int i = 0; i < C.n; i++)
      for (int d = 0; d < C.dim; d++) gradu_s[d] += x_s[i] * tg_s[i * 3 + d];
    for (int i = 0; i < C.n; i++)
      for (int d = 0; d < C.dim; d++) gradu_n[d] += x_n[i] * tg_n[i * 3 + d];
    C.b(cell_s, C.face_pt(dir, 0, q), b);  // param.b(cell_inside, iplocal_s), :426
    double normalflux = dot3(b, F.nF, C.dim);
    double omegaup_s, omegaup_n;
    if (normalflux >= 0.0) {
      omegaup_s = 1.0;
      omegaup_n = 0.0;
    } else {
      omegaup_s = 0.0;
      omegaup_n = 1.0;
    }
    double factor = weight * area;
    double term1 = (omegaup_s * u_s + omegaup_n * u_n) * normalflux * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term1 * phi_s[i];
    for (int i = 0; i < C.n; i++) r_n[i] += -term1 * phi_n[i];
    double term2 = -(F.omega_s * dot3(F.An_s, gradu_s, C.dim) + F.omega_n * dot3(F.An_n, gradu_n, C.dim)) * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term2 * phi_s[i];
    for (int i = 0; i < C.n; i++) r_n[i] += -term2 * phi_n[i];
    double term3 = (u_s - u_n) * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term3 * C.theta * F.omega_s * dot3(F.An_s, &tg_s[i * 3], C.dim);
    for (int i = 0; i < C.n; i++) r_n[i] += term3 * C.theta * F.omega_n * dot3(F.An_n, &tg_n[i * 3], C.dim);
    double term4 = F.penalty * (u_s - u_n) * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term4 * phi_s[i];
    for (int i = 0; i < C.n; i++) r_n[i] += -term4 * phi_n[i];
  }
}

// jacobian_skeleton, convectiondiffusiondg.hh:484-669
void dg_jacobian_skeleton(const Ctx& C, long cell_s, long cell_n, int dir, LocalMatrix& mat_ss,
                          LocalMatrix& mat_sn, LocalMatrix& mat_ns, LocalMatrix& mat_nn, Scratch& S) {
  FaceCoef F = skeleton_coef(C, cell_s, cell_n, dir);
  double area = C.face_area(dir);
  double b[3];
  double *phi_s = S.phi_s.data(), *phi_n = S.phi_n.data();
  double *tg_s = S.grad_s.data(), *tg_n = S.grad_n.data();
  const double theta = C.theta;
  for (int q = 0; q < C.nfq; q++) {
    int pt_s[3], pt_n[3];
    double weight = C.face_point(q, dir, pt_s);
    if (C.pwA) skeleton_coef_at_point(C, F, cell_s, cell_n, dir, q);  // :575-590
    for (int d = 0; d < 3; d++) pt_n[d] = pt_s[d];
    pt_s[dir] = C.T.m;
    pt_n[dir] = C.T.m + 1;
    C.eval_basis(pt_s, phi_s, tg_s);
    C.eval_basis(pt_n, phi_n, tg_n);
    C.b(cell_s, C.face_pt(dir, 0, q), b);  // :613
    double normalflux = dot3(b, F.nF, C.dim);
    double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0;
    double omegaup_n = normalflux >= 0.0 ? 0.0 : 1.0;
    double factor = weight * area;
    double ipfactor = F.penalty * factor;
    const double omega_s = F.omega_s, omega_n = F.omega_n;
    for (int j = 0; j < C.n; j++) {
      double temp1 = -dot3(F.An_s, &tg_s[j * 3], C.dim) * omega_s * factor;
      for (int i = 0; i < C.n; i++) {
        mat_ss.accumulate(i, j, omegaup_s * phi_s[j] * normalflux * factor * phi_s[i]);
        mat_ss.accumulate(i, j, temp1 * phi_s[i]);
        mat_ss.accumulate(i, j, phi_s[j] * factor * theta * omega_s * dot3(F.An_s, &tg_s[i * 3], C.dim));
        mat_ss.accumulate(i, j, phi_s[j] * ipfactor * phi_s[i]);
      }
    }
    for (int j = 0; j < C.n; j++) {
      double temp1 = -dot3(F.An_n, &tg_n[j * 3], C.dim) * omega_n * factor;
      for (int i = 0; i < C.n; i++) {
        mat_sn.accumulate(i, j, omegaup_n * phi_n[j] * normalflux * factor * phi_s[i]);
        mat_sn.accumulate(i, j, temp1 * phi_s[i]);
        mat_sn.accumulate(i, j, -phi_n[j] * factor * theta * omega_s * dot3(F.An_s, &tg_s[i * 3], C.dim));
        mat_sn.accumulate(i, j, -phi_n[j] * ipfactor * phi_s[i]);
      }
    }
    for (int j = 0; j < C.n; j++) {
      double temp1 = -dot3(F.An_s, &tg_s[j * 3], C.dim) * omega_s * factor;
      for (int i = 0; i < C.n; i++) {
        mat_ns.accumulate(i, j, -omegaup_s * phi_s[j] * normalflux * factor * phi_n[i]);
        mat_ns.accumulate(i, j, -temp1 * phi_n[i]);
        mat_ns.accumulate(i, j, phi_s[j] * factor * theta * omega_n * dot3(F.An_n, &tg_n[i * 3], C.dim));
        mat_ns.accumulate(i, j, -phi_s[j] * ipfactor * phi_n[i]);
      }
    }
    for (int j = 0; j < C.n; j++) {
      double temp1 = -dot3(F.An_n, &tg_n[j * 3], C.dim) * omega_n * factor;
      for (int i = 0; i < C.n; i++) {
        mat_nn.accumulate(i, j, -omegaup_n * phi_n[j] * normalflux * factor * phi_n[i]);
        mat_nn.accumulate(i, j, -temp1 * phi_n[i]);
        mat_nn.accumulate(i, j, -phi_n[j] * factor * theta * omega_n * dot3(F.An_n, &tg_n[i * 3], C.dim));
        mat_nn.accumulate(i, j, phi_n[j] * ipfactor * phi_n[i]);
      }
    }
  }
}

struct BndCoef {
  double An_s[3], penalty, nF[3];
};
// coefficient set-up of the boundary integrals, convectiondiffusiondg.hh:710-734
BndCoef boundary_coef(const Ctx& C, long cell, int dir, int side) {
  BndCoef F;
  double A_s[3][3];
  C.A(cell, A_s);
  double h_F = C.vol / C.face_area(dir);
  for (int d = 0; d < 3; d++) F.nF[d] = 0.0;
  F.nF[dir] = side ? 1.0 : -1.0;
  matvec3(A_s, F.nF, F.An_s, C.dim);
  double harmonic_average = C.p->dg_weights == PDB200_DG_WEIGHTS_ON ? dot3(F.An_s, F.nF, C.dim) : 1.0;
  int degree = C.k;
  F.penalty = (C.p->dg_alpha / h_F) * harmonic_average * degree * (degree + C.dim - 1);
  return F;
}

// :752-760 (residual_boundary_integral) and :967-977 (jacobian_boundary): A_s at face quadrature point q
void boundary_coef_at_point(const Ctx& C, BndCoef& F, long cell, int dir, int side, int q) {
  double A_s[3][3];
  C.A_at(cell, C.face_pt(dir, side, q), A_s);
  matvec3(A_s, F.nF, F.An_s, C.dim);
  if (C.p->dg_weights == PDB200_DG_WEIGHTS_ON) {
    double h_F = C.vol / C.face_area(dir);
    double harmonic_average = dot3(F.An_s, F.nF, C.dim);
    int degree = C.k;
    F.penalty = (C.p->dg_alpha / h_F) * harmonic_average * degree * (degree + C.dim - 1);
  }
}

// residual_boundary_integral, convectiondiffusiondg.hh:684-879 (alpha_boundary :884-889,
// jacobian_apply_boundary :893-899 with jacobian_apply=true).  Returns non-zero on the
// "Outflow boundary condition on inflow" exception (:802-806).
int dg_boundary(const Ctx& C, long cell, const int cc[3], int dir, int side, const double* x_s,
                double* r_s, bool jacobian_apply, Scratch& S) {
  BndCoef F = boundary_coef(C, cell, dir, side);
  double area = C.face_area(dir);
  long bf = C.bface(cc, dir, side);
  double b[3];
  double* phi_s = S.phi_s.data();
  double* tg_s = S.grad_s.data();
  for (int q = 0; q < C.nfq; q++) {
    if (C.pwA) boundary_coef_at_point(C, F, cell, dir, side, q);  // :752-760
    int bctype = C.bctype(bf, q);  // param.bctype(ig.intersection(), ip.position()), :763
    if (bctype == PDB200_BC_NONE) continue;
    int pt[3];
    double weight = C.face_point(q, dir, pt);
    pt[dir] = side ? C.T.m + 1 : C.T.m;
    C.eval_basis(pt, phi_s, tg_s);
    double factor = weight * area;
    if (bctype == PDB200_BC_NEUMANN) {
      if (!jacobian_apply) {
        double j = C.j(bf, q);
        for (int i = 0; i < C.n; i++) r_s[i] += j * phi_s[i] * factor;
      }
      continue;
    }
    double u_s = 0.0;
    for (int i = 0; i < C.n; i++) u_s += x_s[i] * phi_s[i];
    C.b(cell, C.face_pt(dir, side, q), b);  // param.b(cell_inside, iplocal_s), :797
    double normalflux = dot3(b, F.nF, C.dim);
    if (bctype == PDB200_BC_OUTFLOW) {
      if (normalflux < -1e-30) {
        g_err = "Outflow boundary condition on inflow!";
        return 1;
      }
      double term1 = u_s * normalflux * factor;
      for (int i = 0; i < C.n; i++) r_s[i] += term1 * phi_s[i];
      if (!jacobian_apply) {
        double o = C.o(bf, q);
        for (int i = 0; i < C.n; i++) r_s[i] += o * phi_s[i] * factor;
      }
      continue;
    }
    double gradu_s[3] = {0, 0, 0};
    for (int i = 0; i < C.n; i++)
      for (int d = 0; d < C.dim; d++) gradu_s[d] += x_s[i] * tg_s[i * 3 + d];
    double g = C.g(bf, q);
    if (jacobian_apply) g = 0.0;
    double omegaup_s, omegaup_n;
    if (normalflux >= 0.0) {
      omegaup_s = 1.0;
      omegaup_n = 0.0;
    } else {
      omegaup_s = 0.0;
      omegaup_n = 1.0;
    }
    double term1 = (omegaup_s * u_s + omegaup_n * g) * normalflux * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term1 * phi_s[i];
    double term2 = dot3(F.An_s, gradu_s, C.dim) * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += -term2 * phi_s[i];
    double term3 = (u_s - g) * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term3 * C.theta * dot3(F.An_s, &tg_s[i * 3], C.dim);
    double term4 = F.penalty * (u_s - g) * factor;
    for (int i = 0; i < C.n; i++) r_s[i] += term4 * phi_s[i];
  }
  return 0;
}

// jacobian_boundary, convectiondiffusiondg.hh:902-1044
int dg_jacobian_boundary(const Ctx& C, long cell, const int cc[3], int dir, int side,
                         LocalMatrix& mat_ss, Scratch& S) {
  BndCoef F = boundary_coef(C, cell, dir, side);
  double area = C.face_area(dir);
  long bf = C.bface(cc, dir, side);
  double b[3];
  double* phi_s = S.phi_s.data();
  double* tg_s = S.grad_s.data();
  for (int q = 0; q < C.nfq; q++) {
    if (C.pwA) boundary_coef_at_point(C, F, cell, dir, side, q);  // :967-977
    int bctype = C.bctype(bf, q);  // :979
    if (bctype == PDB200_BC_NONE || bctype == PDB200_BC_NEUMANN) continue;
    int pt[3];
    double weight = C.face_point(q, dir, pt);
    pt[dir] = side ? C.T.m + 1 : C.T.m;
    C.eval_basis(pt, phi_s, tg_s);
    double factor = weight * area;
    C.b(cell, C.face_pt(dir, side, q), b);  // :995
    double normalflux = dot3(b, F.nF, C.dim);
    if (bctype == PDB200_BC_OUTFLOW) {
      if (normalflux < -1e-30) {
        g_err = "Outflow boundary condition on inflow!";
        return 1;
      }
      for (int j = 0; j < C.n; j++)
        for (int i = 0; i < C.n; i++) mat_ss.accumulate(i, j, phi_s[j] * normalflux * factor * phi_s[i]);
      continue;
    }
    double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0;
    for (int j = 0; j < C.n; j++)
      for (int i = 0; i < C.n; i++) mat_ss.accumulate(i, j, omegaup_s * phi_s[j] * normalflux * factor * phi_s[i]);
    for (int j = 0; j < C.n; j++)
      for (int i = 0; i < C.n; i++) mat_ss.accumulate(i, j, -dot3(F.An_s, &tg_s[j * 3], C.dim) * factor * phi_s[i]);
    for (int j = 0; j < C.n; j++)
      for (int i = 0; i < C.n; i++)
        mat_ss.accumulate(i, j, phi_s[j] * factor * C.theta * dot3(F.An_s, &tg_s[i * 3], C.dim));
    for (int j = 0; j < C.n; j++)
      for (int i = 0; i < C.n; i++) mat_ss.accumulate(i, j, F.penalty * phi_s[j] * phi_s[i] * factor);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// ConvectionDiffusionFEM boundary terms (localoperator/convectiondiffusionfem.hh)
// ---------------------------------------------------------------------------------------------

// alpha_boundary, convectiondiffusionfem.hh:207-275.  linear_only drops the data terms j, o
// (the exact derivative used for jacobian_apply, see fem_jacobian_apply below).
void fem_alpha_boundary(const Ctx& C, long cell, const int cc[3], int dir, int side,
                        const double* x_s, double* r_s, bool linear_only, Scratch& S) {
  long bf = C.bface(cc, dir, side);
  int bctype = C.bctype(bf);  // evaluated at the face centre (:226-229)
  if (bctype == PDB200_BC_DIRICHLET) return;
  double area = C.face_area(dir);
  double nF[3] = {0, 0, 0}, b[3];
  nF[dir] = side ? 1.0 : -1.0;
  double* phi = S.phi_s.data();
  double* grad = S.grad_s.data();
  for (int q = 0; q < C.nfq; q++) {
    int pt[3];
    double weight = C.face_point(q, dir, pt);
    pt[dir] = side ? C.T.m + 1 : C.T.m;
    C.eval_basis(pt, phi, grad);
    if (bctype == PDB200_BC_NEUMANN && !linear_only) {
      double j = C.j(bf, q);
      double factor = weight * area;
      for (int i = 0; i < C.n; i++) r_s[i] += j * phi[i] * factor;
    }
    if (bctype == PDB200_BC_OUTFLOW) {
      double u = 0.0;
      for (int i = 0; i < C.n; i++) u += x_s[i] * phi[i];
      C.b(cell, C.face_pt(dir, side, q), b);  // param.b(cell_inside, local), convectiondiffusionfem.hh:254
      double o = linear_only ? 0.0 : C.o(bf, q);
      double factor = weight * area;
      for (int i = 0; i < C.n; i++) r_s[i] += (dot3(b, nF, C.dim) * u + o) * phi[i] * factor;
    }
  }
}

// jacobian_boundary, convectiondiffusionfem.hh:279-325
void fem_jacobian_boundary(const Ctx& C, long cell, const int cc[3], int dir, int side,
                           LocalMatrix& mat, Scratch& S) {
  long bf = C.bface(cc, dir, side);
  int bctype = C.bctype(bf);
  if (bctype == PDB200_BC_DIRICHLET) return;
  if (bctype == PDB200_BC_NEUMANN) return;
  double area = C.face_area(dir);
  double nF[3] = {0, 0, 0}, b[3];
  nF[dir] = side ? 1.0 : -1.0;
  double* phi = S.phi_s.data();
  double* grad = S.grad_s.data();
  for (int q = 0; q < C.nfq; q++) {
    int pt[3];
    double weight = C.face_point(q, dir, pt);
    pt[dir] = side ? C.T.m + 1 : C.T.m;
    C.eval_basis(pt, phi, grad);
    C.b(cell, C.face_pt(dir, side, q), b);  // convectiondiffusionfem.hh:314
    double factor = weight * area;
    if (bctype == PDB200_BC_OUTFLOW) {
      for (int j = 0; j < C.n; j++)
        for (int i = 0; i < C.n; i++) mat.accumulate(i, j, dot3(b, nF, C.dim) * phi[j] * phi[i] * factor);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// DOF numbering: LocalFunctionSpace::bind + LFSIndexCache::update + leaf ordering
//   gridfunctionspace/localfunctionspace.hh:616-654, gridfunctionspace/lfsindexcache.hh:603-633,
//   ordering/leaforderingbase.hh:97-203, ordering/leafgridviewordering.hh:166-184
// ---------------------------------------------------------------------------------------------

struct DofMap {
  const Ctx& C;
  long ndofs;
  long codim_block_off[4];  // offset of the block of entities with (dim - codim) = e-dim 0..3
  long group_off[8];        // offset (inside its e-dim block) of the entity group with bitset s
  explicit DofMap(const Ctx& C_) : C(C_) {
    if (C.dg) {
      ndofs = C.ncells * C.n;
      return;
    }
    if (C.k == 1) {
      ndofs = 1;
      for (int d = 0; d < C.dim; d++) ndofs *= C.N[d] + 1;
      return;
    }
    // Q2: one DOF per sub-entity.  Entities of dimension e are grouped by extension bitset s in
    // increasing integer value; inside a group lexicographic, x fastest (YaspGrid index sets).
    long count_by_edim[4] = {0, 0, 0, 0};
    for (int edim = 0; edim <= C.dim; edim++)
      for (int s = 0; s < (1 << C.dim); s++)
        if (__builtin_popcount(s) == edim) {
          group_off[s] = count_by_edim[edim];
          count_by_edim[edim] += group_size(s);
        }
    long off = 0;
    for (int edim = 0; edim <= C.dim; edim++) {
      codim_block_off[edim] = off;  // GlobalGeometryTypeIndex order: vertex < line < quad < hexa
      off += count_by_edim[edim];
    }
    ndofs = off;
  }
  long group_size(int s) const {
    long sz = 1;
    for (int d = 0; d < C.dim; d++) sz *= ((s >> d) & 1) ? C.N[d] : C.N[d] + 1;
    return sz;
  }
  // container index of local DOF i of the cell with coordinates c
  long index(const int c[3], long cell, int i) const {
    if (C.dg) return cell * C.n + i;  // DG: LocalKey(0,0,i), finiteelement/qkdglagrange.hh:233-237
    int lat[3] = {0, 0, 0}, ii = i;
    for (int d = 0; d < C.dim; d++) {
      lat[d] = ii % C.n1;
      ii /= C.n1;
    }
    if (C.k == 1) {
      long idx = 0, stride = 1;
      for (int d = 0; d < C.dim; d++) {
        idx += stride * (c[d] + lat[d]);
        stride *= C.N[d] + 1;
      }
      return idx;
    }
    int s = 0, edim = 0;
    long idx = 0, stride = 1;
    for (int d = 0; d < C.dim; d++) {
      int ext = lat[d] == 1;
      s |= ext << d;
      edim += ext;
      int a = c[d] + (lat[d] == 2);
      idx += stride * a;
      stride *= ext ? C.N[d] : C.N[d] + 1;
    }
    return codim_block_off[edim] + group_off[s] + idx;
  }
};

// constraints(bctype, gfs, cc): constraints/common/constraints.hh:588-687 with
// ConformingDirichletConstraints::boundary (constraints/conforming.hh:53-93),
// OverlappingConformingDirichletConstraints::processor (:108-138) and
// P0ParallelConstraints::processor (constraints/p0.hh:31-41)
std::vector<char> constrained_flags(const Ctx& C, const DofMap& M) {
  std::vector<char> flag(M.ndofs, 0);
  for (long cell = 0; cell < C.ncells; cell++) {
    int c[3];
    C.cell_coord(cell, c);
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++) {
        bool onb = side ? c[dir] == C.N[dir] - 1 : c[dir] == 0;
        if (!onb) continue;
        bool processor = C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR;
        if (C.dg) {
          if (processor)
            for (int i = 0; i < C.n; i++) flag[M.index(c, cell, i)] = 1;
          continue;
        }
        if (!processor && C.bctype(C.bface(c, dir, side)) != PDB200_BC_DIRICHLET) continue;
        for (int i = 0; i < C.n; i++) {
          int ii = i, lat = 0;
          for (int d = 0; d <= dir; d++) {
            lat = ii % C.n1;
            ii /= C.n1;
          }
          if (lat == (side ? C.k : 0)) flag[M.index(c, cell, i)] = 1;
        }
      }
  }
  return flag;
}

// ---------------------------------------------------------------------------------------------
// DefaultAssembler::assemble, gridoperator/default/assembler.hh:85-279, specialised to the three
// engines.  mode 0: residual (default/residualengine.hh), 1: jacobian_apply
// (default/jacobianapplyengine.hh).
// ---------------------------------------------------------------------------------------------

int assemble_vector(const Ctx& C, const DofMap& M, const double* x, double* r, int mode,
                    long cell_begin, long cell_end) {
  Scratch S(C.n);
  std::vector<double> xl(C.n), xn(C.n), rl(C.n), rn(C.n);
  std::vector<long> idx_s(C.n), idx_n(C.n);
  const bool japply = mode == 1;
  for (long cell = cell_begin; cell < cell_end; cell++) {
    int c[3];
    C.cell_coord(cell, c);
    for (int i = 0; i < C.n; i++) idx_s[i] = M.index(c, cell, i);
    std::fill(rl.begin(), rl.end(), 0.0);  // onBindLFSV
    // assembleVVolume -> lambda_volume (residual engine only; DG only: FEM folds f into alpha)
    if (!japply && C.dg) dg_lambda_volume(C, cell, rl.data(), S);
    for (int i = 0; i < C.n; i++) xl[i] = x[idx_s[i]];  // loadCoefficientsLFSUInside
    // assembleUVVolume
    if (C.dg)
      alpha_volume(C, cell, xl.data(), rl.data(), S, false);
    else
      alpha_volume(C, cell, xl.data(), rl.data(), S, !japply);
    // intersections in YaspGrid order 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++) {
        bool onb = side ? c[dir] == C.N[dir] - 1 : c[dir] == 0;
        if (!onb) {
          if (!C.dg) continue;   // doAlphaSkeleton = false for ConvectionDiffusionFEM
          if (side == 1) continue;  // neighbour has the larger index: visited from there (:181)
          int cn[3] = {c[0], c[1], c[2]};
          cn[dir] -= 1;
          long celln = C.cell_index(cn);
          for (int i = 0; i < C.n; i++) idx_n[i] = M.index(cn, celln, i);
          std::fill(rn.begin(), rn.end(), 0.0);               // onBindLFSVOutside
          for (int i = 0; i < C.n; i++) xn[i] = x[idx_n[i]];  // loadCoefficientsLFSUOutside
          dg_alpha_skeleton(C, cell, celln, dir, xl.data(), xn.data(), rl.data(), rn.data(), S);
          for (int i = 0; i < C.n; i++) r[idx_n[i]] += rn[i];  // onUnbindLFSVOutside
        } else {
          if (C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;  // :239-250 no-op
          if (C.dg) {
            if (dg_boundary(C, cell, c, dir, side, xl.data(), rl.data(), japply, S)) return 1;
          } else {
            fem_alpha_boundary(C, cell, c, dir, side, xl.data(), rl.data(), japply, S);
          }
        }
      }
    for (int i = 0; i < C.n; i++) r[idx_s[i]] += rl[i];  // onUnbindLFSV
  }
  return 0;
}

// postAssembly -> constrain_residual (constraints/common/constraints.hh:904-915); only Dirichlet
// (empty-row) constraints occur on this path, so it reduces to zeroing the constrained entries.
void constrain_residual(const std::vector<char>& flag, double* r) {
  for (size_t i = 0; i < flag.size(); i++)
    if (flag[i]) r[i] = 0.0;
}

int run_vector(const pdb200_problem* p, const double* x, double* r, int mode, int nthreads) {
  try {
    Ctx C(p);
    DofMap M(C);
    int rc = 0;
    if (nthreads <= 1) {
      rc = assemble_vector(C, M, x, r, mode, 0, C.ncells);
    } else {
      // Multi-core baseline: the reference runs one MPI rank per core on an overlapping YaspGrid.
      // Here the cell range is cut into 2*T slabs along the last direction and slabs of one
      // parity are processed concurrently; a cell only writes into itself and lower neighbours,
      // so same-parity slabs never touch the same rows.  Only the summation order at slab
      // interfaces differs from the serial loop.
      int last = C.dim - 1;
      long plane = C.ncells / C.N[last];
      int nslab = std::min(2 * nthreads, C.N[last]);
      for (int parity = 0; parity < 2; parity++) {
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1) reduction(| : rc)
        for (int s = parity; s < nslab; s += 2) {
          long z0 = (long)C.N[last] * s / nslab, z1 = (long)C.N[last] * (s + 1) / nslab;
          rc |= assemble_vector(C, M, x, r, mode, z0 * plane, z1 * plane);
        }
      }
    }
    if (rc) return rc;
    constrain_residual(constrained_flags(C, M), r);
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

// ---------------------------------------------------------------------------------------------
// pattern + jacobian
// ---------------------------------------------------------------------------------------------

struct Csr {
  std::vector<uint64_t> rowptr, colidx;
};

// fill_pattern: gridoperator/default/patternengine.hh:146-204 with FullVolumePattern /
// FullSkeletonPattern (localoperator/pattern.hh:13-47) -> BCRSPattern::add_link
// (backend/istl/bcrspattern.hh:96-119) -> allocate_bcrs_matrix (bcrsmatrixbackend.hh:90-121),
// where dune-istl setIndices stores the columns of each row in ascending order.
Csr build_pattern(const Ctx& C, const DofMap& M) {
  std::vector<std::vector<uint64_t>> rows(M.ndofs);
  std::vector<long> idx_s(C.n), idx_n(C.n);
  auto add_block = [&](const std::vector<long>& ri, const std::vector<long>& ci) {
    for (int i = 0; i < C.n; i++) {
      auto& row = rows[ri[i]];
      for (int j = 0; j < C.n; j++)
        if (std::find(row.begin(), row.end(), (uint64_t)ci[j]) == row.end()) row.push_back(ci[j]);
    }
  };
  for (long cell = 0; cell < C.ncells; cell++) {
    int c[3];
    C.cell_coord(cell, c);
    for (int i = 0; i < C.n; i++) idx_s[i] = M.index(c, cell, i);
    if (C.dg)
      for (int dir = 0; dir < C.dim; dir++) {
        if (c[dir] == 0) continue;
        int cn[3] = {c[0], c[1], c[2]};
        cn[dir] -= 1;
        long celln = C.cell_index(cn);
        for (int i = 0; i < C.n; i++) idx_n[i] = M.index(cn, celln, i);
        add_block(idx_s, idx_n);  // localpattern_sn
        add_block(idx_n, idx_s);  // localpattern_ns
      }
    add_block(idx_s, idx_s);
  }
  Csr P;
  P.rowptr.resize(M.ndofs + 1);
  P.rowptr[0] = 0;
  for (long r = 0; r < M.ndofs; r++) {
    std::sort(rows[r].begin(), rows[r].end());
    P.rowptr[r + 1] = P.rowptr[r] + rows[r].size();
  }
  P.colidx.reserve(P.rowptr[M.ndofs]);
  for (long r = 0; r < M.ndofs; r++) P.colidx.insert(P.colidx.end(), rows[r].begin(), rows[r].end());
  return P;
}

// scatter_jacobian (gridoperator/common/assemblerutilities.hh:449-460) -> UncachedMatrixView::add
// (backend/common/uncachedmatrixview.hh:259-262): A(ri,ci) += v for every entry != 0.0
void scatter_jacobian(const Csr& P, double* values, const LocalMatrix& al, const std::vector<long>& ri,
                      const std::vector<long>& ci) {
  // LocalMatrix iterator runs over the column-major container
  for (int j = 0; j < al.cols; j++)
    for (int i = 0; i < al.rows; i++) {
      double v = al(i, j);
      if (v == 0.0) continue;
      auto b = P.colidx.begin() + P.rowptr[ri[i]], e = P.colidx.begin() + P.rowptr[ri[i] + 1];
      auto it = std::lower_bound(b, e, (uint64_t)ci[j]);
      if (it == e || *it != (uint64_t)ci[j]) throw std::runtime_error("entry not in pattern");
      values[it - P.colidx.begin()] += v;
    }
}

// GridOperator::jacobian -> DefaultLocalJacobianAssemblerEngine (default/jacobianengine.hh)
int assemble_jacobian(const Ctx& C, const DofMap& M, const Csr& P, double* values) {
  Scratch S(C.n);
  LocalMatrix al, al_sn, al_ns, al_nn;
  std::vector<long> idx_s(C.n), idx_n(C.n);
  for (long cell = 0; cell < C.ncells; cell++) {
    int c[3];
    C.cell_coord(cell, c);
    for (int i = 0; i < C.n; i++) idx_s[i] = M.index(c, cell, i);
    al.assign(C.n, C.n);  // onBindLFSUV
    jacobian_volume(C, cell, al, S);
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++) {
        bool onb = side ? c[dir] == C.N[dir] - 1 : c[dir] == 0;
        if (!onb) {
          if (!C.dg || side == 1) continue;
          int cn[3] = {c[0], c[1], c[2]};
          cn[dir] -= 1;
          long celln = C.cell_index(cn);
          for (int i = 0; i < C.n; i++) idx_n[i] = M.index(cn, celln, i);
          al_sn.assign(C.n, C.n);
          al_ns.assign(C.n, C.n);
          al_nn.assign(C.n, C.n);
          dg_jacobian_skeleton(C, cell, celln, dir, al, al_sn, al_ns, al_nn, S);
          scatter_jacobian(P, values, al_sn, idx_s, idx_n);  // onUnbindLFSUVOutside
          scatter_jacobian(P, values, al_ns, idx_n, idx_s);
          scatter_jacobian(P, values, al_nn, idx_n, idx_n);
        } else {
          if (C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;
          if (C.dg) {
            if (dg_jacobian_boundary(C, cell, c, dir, side, al, S)) return 1;
          } else {
            fem_jacobian_boundary(C, cell, c, dir, side, al, S);
          }
        }
      }
    scatter_jacobian(P, values, al, idx_s, idx_s);  // onUnbindLFSUV
  }
  // postAssembly -> handle_dirichlet_constraints -> set_trivial_rows -> clear_row(ri, 1)
  // (assemblerutilities.hh:666-684, backend/istl/bcrsmatrix.hh:254-258)
  std::vector<char> flag = constrained_flags(C, M);
  for (long r = 0; r < M.ndofs; r++)
    if (flag[r])
      for (uint64_t e = P.rowptr[r]; e < P.rowptr[r + 1]; e++) values[e] = P.colidx[e] == (uint64_t)r ? 1.0 : 0.0;
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Sampled rows of pattern + Jacobian: the SAME visits in the SAME order as build_pattern / assemble_jacobian,
// restricted to the cells that contribute to one row, so that a parity test at BASELINE sizes (cfg4: 2.1e9 non-zeros)
// can check 1e5 rows without the 17 GB matrix.  Bit-identical to the rows of oracle_jacobian (checked on small
// grids by tests/test_oracle_rows.py).
// ---------------------------------------------------------------------------------------------

// inverse of DofMap::index for the conforming spaces: lattice coordinates (k N_d + 1 points per direction) of DOF r
void lattice_of_dof(const Ctx& C, const DofMap& M, long r, int lat[3]) {
  lat[0] = lat[1] = lat[2] = 0;
  if (C.k == 1) {
    for (int d = 0; d < C.dim; d++) {
      lat[d] = (int)(r % (C.N[d] + 1));
      r /= C.N[d] + 1;
    }
    return;
  }
  int edim = C.dim;
  while (edim > 0 && r < M.codim_block_off[edim]) edim--;
  r -= M.codim_block_off[edim];
  int sbest = -1;
  for (int s = 0; s < (1 << C.dim); s++)
    if (__builtin_popcount(s) == edim && M.group_off[s] <= r && (sbest < 0 || M.group_off[s] > M.group_off[sbest])) sbest = s;
  r -= M.group_off[sbest];
  for (int d = 0; d < C.dim; d++) {
    int ext = (sbest >> d) & 1;
    long sz = ext ? C.N[d] : C.N[d] + 1;
    lat[d] = 2 * (int)(r % sz) + ext;
    r /= sz;
  }
}

struct RowAcc {
  std::vector<uint64_t> cols;  // ascending
  std::vector<double> vals;
  void add(uint64_t col, double v) {
    if (v == 0.0) return;  // scatter_jacobian drops exact zeros, assemblerutilities.hh:434,456
    auto it = std::lower_bound(cols.begin(), cols.end(), col);
    if (it == cols.end() || *it != col) throw std::runtime_error("entry not in pattern");
    vals[it - cols.begin()] += v;
  }
};

void jacobian_row(const Ctx& C, const DofMap& M, long r, bool constrained, RowAcc& R) {
  Scratch S(C.n);
  LocalMatrix al, al_sn, al_ns, al_nn;
  std::vector<long> idx_s(C.n), idx_n(C.n);
  R.cols.clear();
  if (C.dg) {
    const long e = r / C.n;
    const int i = (int)(r % C.n);
    int c[3];
    C.cell_coord(e, c);
    // pattern: FullVolumePattern + FullSkeletonPattern (both directions) -> own cell and all face neighbours
    for (int j = 0; j < C.n; j++) R.cols.push_back(e * C.n + j);
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++) {
        if (side ? c[dir] == C.N[dir] - 1 : c[dir] == 0) continue;
        int cn[3] = {c[0], c[1], c[2]};
        cn[dir] += side ? 1 : -1;
        long en = C.cell_index(cn);
        for (int j = 0; j < C.n; j++) R.cols.push_back(en * C.n + j);
      }
    std::sort(R.cols.begin(), R.cols.end());
    R.vals.assign(R.cols.size(), 0.0);
    // visit of cell e (assemble_jacobian): volume, lower interior faces (e is the inside cell), boundary faces
    for (int ii = 0; ii < C.n; ii++) idx_s[ii] = e * C.n + ii;
    al.assign(C.n, C.n);
    jacobian_volume(C, e, al, S);
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++) {
        bool onb = side ? c[dir] == C.N[dir] - 1 : c[dir] == 0;
        if (!onb) {
          if (side == 1) continue;
          int cn[3] = {c[0], c[1], c[2]};
          cn[dir] -= 1;
          long celln = C.cell_index(cn);
          al_sn.assign(C.n, C.n);
          al_ns.assign(C.n, C.n);
          al_nn.assign(C.n, C.n);
          dg_jacobian_skeleton(C, e, celln, dir, al, al_sn, al_ns, al_nn, S);
          for (int j = 0; j < C.n; j++) R.add(celln * C.n + j, al_sn(i, j));
        } else {
          if (C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;
          if (dg_jacobian_boundary(C, e, c, dir, side, al, S)) throw std::runtime_error(g_err);
        }
      }
    for (int j = 0; j < C.n; j++) R.add(e * C.n + j, al(i, j));
    // visits of the upper neighbours u = e + e_dir (ascending): e is their outside cell
    for (int dir = 0; dir < C.dim; dir++) {
      if (c[dir] == C.N[dir] - 1) continue;
      int cu[3] = {c[0], c[1], c[2]};
      cu[dir] += 1;
      long u = C.cell_index(cu);
      LocalMatrix scratch_ss;
      scratch_ss.assign(C.n, C.n);
      al_sn.assign(C.n, C.n);
      al_ns.assign(C.n, C.n);
      al_nn.assign(C.n, C.n);
      dg_jacobian_skeleton(C, u, e, dir, scratch_ss, al_sn, al_ns, al_nn, S);
      for (int j = 0; j < C.n; j++) R.add(u * C.n + j, al_ns(i, j));
      for (int j = 0; j < C.n; j++) R.add(e * C.n + j, al_nn(i, j));
    }
  } else {
    int lat[3];
    lattice_of_dof(C, M, r, lat);
    // cells whose closure holds the lattice point, ascending cell index
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < C.dim; d++) {
      if (lat[d] % C.k != 0) {
        lo[d] = hi[d] = lat[d] / C.k;
      } else {
        lo[d] = std::max(0, lat[d] / C.k - 1);
        hi[d] = std::min(C.N[d] - 1, lat[d] / C.k);
      }
    }
    std::vector<long> cells;
    for (int z = lo[2]; z <= hi[2]; z++)
      for (int y = lo[1]; y <= hi[1]; y++)
        for (int x = lo[0]; x <= hi[0]; x++) {
          int cc[3] = {x, y, z};
          cells.push_back(C.cell_index(cc));
        }
    for (long cell : cells) {
      int c[3];
      C.cell_coord(cell, c);
      for (int j = 0; j < C.n; j++) R.cols.push_back(M.index(c, cell, j));
    }
    std::sort(R.cols.begin(), R.cols.end());
    R.cols.erase(std::unique(R.cols.begin(), R.cols.end()), R.cols.end());
    R.vals.assign(R.cols.size(), 0.0);
    for (long cell : cells) {
      int c[3];
      C.cell_coord(cell, c);
      int i = -1;
      for (int j = 0; j < C.n; j++) {
        idx_s[j] = M.index(c, cell, j);
        if (idx_s[j] == r) i = j;
      }
      if (i < 0) throw std::runtime_error("row not found in an adjacent cell");
      al.assign(C.n, C.n);
      jacobian_volume(C, cell, al, S, i);
      for (int dir = 0; dir < C.dim; dir++)
        for (int side = 0; side < 2; side++) {
          bool onb = side ? c[dir] == C.N[dir] - 1 : c[dir] == 0;
          if (!onb || C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;
          fem_jacobian_boundary(C, cell, c, dir, side, al, S);
        }
      for (int j = 0; j < C.n; j++) R.add(idx_s[j], al(i, j));
    }
  }
  if (constrained)  // set_trivial_rows, assemblerutilities.hh:666-684
    for (size_t e = 0; e < R.cols.size(); e++) R.vals[e] = R.cols[e] == (uint64_t)r ? 1.0 : 0.0;
}

// is DOF r constrained?  (constrained_flags for one DOF, without the O(N) flag vector)
bool dof_constrained(const Ctx& C, const DofMap& M, long r) {
  if (C.dg) {
    int c[3];
    C.cell_coord(r / C.n, c);
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++)
        if ((side ? c[dir] == C.N[dir] - 1 : c[dir] == 0) && C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) return true;
    return false;
  }
  int lat[3];
  lattice_of_dof(C, M, r, lat);
  // a DOF is constrained iff it lies in the closure of a Dirichlet (or processor) boundary face
  for (int dir = 0; dir < C.dim; dir++)
    for (int side = 0; side < 2; side++) {
      if (lat[dir] != (side ? C.k * C.N[dir] : 0)) continue;
      if (C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) return true;
      // boundary faces (cells of the layer) whose closure holds the point
      int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      for (int d = 0; d < C.dim; d++) {
        if (d == dir) {
          lo[d] = hi[d] = side ? C.N[d] - 1 : 0;
        } else if (lat[d] % C.k != 0) {
          lo[d] = hi[d] = lat[d] / C.k;
        } else {
          lo[d] = std::max(0, lat[d] / C.k - 1);
          hi[d] = std::min(C.N[d] - 1, lat[d] / C.k);
        }
      }
      for (int z = lo[2]; z <= hi[2]; z++)
        for (int y = lo[1]; y <= hi[1]; y++)
          for (int x = lo[0]; x <= hi[0]; x++) {
            int cc[3] = {x, y, z};
            if (C.bctype(C.bface(cc, dir, side)) == PDB200_BC_DIRICHLET) return true;
          }
    }
  return false;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

#define ORACLE_TRY try {
#define ORACLE_CATCH                \
  }                                 \
  catch (std::exception & e) {      \
    g_err = e.what();               \
    return 2;                       \
  }                                 \
  return 0;

int oracle_num_dofs(const pdb200_problem* p, uint64_t* n) {
  ORACLE_TRY
  Ctx C(p);
  DofMap M(C);
  *n = M.ndofs;
  ORACLE_CATCH
}

int oracle_sizes(const pdb200_problem* p, uint32_t* local_size, uint32_t* m, uint64_t* nbfaces) {
  ORACLE_TRY
  Ctx C(p);
  *local_size = C.n;
  *m = C.T.m;
  *nbfaces = C.nbf;
  ORACLE_CATCH
}

int oracle_boundary_face_offset(const pdb200_problem* p, int dir, int side, uint64_t* first) {
  ORACLE_TRY
  Ctx C(p);
  *first = C.bf_off[dir][side];
  ORACLE_CATCH
}

int oracle_quadrature(const pdb200_problem* p, double* points, double* weights) {
  ORACLE_TRY
  Ctx C(p);
  for (int i = 0; i < C.T.m; i++) {
    points[i] = C.T.xq[i];
    weights[i] = C.T.wq[i];
  }
  ORACLE_CATCH
}

int oracle_cell_dof_indices(const pdb200_problem* p, uint64_t cell, uint64_t* idx) {
  ORACLE_TRY
  Ctx C(p);
  DofMap M(C);
  int c[3];
  C.cell_coord((long)cell, c);
  for (int i = 0; i < C.n; i++) idx[i] = M.index(c, (long)cell, i);
  ORACLE_CATCH
}

int oracle_constrained_dofs(const pdb200_problem* p, uint64_t* count, uint64_t* idx) {
  ORACLE_TRY
  Ctx C(p);
  DofMap M(C);
  std::vector<char> flag = constrained_flags(C, M);
  uint64_t k = 0;
  for (size_t i = 0; i < flag.size(); i++)
    if (flag[i]) {
      if (idx) idx[k] = i;
      k++;
    }
  *count = k;
  ORACLE_CATCH
}

// GridOperator::residual, gridoperator/gridoperator.hh:176-181:  r += R(x)
int oracle_residual(const pdb200_problem* p, const double* x, double* r) { return run_vector(p, x, r, 0, 1); }

// GridOperator::jacobian_apply (linear), gridoperator/gridoperator.hh:192-197:  y += J z.
// For ConvectionDiffusionFEM the reference uses the finite-difference mixins
// (convectiondiffusionfem.hh:40-41); this returns the exact derivative (documented deviation,
// SURVEY.md §8a row 6), oracle_fem_jacobian_apply_fd below restates the mixin.
int oracle_jacobian_apply(const pdb200_problem* p, const double* z, double* y) { return run_vector(p, z, y, 1, 1); }

// multi-threaded variants used only as the CPU baseline in bench.py
int oracle_jacobian_apply_mt(const pdb200_problem* p, const double* z, double* y, int nthreads) {
  return run_vector(p, z, y, 1, nthreads);
}
int oracle_residual_mt(const pdb200_problem* p, const double* x, double* r, int nthreads) {
  return run_vector(p, x, r, 0, nthreads);
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// NumericalJacobianApplyVolume / NumericalJacobianApplyBoundary (linear variants),
// localoperator/numericaljacobianapply.hh:54-85 and :307-341, epsilon = 1e-7, as inherited by
// ConvectionDiffusionFEM (convectiondiffusionfem.hh:40-41).
int oracle_fem_jacobian_apply_fd(const pdb200_problem* p, const double* x, double* y) {
  ORACLE_TRY
  Ctx C(p);
  if (C.dg) throw std::runtime_error("finite-difference jacobian_apply is the FEM code path");
  DofMap M(C);
  Scratch S(C.n);
  const double epsilon = 1e-7;
  std::vector<double> xl(C.n), u(C.n), down(C.n), up(C.n), yl(C.n);
  std::vector<long> idx(C.n);
  for (long cell = 0; cell < C.ncells; cell++) {
    int c[3];
    C.cell_coord(cell, c);
    for (int i = 0; i < C.n; i++) idx[i] = M.index(c, cell, i);
    for (int i = 0; i < C.n; i++) xl[i] = x[idx[i]];
    std::fill(yl.begin(), yl.end(), 0.0);
    auto fd = [&](auto&& alpha) {
      u = xl;
      std::fill(down.begin(), down.end(), 0.0);
      alpha(u.data(), down.data());
      for (int j = 0; j < C.n; j++) {
        std::fill(up.begin(), up.end(), 0.0);
        double delta = epsilon * (1.0 + std::abs(u[j]));
        u[j] += delta;
        alpha(u.data(), up.data());
        for (int i = 0; i < C.n; i++) yl[i] += ((up[i] - down[i]) / delta) * xl[j];
        u[j] = xl[j];
      }
    };
    fd([&](const double* uu, double* rr) { alpha_volume(C, cell, uu, rr, S, true); });
    for (int dir = 0; dir < C.dim; dir++)
      for (int side = 0; side < 2; side++) {
        bool onb = side ? c[dir] == C.N[dir] - 1 : c[dir] == 0;
        if (!onb || C.p->side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;
        fd([&](const double* uu, double* rr) { fem_alpha_boundary(C, cell, c, dir, side, uu, rr, false, S); });
      }
    for (int i = 0; i < C.n; i++) y[idx[i]] += yl[i];
  }
  constrain_residual(constrained_flags(C, M), y);
  ORACLE_CATCH
}

int oracle_pattern(const pdb200_problem* p, uint64_t* nrows, uint64_t* nnz, uint64_t* rowptr, uint64_t* colidx) {
  ORACLE_TRY
  Ctx C(p);
  DofMap M(C);
  Csr P = build_pattern(C, M);
  *nrows = M.ndofs;
  *nnz = P.colidx.size();
  if (rowptr) std::copy(P.rowptr.begin(), P.rowptr.end(), rowptr);
  if (colidx) std::copy(P.colidx.begin(), P.colidx.end(), colidx);
  ORACLE_CATCH
}

// GridOperator::jacobian, gridoperator/gridoperator.hh:184-189: values (CSR order of
// oracle_pattern) += dR/dx, then constrained rows := unit rows
int oracle_jacobian(const pdb200_problem* p, const double* x, double* values) {
  (void)x;  // both local operators are linear: the matrix does not depend on x
  ORACLE_TRY
  Ctx C(p);
  DofMap M(C);
  Csr P = build_pattern(C, M);
  if (assemble_jacobian(C, M, P, values)) return 1;
  ORACLE_CATCH
}

// rows of fill_pattern + jacobian for the listed DOFs: rowlen[s] entries per sampled row s, padded to maxlen in
// colidx / values (row s at offset s * maxlen); maxlen >= (2k+1)^dim (conforming) or (2 dim + 1) n (QkDG)
int oracle_jacobian_rows(const pdb200_problem* p, const uint64_t* rows, uint64_t nrows, int nthreads, uint64_t maxlen,
                         uint64_t* rowlen, uint64_t* colidx, double* values) {
  ORACLE_TRY
  Ctx C(p);
  DofMap M(C);
  std::string err;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 64)
  for (long s = 0; s < (long)nrows; s++) {
    try {
      RowAcc R;
      long r = (long)rows[s];
      if (r < 0 || r >= M.ndofs) throw std::runtime_error("row out of range");
      jacobian_row(C, M, r, dof_constrained(C, M, r), R);
      if (R.cols.size() > maxlen) throw std::runtime_error("maxlen too small");
      rowlen[s] = R.cols.size();
      for (size_t e = 0; e < R.cols.size(); e++) {
        colidx[s * maxlen + e] = R.cols[e];
        values[s * maxlen + e] = R.vals[e];
      }
    } catch (std::exception& ex) {
#pragma omp critical
      err = ex.what();
    }
  }
  if (!err.empty()) throw std::runtime_error(err);
  ORACLE_CATCH
}

}  // extern "C"

// test_gridoperator.cc — the C++ host mirror (dune-pdelab_b200/host/gridoperator.hh) exercised the way
// the reference's own tests use Dune::PDELab::GridOperator.  Needs a CUDA device (pytest -m gpu
// builds and runs it, tests/test_gpu_cpp_host.py).  Restated reference tests:
//   test/testconvectiondiffusiondg.cc:46-168   DG k=1, 2D 16^2, u = exp(-|x-1/2|^2), err^2 <= 1e-6
//   test/matrixfree/matrix_free_linear.cc      assembled vs matrix-free: same Krylov iteration count
//   test/testmatrixfree.cc:150-178             Q2 conforming FEM 2D 32^2, err^2 <= 1e-7, both ways
//   test/test-blocked-istl-ordering.cc:45-72   flat and blocked DG vectors are the same bytes
//   gridoperator/gridoperator.hh:200-205       jacobian_apply(u,z,y) throws for a linear operator
//   test/testinstationaryfastdgassembler.cc    DG k=1 2D 8^2 + L2, Alexander2, one step dt=0.1, err^2 <= 5e-6
// and every GPU result is compared with the CPU oracle (liboracle.so, test infrastructure) fed
// with the very same pdb200_problem.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <vector>

#include "../../dune-pdelab_b200/host/gridoperator.hh"
#include "../../dune-pdelab_b200/host/onestep.hh"

extern "C" {  // oracle/pdelab_oracle.cc
int oracle_residual(const pdb200_problem*, const double*, double*);
int oracle_jacobian_apply(const pdb200_problem*, const double*, double*);
int oracle_jacobian(const pdb200_problem*, const double*, double*);
int oracle_pattern(const pdb200_problem*, uint64_t*, uint64_t*, uint64_t*, uint64_t*);
const char* oracle_last_error(void);
}

namespace PDELab = Dune::PDELab::B200;

static int failures = 0;
#define EXPECT(cond, msg)                                                     \
  do {                                                                        \
    if (!(cond)) {                                                            \
      std::cerr << "FAIL " << __LINE__ << ": " << msg << std::endl;           \
      failures++;                                                             \
    } else {                                                                  \
      std::cout << "ok   " << msg << std::endl;                               \
    }                                                                         \
  } while (0)

// the parameter class of test/testconvectiondiffusiondg.cc:10-44, same call-back interface
template <typename GridView, typename RangeType>
class PoissonProblem : public PDELab::ConvectionDiffusionModelProblem<GridView, RangeType> {
 public:
  template <typename Element, typename Coord>
  auto f(const Element& element, const Coord& x) const {
    auto global = element.geometry().global(x);
    double c = 0;
    for (std::size_t i = 0; i < global.size(); i++) c += (0.5 - global[i]) * (0.5 - global[i]);
    const double dim = (double)global.size();
    return (2.0 * dim - 4.0 * c) * std::exp(-c);  // -laplace exp(-c) in any dimension (4(1-c)g in 2D)
  }
  template <typename Element, typename Coord>
  auto bctype(const Element&, const Coord&) const {
    return PDELab::ConvectionDiffusionBoundaryConditions::Dirichlet;
  }
  template <typename Element, typename Coord>
  RangeType g(const Element& element, const Coord& x) const {
    auto global = element.geometry().global(x);
    double c = 0;
    for (std::size_t i = 0; i < global.size(); i++) c += (0.5 - global[i]) * (0.5 - global[i]);
    return std::exp(-c);
  }
};

// heterogeneous diagonal permeability + reaction: exercises the sampled coefficient arrays
template <typename GridView, typename RangeType>
class LayeredProblem : public PoissonProblem<GridView, RangeType> {
 public:
  using Traits = PDELab::ConvectionDiffusionParameterTraits<GridView, RangeType>;
  template <typename Element, typename Coord>
  typename Traits::PermTensorType A(const Element& e, const Coord& x) const {
    auto g = e.geometry().global(x);
    typename Traits::PermTensorType K(0.0);
    for (int i = 0; i < Traits::dimDomain; i++) K[i][i] = (g[Traits::dimDomain - 1] > 0.5 ? 10.0 : 0.1) * (1 + i);
    return K;
  }
  template <typename Element, typename Coord>
  RangeType c(const Element&, const Coord&) const {
    return 2.0;
  }
};

// spatially varying fields: a rotating velocity, a reaction term that varies inside the cells, a diffusion tensor that
// is NOT constant per cell and a boundary type that changes inside boundary faces — every call-back must be sampled
// where the reference evaluates it (convectiondiffusiondg.hh:143-146,178,181,367-382,426,752-763)
template <typename GridView, typename RangeType>
class RotatingProblem : public PoissonProblem<GridView, RangeType> {
 public:
  using Traits = PDELab::ConvectionDiffusionParameterTraits<GridView, RangeType>;
  static constexpr bool permeabilityIsConstantPerCell() { return false; }
  template <typename Element, typename Coord>
  typename Traits::PermTensorType A(const Element& e, const Coord& x) const {
    auto g = e.geometry().global(x);
    typename Traits::PermTensorType K(0.0);
    for (int i = 0; i < Traits::dimDomain; i++) K[i][i] = (1.0 + 0.5 * std::sin(3.0 * g[0] + 2.0 * g[1])) * (1 + 0.3 * i);
    return K;
  }
  template <typename Element, typename Coord>
  typename Traits::RangeType b(const Element& e, const Coord& x) const {
    auto g = e.geometry().global(x);
    typename Traits::RangeType v(0.0);
    v[0] = -(g[1] - 0.5);
    v[1] = g[0] - 0.5;
    return v;
  }
  template <typename Element, typename Coord>
  RangeType c(const Element& e, const Coord& x) const {
    auto g = e.geometry().global(x);
    return 1.0 + g[0] * g[1];
  }
  template <typename Intersection, typename Coord>
  auto bctype(const Intersection& is, const Coord& x) const {
    auto g = is.geometry().global(x);
    double s = 0;
    for (std::size_t i = 0; i < g.size(); i++) s += (7.0 - 2.0 * i) * g[i];
    return std::sin(s) > 0.2 ? PDELab::ConvectionDiffusionBoundaryConditions::Dirichlet
                             : PDELab::ConvectionDiffusionBoundaryConditions::Neumann;
  }
  template <typename Intersection, typename Coord>
  RangeType j(const Intersection& is, const Coord& x) const {
    return 0.25 + is.geometry().global(x)[0];
  }
};

template <class V>
double rel_err(const V& a, const std::vector<double>& b) {
  double e = 0, n = 0;
  for (std::size_t i = 0; i < b.size(); i++) {
    e = std::max(e, std::abs(a[i] - b[i]));
    n = std::max(n, std::abs(b[i]));
  }
  return e / (n > 0 ? n : 1.0);
}

// unpreconditioned CG on A x = b; returns the iteration count
template <class Apply, class V>
int cg(Apply&& A, const V& b, V& x, double reduction, int maxit) {
  V r(b), p(b), q(b);
  A(x, q);
  r = b;
  r -= q;
  p = r;
  double rr = r.dot(r);
  const double stop = reduction * reduction * rr;
  int it = 0;
  while (it < maxit && rr > stop) {
    A(p, q);
    const double alpha = rr / p.dot(q);
    x.axpy(alpha, p);
    r.axpy(-alpha, q);
    const double rr2 = r.dot(r);
    p *= rr2 / rr;
    p += r;
    rr = rr2;
    it++;
  }
  return it;
}

// int (u_h - g)^2 with a 6-point Gauss rule per direction (integrateGridFunction(..., 10))
template <class GFS, class V, class Problem>
double l2_error_squared(const GFS& gfs, const V& u, const Problem& problem, pdb200_handle h, int k) {
  constexpr int dim = GFS::GridView::dimension;
  const auto& gv = gfs.gridView();
  const int m = 6, n1 = k + 1;
  std::vector<double> xq(m), wq(m);
  pdb200_gauss_legendre(m, xq.data(), wq.data());
  int n = 1, nq = 1;
  for (int d = 0; d < dim; d++) n *= n1, nq *= m;
  std::vector<uint64_t> idx(n);
  double err = 0;
  auto lag = [&](int i, double x) {
    double v = 1;
    for (int j = 0; j < n1; j++)
      if (j != i) v *= (k * x - j) / double(i - j);
    return v;
  };
  for (long long e = 0; e < gv.size(0); e++) {
    const auto cell = gv.cell(e);
    pdb200_cell_dof_indices(h, (uint64_t)e, idx.data());
    for (int q = 0; q < nq; q++) {
      PDELab::FieldVector<double, dim> x;
      double w = cell.geometry().volume();
      int r = q;
      for (int d = 0; d < dim; d++) {
        x[d] = xq[r % m];
        w *= wq[r % m];
        r /= m;
      }
      double uh = 0;
      for (int i = 0; i < n; i++) {
        double phi = 1;
        int ii = i;
        for (int d = 0; d < dim; d++) {
          phi *= lag(ii % n1, x[d]);
          ii /= n1;
        }
        uh += u[idx[i]] * phi;
      }
      const double diff = uh - problem.g(cell, x);
      err += w * diff * diff;
    }
  }
  return err;
}

// ---- test/testconvectiondiffusiondg.cc + test/matrixfree/matrix_free_linear.cc -------------------
template <int dim, int degree, template <class, class> class ProblemT>
void dg_case(int ncells, double alpha, double tol, const char* name, bool check_error) {
  using Grid = PDELab::YaspGrid<dim>;
  std::array<int, dim> cells;
  cells.fill(ncells);
  Grid grid(PDELab::FieldVector<double, dim>(1.0), cells);
  using GridView = typename Grid::LeafGridView;
  GridView gridView = grid.leafGridView();
  using DomainField = double;
  using RangeType = double;
  using FiniteElementMap = PDELab::QkDGLocalFiniteElementMap<DomainField, RangeType, degree, dim>;
  FiniteElementMap finiteElementMap;
  using Constraints = PDELab::NoConstraints;
  using VectorBackend = PDELab::ISTL::VectorBackend<PDELab::ISTL::Blocking::fixed, FiniteElementMap::maxLocalSize()>;
  using GridFunctionSpace = PDELab::GridFunctionSpace<GridView, FiniteElementMap, Constraints, VectorBackend>;
  GridFunctionSpace gridFunctionSpace(gridView, finiteElementMap);
  gridFunctionSpace.name("numerical_solution");
  using Problem = ProblemT<GridView, RangeType>;
  Problem problem;
  using LocalOperator = PDELab::ConvectionDiffusionDG<Problem, FiniteElementMap>;
  LocalOperator localOperator(problem, PDELab::ConvectionDiffusionDGMethod::SIPG,
                              PDELab::ConvectionDiffusionDGWeights::weightsOn, alpha);
  using MatrixBackend = PDELab::ISTL::BCRSMatrixBackend;
  MatrixBackend matrixBackend(std::pow(2, dim) * gridFunctionSpace.maxLocalSize());
  using GridOperator =
      PDELab::GridOperator<GridFunctionSpace, GridFunctionSpace, LocalOperator, MatrixBackend, DomainField, RangeType, RangeType>;
  GridOperator gridOperator(gridFunctionSpace, gridFunctionSpace, localOperator, matrixBackend);
  using V = typename GridOperator::Domain;
  const std::size_t N = gridFunctionSpace.size();
  EXPECT(gridOperator.globalSizeU() == N, name << ": globalSizeU == gfs.size() == " << N);
  EXPECT(&gridOperator.assembler().trialGridFunctionSpace() == &gridFunctionSpace &&
             gridOperator.localAssembler().doPreProcessing() && gridOperator.localAssembler().doPostProcessing(),
         name << ": assembler() and the pre-/post-processing flags (gridoperator.hh:115-151)");
  if (std::is_same<Problem, RotatingProblem<GridView, RangeType>>::value) {
    // the sampled arrays hold the call-backs' values at the reference's evaluation points (layout (2) of pdelab_b200.h)
    const pdb200_problem& P = gridOperator.problem();
    const int all = PDB200_POINTWISE_A | PDB200_POINTWISE_B | PDB200_POINTWISE_C | PDB200_POINTWISE_BCTYPE;
    EXPECT(P.pointwise == all && P.a_mode == PDB200_A_DIAGONAL, name << ": A, b, c, bctype handed over per quadrature point");
    const int m = degree + 1;
    std::vector<double> xq(m), wq(m);
    pdb200_gauss_legendre(m, xq.data(), wq.data());
    int nq = 1, nfq = 1;
    for (int d = 0; d < dim; d++) nq *= m;
    for (int d = 1; d < dim; d++) nfq *= m;
    const int NP = nq + 2 * dim * nfq;
    const double h = 1.0 / ncells;
    double worst = 0;
    for (long long e = 0; e < gridView.size(0); e++) {
      int c[3] = {0, 0, 0};
      long long r = e;
      for (int d = 0; d < dim; d++) {
        c[d] = (int)(r % ncells);
        r /= ncells;
      }
      for (int pt = 0; pt < NP; pt++) {
        double X[3] = {0, 0, 0};
        if (pt < nq) {
          int q = pt;
          for (int d = 0; d < dim; d++) {
            X[d] = (c[d] + xq[q % m]) * h;
            q /= m;
          }
        } else {
          const int f = (pt - nq) / nfq, dir = f / 2, side = f % 2;
          int q = (pt - nq) % nfq;
          for (int d = 0; d < dim; d++) {
            if (d == dir) {
              X[d] = (c[d] + side) * h;
            } else {
              X[d] = (c[d] + xq[q % m]) * h;
              q /= m;
            }
          }
        }
        worst = std::max(worst, std::abs(P.b[(e * NP + pt) * dim + 0] + (X[1] - 0.5)));
        worst = std::max(worst, std::abs(P.b[(e * NP + pt) * dim + 1] - (X[0] - 0.5)));
        for (int i = 0; i < dim; i++)
          worst = std::max(worst, std::abs(P.A[(e * NP + pt) * dim + i] - (1.0 + 0.5 * std::sin(3.0 * X[0] + 2.0 * X[1])) * (1 + 0.3 * i)));
        if (pt < nq) worst = std::max(worst, std::abs(P.c[e * nq + pt] - (1.0 + X[0] * X[1])));
      }
    }
    EXPECT(worst < 1e-14, name << ": sampled A, b, c equal the call-backs at the quadrature points (max deviation " << worst << ")");
  }

  // --- oracle parity of residual / jacobian_apply / jacobian on random data ---------------------
  std::mt19937_64 rng;  // default seed 5489, test/test-blocked-istl-ordering.cc:45-48
  std::uniform_real_distribution<double> dist(0, 1);
  V x(gridFunctionSpace), r(gridFunctionSpace, 0.0), y(gridFunctionSpace, 0.0);
  for (std::size_t i = 0; i < N; i++) x[i] = dist(rng);
  gridOperator.residual(x, r);
  std::vector<double> want(N, 0.0);
  if (oracle_residual(&gridOperator.problem(), x.data(), want.data())) std::cerr << oracle_last_error() << std::endl;
  EXPECT(rel_err(r, want) < 1e-12, name << ": residual vs oracle " << rel_err(r, want));
  gridOperator.jacobian_apply(x, y);
  std::fill(want.begin(), want.end(), 0.0);
  oracle_jacobian_apply(&gridOperator.problem(), x.data(), want.data());
  EXPECT(rel_err(y, want) < 1e-12, name << ": jacobian_apply vs oracle " << rel_err(y, want) << " [" << gridOperator.lastKernel() << "]");
  // accumulate semantics: a second call adds again (jacobianapplyengine.hh:197-202)
  gridOperator.jacobian_apply(x, y);
  for (auto& v : want) v *= 2;
  EXPECT(rel_err(y, want) < 1e-12, name << ": jacobian_apply accumulates into y");

  typename GridOperator::Jacobian A(gridOperator);
  A = 0.0;
  gridOperator.jacobian(x, A);
  uint64_t nr = 0, nnz = 0;
  oracle_pattern(&gridOperator.problem(), &nr, &nnz, nullptr, nullptr);
  std::vector<uint64_t> orp(nr + 1), oci(nnz);
  oracle_pattern(&gridOperator.problem(), &nr, &nnz, orp.data(), oci.data());
  EXPECT(orp == A.rowptr() && oci == A.colidx(), name << ": sparsity pattern bit-exact vs oracle (" << nnz << " nnz)");
  std::vector<double> ov(nnz, 0.0);
  oracle_jacobian(&gridOperator.problem(), x.data(), ov.data());
  EXPECT(rel_err(A.values(), ov) < 1e-12, name << ": jacobian values vs oracle " << rel_err(A.values(), ov));
  // J z == jacobian_apply(z)
  V Jx(gridFunctionSpace, 0.0), y1(gridFunctionSpace, 0.0);
  A.mv(x, Jx);
  PDELab::OnTheFlyOperator<V, V, GridOperator> otf(gridOperator);
  otf.apply(x, y1);
  EXPECT(rel_err(y1, Jx.native()) < 1e-12, name << ": OnTheFlyOperator::apply == J x " << rel_err(y1, Jx.native()));
  // the non-linear overload throws for a linear local operator (gridoperator.hh:200-205)
  bool threw = false;
  try {
    gridOperator.jacobian_apply(x, x, y);
  } catch (PDELab::Exception& e) {
    threw = true;
  }
  EXPECT(threw, name << ": jacobian_apply(u,z,y) throws for a linear operator");

  if (!check_error) return;
  // --- solve matrix-based and matrix-free (StationaryLinearProblemSolver::apply, linearproblem.hh:203-246)
  V u(gridFunctionSpace, 0.0), res(gridFunctionSpace, 0.0), z(gridFunctionSpace, 0.0), rhs(gridFunctionSpace, 0.0);
  gridOperator.residual(u, res);  // r = R(u0)
  rhs = res;
  rhs *= -1.0;
  int it_mat = cg([&](const V& a, V& b) { A.mv(a, b); }, rhs, z, 1e-12, 5000);
  V u_mat(u);
  u_mat += z;
  z = 0.0;
  int it_free = cg([&](const V& a, V& b) { otf.apply(a, b); }, rhs, z, 1e-12, 5000);
  V u_free(u);
  u_free += z;
  const double e_mat = l2_error_squared(gridFunctionSpace, u_mat, problem, gridOperator.handle(), degree);
  const double e_free = l2_error_squared(gridFunctionSpace, u_free, problem, gridOperator.handle(), degree);
  EXPECT(e_mat <= tol && !std::isnan(e_mat), name << ": l2errorsquared matrix based " << e_mat << " <= " << tol);
  EXPECT(e_free <= tol && !std::isnan(e_free), name << ": l2errorsquared matrix free " << e_free << " <= " << tol);
  EXPECT(it_mat == it_free, name << ": same CG iteration count assembled/matrix-free (" << it_mat << ", " << it_free << ")");

  // --- test/matrixfree/matrix_free_linear.cc:350-393: StationaryLinearProblemSolver with an assembled and
  // a matrix-free BiCGSTAB back-end, both running entirely on the device
  {
    using LSM = PDELab::ISTLBackend_SEQ_BCGS_Richardson<GridOperator>;
    using LSF = PDELab::ISTLBackend_SEQ_MatrixFree_BCGS_Richardson<GridOperator>;
    LSM linearSolver(gridOperator, 5000, 0);
    LSF linearSolverMatrixFree(gridOperator, 5000, 0);
    V c1(gridFunctionSpace, 0.0), c2(gridFunctionSpace, 0.0);
    PDELab::StationaryLinearProblemSolver<GridOperator, LSM, V> solver(gridOperator, linearSolver, c1, 1e-12);
    solver.apply();
    PDELab::StationaryLinearProblemSolver<GridOperator, LSF, V> solverMatrixFree(gridOperator, linearSolverMatrixFree, c2, 1e-12);
    solverMatrixFree.apply();
    const auto r1 = solver.result(), r2 = solverMatrixFree.result();
    const double e1 = l2_error_squared(gridFunctionSpace, c1, problem, gridOperator.handle(), degree);
    const double e2 = l2_error_squared(gridFunctionSpace, c2, problem, gridOperator.handle(), degree);
    EXPECT(r1.converged && e1 <= tol, name << ": StationaryLinearProblemSolver (assembled BiCGSTAB on device) err^2 " << e1
                                            << ", " << r1.linear_solver_iterations << " iterations");
    EXPECT(r2.converged && e2 <= tol, name << ": StationaryLinearProblemSolver (matrix-free BiCGSTAB on device) err^2 " << e2
                                            << ", " << r2.linear_solver_iterations << " iterations");
    // BiCGSTAB amplifies the different summation orders of the two operators (row gather vs Kronecker
    // kernel): near the 1e-12 floor the stopping test moves by a few per cent of the iterations
    EXPECT(std::abs(r1.linear_solver_iterations - r2.linear_solver_iterations) <=
               std::max(5, (r1.linear_solver_iterations + r2.linear_solver_iterations) / 20),
           name << ": comparable iteration counts (" << r1.linear_solver_iterations << ", " << r2.linear_solver_iterations << ")");
    EXPECT(r2.defect <= 1e-12 * r2.first_defect * 1.000001, name << ": defect reduced by 1e-12");
    // the matrix-free back-end of matrix_free_linear.cc:338-341 proper: BiCGSTAB preconditioned with block Jacobi
    using LSB = PDELab::ISTLBackend_SEQ_MatrixFree_BCGS_BlockJacobi<GridOperator>;
    LSB linearSolverBlockJacobi(gridOperator, 5000, 0);
    V c3(gridFunctionSpace, 0.0);
    PDELab::StationaryLinearProblemSolver<GridOperator, LSB, V> solverBJ(gridOperator, linearSolverBlockJacobi, c3, 1e-12);
    solverBJ.apply();
    const auto r3 = solverBJ.result();
    const double e3 = l2_error_squared(gridFunctionSpace, c3, problem, gridOperator.handle(), degree);
    EXPECT(r3.converged && e3 <= tol && r3.linear_solver_iterations < r2.linear_solver_iterations,
           name << ": matrix-free BiCGSTAB + block Jacobi err^2 " << e3 << ", " << r3.linear_solver_iterations
                << " iterations (unpreconditioned: " << r2.linear_solver_iterations << ")");
    // and with the matrix-free block SOR preconditioner (blocksorpreconditioner.hh)
    using LSS = PDELab::ISTLBackend_SEQ_MatrixFree_BCGS_BlockSOR<GridOperator>;
    LSS linearSolverBlockSOR(gridOperator, 5000, 0);
    V c4(gridFunctionSpace, 0.0);
    PDELab::StationaryLinearProblemSolver<GridOperator, LSS, V> solverSOR(gridOperator, linearSolverBlockSOR, c4, 1e-12);
    solverSOR.apply();
    const auto r4 = solverSOR.result();
    const double e4 = l2_error_squared(gridFunctionSpace, c4, problem, gridOperator.handle(), degree);
    EXPECT(r4.converged && e4 <= tol && r4.linear_solver_iterations < r2.linear_solver_iterations,
           name << ": matrix-free BiCGSTAB + block SOR err^2 " << e4 << ", " << r4.linear_solver_iterations
                << " iterations (block Jacobi: " << r3.linear_solver_iterations << ")");
  }
}

// ---- test/testmatrixfree.cc: conforming Q2, Dirichlet constraints + interpolate --------------------
template <int dim, int degree>
void fem_case(int ncells, double tol, const char* name) {
  using Grid = PDELab::YaspGrid<dim>;
  std::array<int, dim> cells;
  cells.fill(ncells);
  Grid grid(PDELab::FieldVector<double, dim>(1.0), cells);
  using GV = typename Grid::LeafGridView;
  GV gv = grid.leafGridView();
  using FEM = PDELab::QkLocalFiniteElementMap<GV, double, double, degree>;
  FEM fem(gv);
  using GFS = PDELab::GridFunctionSpace<GV, FEM, PDELab::ConformingDirichletConstraints, PDELab::ISTL::VectorBackend<>>;
  GFS gfs(gv, fem);
  using Problem = PoissonProblem<GV, double>;
  Problem problem;
  using LOP = PDELab::ConvectionDiffusionFEM<Problem, FEM>;
  LOP lop(problem);
  using MBE = PDELab::ISTL::BCRSMatrixBackend;
  using CC = typename GFS::template ConstraintsContainer<double>::Type;
  CC cc;
  using GO = PDELab::GridOperator<GFS, GFS, LOP, MBE, double, double, double, CC, CC>;
  GO go(gfs, cc, gfs, cc, lop, MBE(std::pow(2 * degree + 1, dim)));
  PDELab::constraints(go, cc);
  std::size_t expect_con = 1, inner = 1;
  for (int d = 0; d < dim; d++) expect_con *= degree * ncells + 1, inner *= degree * ncells - 1;
  EXPECT(cc.size() == expect_con - inner, name << ": constrained DOFs = " << cc.size());
  using V = typename GO::Domain;
  const std::size_t N = gfs.size();

  // oracle parity on random data
  std::mt19937_64 rng;
  std::uniform_real_distribution<double> dist(0, 1);
  V x(gfs), r(gfs, 0.0);
  for (std::size_t i = 0; i < N; i++) x[i] = dist(rng);
  go.residual(x, r);
  std::vector<double> want(N, 0.0);
  oracle_residual(&go.problem(), x.data(), want.data());
  EXPECT(rel_err(r, want) < 1e-12, name << ": residual vs oracle " << rel_err(r, want));
  typename GO::Jacobian A(go);
  A = 0.0;
  go.jacobian(x, A);
  uint64_t nr = 0, nnz = 0;
  oracle_pattern(&go.problem(), &nr, &nnz, nullptr, nullptr);
  std::vector<uint64_t> orp(nr + 1), oci(nnz);
  oracle_pattern(&go.problem(), &nr, &nnz, orp.data(), oci.data());
  EXPECT(orp == A.rowptr() && oci == A.colidx(), name << ": sparsity pattern bit-exact vs oracle (" << nnz << " nnz)");
  std::vector<double> ov(nnz, 0.0);
  oracle_jacobian(&go.problem(), x.data(), ov.data());
  EXPECT(rel_err(A.values(), ov) < 1e-12, name << ": jacobian values vs oracle " << rel_err(A.values(), ov));
  // constrained rows are unit rows (assemblerutilities.hh:666-684)
  bool unit = true;
  for (auto i : cc.dofs)
    for (uint64_t k2 = A.rowptr()[i]; k2 < A.rowptr()[i + 1]; k2++)
      unit &= A.values()[k2] == (A.colidx()[k2] == i ? 1.0 : 0.0);
  EXPECT(unit, name << ": constrained rows are identity rows");

  // solve: u = interpolate(g); z from J z = -R(u); matrix-based and matrix-free
  PDELab::ConvectionDiffusionDirichletExtensionAdapter<Problem> g(gv, problem);
  V u(gfs, 0.0);
  go.B200_interpolate(g, u);
  PDELab::set_nonconstrained_dofs(cc, 0.0, u);
  V res(gfs, 0.0), rhs(gfs, 0.0), z(gfs, 0.0);
  go.residual(u, res);
  rhs = res;
  rhs *= -1.0;
  // the Dirichlet-eliminated system is non-symmetric in its constrained columns (quirk viii) but the
  // right-hand side vanishes there, so CG on the free block is what the iteration sees
  PDELab::OnTheFlyOperator<V, V, GO> otf(go);
  int it_mat = cg([&](const V& a, V& b) { A.mv(a, b); }, rhs, z, 1e-12, 5000);
  V u_mat(u);
  u_mat += z;
  z = 0.0;
  int it_free = cg([&](const V& a, V& b) { otf.apply(a, b); }, rhs, z, 1e-12, 5000);
  V u_free(u);
  u_free += z;
  const double e_mat = l2_error_squared(gfs, u_mat, problem, go.handle(), degree);
  const double e_free = l2_error_squared(gfs, u_free, problem, go.handle(), degree);
  EXPECT(e_mat <= tol, name << ": l2errorsquared matrix based " << e_mat << " <= " << tol);
  EXPECT(e_free <= tol, name << ": l2errorsquared matrix free " << e_free << " <= " << tol);
  EXPECT(std::abs(it_mat - it_free) <= 1, name << ": CG iterations assembled/matrix-free (" << it_mat << ", " << it_free << ")");
}

// ---- test/test-blocked-istl-ordering.cc:24-72: L2 operator, QkDG k=2 on 4^3 cells, flat vs Blocking::fixed --------
void l2_blocked_ordering_case() {
  constexpr int dim = 3;
  using Grid = PDELab::YaspGrid<dim>;
  std::array<int, dim> cells;
  cells.fill(4);
  Grid grid(PDELab::FieldVector<double, dim>(1.0), cells);
  using GridView = typename Grid::LeafGridView;
  GridView gv = grid.leafGridView();
  using FEM = PDELab::QkDGLocalFiniteElementMap<double, double, 2, dim>;
  FEM fem;
  using FlatBackend = PDELab::ISTL::VectorBackend<>;
  using BlockedBackend = PDELab::ISTL::VectorBackend<PDELab::ISTL::Blocking::fixed, FEM::maxLocalSize()>;
  using FlatGFS = PDELab::GridFunctionSpace<GridView, FEM, PDELab::NoConstraints, FlatBackend>;
  using BlockedGFS = PDELab::GridFunctionSpace<GridView, FEM, PDELab::NoConstraints, BlockedBackend>;
  FlatGFS flat_gfs(gv, fem);
  BlockedGFS blocked_gfs(gv, fem);
  using LOP = PDELab::L2<GridView, FEM>;
  LOP lop;
  using MB = PDELab::ISTL::BCRSMatrixBackend;
  MB mb(1);
  using FlatGO = PDELab::GridOperator<FlatGFS, FlatGFS, LOP, MB, double, double, double>;
  using BlockedGO = PDELab::FastDGGridOperator<BlockedGFS, BlockedGFS, LOP, MB, double, double, double>;
  FlatGO flat_go(flat_gfs, flat_gfs, lop, mb);
  BlockedGO blocked_go(blocked_gfs, blocked_gfs, lop, mb);
  typename FlatGO::Domain flat_x(flat_gfs, 0.0), flat_r(flat_gfs, 0.0);
  typename BlockedGO::Domain blocked_x(blocked_gfs, 0.0), blocked_r(blocked_gfs, 0.0);
  std::mt19937_64 rng;
  std::uniform_real_distribution<double> dist(0.0, 1.0);
  const std::size_t N = flat_gfs.size();
  for (std::size_t i = 0; i < N; i++) blocked_x[i] = flat_x[i] = dist(rng);
  flat_go.residual(flat_x, flat_r);
  blocked_go.residual(blocked_x, blocked_r);
  bool same = true;
  for (std::size_t i = 0; i < N; i++) same = same && flat_r[i] == blocked_r[i];
  EXPECT(same, "L2 QkDG k=2 4^3 (test-blocked-istl-ordering): flat and blocked residuals coincide");
  // and the values are the mass operator: compare with the oracle fed with the same problem, and with
  // the closed form  int u * 1 = sum_i r_i  = sum over cells of |K| * mean-weighted u
  std::vector<double> want(N, 0.0);
  if (oracle_residual(&flat_go.problem(), flat_x.data(), want.data())) std::cerr << oracle_last_error() << std::endl;
  EXPECT(rel_err(flat_r, want) < 1e-12, "L2 residual vs oracle " << rel_err(flat_r, want) << " [" << flat_go.lastKernel() << "]");
  // 1^T M x = int u_h: with the Q2 Lagrange weights (1/6, 4/6, 1/6) per direction
  double sum_r = 0, integral = 0;
  const double w1[3] = {1.0 / 6, 4.0 / 6, 1.0 / 6};
  for (std::size_t i = 0; i < N; i++) {
    sum_r += flat_r[i];
    const int l = (int)(i % 27);
    integral += flat_x[i] * w1[l % 3] * w1[(l / 3) % 3] * w1[l / 9] / 64.0;
  }
  EXPECT(std::abs(sum_r - integral) < 1e-13, "L2: sum of the residual equals the integral of u_h (" << sum_r << ")");
}

// ---- test/testinstationaryfastdgassembler.cc: OneStepGridOperator + OneStepMethod ------------------------
// ParameterA of the reference test (:17-104): A = I, f = (2 d - 4 |x|^2) exp(-|x|^2), Dirichlet g = exp(-|x|^2)
template <typename GV, typename RF>
class ParameterA : public PDELab::ConvectionDiffusionModelProblem<GV, RF> {
 public:
  template <typename Element, typename Coord>
  RF f(const Element& e, const Coord& x) const {
    auto xg = e.geometry().global(x);
    double norm = 0;
    for (std::size_t i = 0; i < xg.size(); i++) norm += xg[i] * xg[i];
    return (2.0 * xg.size() - 4.0 * norm) * std::exp(-norm);
  }
  template <typename Element, typename Coord>
  RF g(const Element& e, const Coord& x) const {
    auto xg = e.geometry().global(x);
    double norm = 0;
    for (std::size_t i = 0; i < xg.size(); i++) norm += xg[i] * xg[i];
    return std::exp(-norm);
  }
  void setTime(RF t) { time = t; }
  RF time = 0.0;
};

void instationary_dg_case() {
  constexpr int dim = 2, degree = 1;
  const char* name = "instationary DG k=1 2D 8^2 (testinstationaryfastdgassembler)";
  using Grid = PDELab::YaspGrid<dim>;
  std::array<int, dim> cells;
  cells.fill(8);  // one cell, globalRefine(3)
  Grid grid(PDELab::FieldVector<double, dim>(1.0), cells);
  using GV = typename Grid::LeafGridView;
  GV gv = grid.leafGridView();
  using FEM = PDELab::QkDGLocalFiniteElementMap<double, double, degree, dim>;
  FEM fem;
  using VBE = PDELab::ISTL::VectorBackend<PDELab::ISTL::Blocking::fixed, FEM::maxLocalSize()>;
  using GFS = PDELab::GridFunctionSpace<GV, FEM, PDELab::NoConstraints, VBE>;
  GFS gfs(gv, fem);
  using Problem = ParameterA<GV, double>;
  Problem problem;
  using LOP = PDELab::ConvectionDiffusionDG<Problem, FEM>;
  LOP lop(problem, PDELab::ConvectionDiffusionDGMethod::SIPG, PDELab::ConvectionDiffusionDGWeights::weightsOn, 2.0);
  using MLOP = PDELab::L2<GV, FEM>;
  MLOP mlop(2 * degree);
  using MBE = PDELab::ISTL::BCRSMatrixBackend;
  MBE mbe(9);
  using CC = typename GFS::template ConstraintsContainer<double>::Type;
  CC cc;
  using GO0 = PDELab::FastDGGridOperator<GFS, GFS, LOP, MBE, double, double, double, CC, CC>;
  GO0 go0(gfs, cc, gfs, cc, lop, mbe);
  using GO1 = PDELab::FastDGGridOperator<GFS, GFS, MLOP, MBE, double, double, double, CC, CC>;
  GO1 go1(gfs, cc, gfs, cc, mlop, mbe);
  using IGO = PDELab::OneStepGridOperator<GO0, GO1>;
  IGO igo(go0, go1);
  using V = typename IGO::Traits::Domain;
  const std::size_t N = gfs.size();
  using G = PDELab::ConvectionDiffusionDirichletExtensionAdapter<Problem>;
  G g(gv, problem);
  V x(gfs, 0.0);
  go0.B200_interpolate(g, x);  // Dune::PDELab::interpolate(g, gfs, x)

  // --- the one-step residual against the engines restated on the oracle (stage 1 of Alexander2:
  //     a10 = -1, b10 = 0, b11 = alpha):  r = b11 dt R0(y) + R1(y) - R1(x)
  PDELab::Alexander2Parameter<double> method;
  const double dt = 0.1;
  {
    std::mt19937_64 rng;
    std::uniform_real_distribution<double> dist(0, 1);
    V y(gfs), r(gfs, 0.0);
    for (std::size_t i = 0; i < N; i++) y[i] = dist(rng);
    igo.preStep(method, 0.0, dt);
    std::vector<V*> xs(1, &x);
    igo.preStage(1, xs);
    igo.residual(y, r);
    std::vector<double> r0(N, 0.0), r1y(N, 0.0), r1x(N, 0.0), want(N);
    oracle_residual(&go0.problem(), y.data(), r0.data());
    oracle_residual(&go1.problem(), y.data(), r1y.data());
    oracle_residual(&go1.problem(), x.data(), r1x.data());
    for (std::size_t i = 0; i < N; i++) want[i] = method.b(1, 1) * dt * r0[i] + r1y[i] - r1x[i];
    EXPECT(rel_err(r, want) < 1e-12, name << ": one-step residual vs oracle engines " << rel_err(r, want));
    EXPECT(std::abs(igo.timeAtStage(1) - method.d(1) * dt) < 1e-15, name << ": timeAtStage");
    // assembled one-step Jacobian times z == matrix-free one-step jacobian_apply
    typename IGO::Jacobian A(igo);
    A = 0.0;
    igo.jacobian(y, A);
    V Jy(gfs, 0.0), Jf(gfs, 0.0);
    A.mv(y, Jy);
    igo.jacobian_apply(y, Jf);
    EXPECT(rel_err(Jf, Jy.native()) < 1e-12, name << ": one-step jacobian * z == jacobian_apply " << rel_err(Jf, Jy.native()));
    bool threw = false;
    try {
      igo.jacobian_apply(y, y, r);
    } catch (PDELab::Exception&) {
      threw = true;
    }
    EXPECT(threw, name << ": non-linear jacobian_apply throws for linear operators (onestep.hh:187-192)");
  }

  // --- the time loop of the reference test, matrix-free (CG + block Jacobi) and assembled (CG + Jacobi)
  auto run = [&](auto& ls, const char* what) {
    using LS = std::decay_t<decltype(ls)>;
    using PDESOLVER = PDELab::StationaryLinearProblemSolver<IGO, LS, V>;
    PDESOLVER pdesolver(igo, ls, 1e-10);
    PDELab::OneStepMethod<double, IGO, PDESOLVER, V, V> osm(method, igo, pdesolver);
    osm.setVerbosityLevel(0);
    V xt(x);
    double time = 0.0;
    const double T = 0.1;
    while (time < T - 1e-10) {
      V xnew(gfs, 0.0);
      osm.apply(time, dt, xt, xnew);
      xt = xnew;
      time += dt;
    }
    const double err = l2_error_squared(gfs, xt, problem, go0.handle(), degree);
    EXPECT(err <= 5e-6 && !std::isnan(err), name << ": " << what << " l2 error squared " << err << " <= 5e-6, "
                                                 << osm.result().total.linear_solver_iterations << " linear iterations");
    return xt;
  };
  PDELab::ISTLBackend_SEQ_MatrixFree_CG_BlockJacobi<IGO> ls_free(igo, 10000, 0);
  PDELab::ISTLBackend_SEQ_CG_Jac<IGO> ls_mat(igo, 10000, 0);
  V x_free = run(ls_free, "matrix-free CG + block Jacobi");
  V x_mat = run(ls_mat, "assembled CG + Jacobi");
  EXPECT(rel_err(x_free, x_mat.native()) < 1e-8, name << ": both solvers reach the same state " << rel_err(x_free, x_mat.native()));

  // --- the same problem stepped explicitly (ExplicitOneStepMethod, instationary/explicitonestep.hh): RK4, dt well
  // inside the diffusive stability limit; the mass solve is the exact block inverse on the device
  {
    using EIGO = PDELab::OneStepGridOperator<GO0, GO1, false>;
    EIGO eigo(go0, go1);
    PDELab::RK4Parameter<double> rk4;
    using ELS = PDELab::ISTLBackend_SEQ_MatrixFree_CG_BlockJacobi<EIGO>;
    ELS els(eigo, 100, 0);
    PDELab::ExplicitOneStepMethod<double, EIGO, ELS, V, V> eosm(rk4, eigo, els);
    eosm.setVerbosityLevel(0);
    V xt(x);
    double time = 0.0;
    for (int step = 0; step < 20; step++) {
      V xnew(gfs, 0.0);
      eosm.apply(time, 2e-5, xt, xnew);
      xt = xnew;
      time += 2e-5;
    }
    const double err = l2_error_squared(gfs, xt, problem, go0.handle(), degree);
    EXPECT(err <= 5e-6 && !std::isnan(err), name << ": explicit RK4, 20 steps: l2 error squared " << err << " <= 5e-6");
    bool threw = false;
    try {
      V r(gfs, 0.0);
      eigo.residual(xt, r);
    } catch (PDELab::Exception&) {
      threw = true;
    }
    EXPECT(threw, name << ": residual() of the explicit operator throws (onestep.hh:143-144)");
    threw = false;
    try {
      PDELab::ExplicitOneStepMethod<double, EIGO, ELS, V, V> bad(method, eigo, els);
    } catch (PDELab::Exception&) {
      threw = true;
    }
    EXPECT(threw, name << ": explicit method with an implicit scheme throws (explicitonestep.hh:226-228)");
  }
}

// conforming Q2 with Dirichlet constraints interpolated at every stage (implicitonestep.hh:264-400)
void instationary_fem_case() {
  constexpr int dim = 2, degree = 2;
  const char* name = "instationary Q2 2D 16^2, fractional step";
  using Grid = PDELab::YaspGrid<dim>;
  std::array<int, dim> cells;
  cells.fill(16);
  Grid grid(PDELab::FieldVector<double, dim>(1.0), cells);
  using GV = typename Grid::LeafGridView;
  GV gv = grid.leafGridView();
  using FEM = PDELab::QkLocalFiniteElementMap<GV, double, double, degree>;
  FEM fem(gv);
  using GFS = PDELab::GridFunctionSpace<GV, FEM, PDELab::ConformingDirichletConstraints, PDELab::ISTL::VectorBackend<>>;
  GFS gfs(gv, fem);
  using Problem = ParameterA<GV, double>;
  Problem problem;
  using LOP = PDELab::ConvectionDiffusionFEM<Problem, FEM>;
  LOP lop(problem);
  using MLOP = PDELab::L2<GV, FEM>;
  MLOP mlop;
  using MBE = PDELab::ISTL::BCRSMatrixBackend;
  using CC = typename GFS::template ConstraintsContainer<double>::Type;
  CC cc;
  using GO0 = PDELab::GridOperator<GFS, GFS, LOP, MBE, double, double, double, CC, CC>;
  using GO1 = PDELab::GridOperator<GFS, GFS, MLOP, MBE, double, double, double, CC, CC>;
  GO0 go0(gfs, cc, gfs, cc, lop, MBE(25));
  GO1 go1(gfs, cc, gfs, cc, mlop, MBE(25));
  PDELab::constraints(go0, cc);
  using IGO = PDELab::OneStepGridOperator<GO0, GO1>;
  IGO igo(go0, go1);
  using V = typename IGO::Traits::Domain;
  PDELab::ConvectionDiffusionDirichletExtensionAdapter<Problem> g(gv, problem);
  V x(gfs, 0.0);
  go0.B200_interpolate(g, x);
  PDELab::set_nonconstrained_dofs(cc, 0.0, x);  // start from zero in the interior: the heat equation relaxes to exp(-|x|^2)
  using LS = PDELab::ISTLBackend_SEQ_MatrixFree_CG_Richardson<IGO>;
  LS ls(igo, 10000, 0);
  using PDESOLVER = PDELab::StationaryLinearProblemSolver<IGO, LS, V>;
  PDESOLVER pdesolver(igo, ls, 1e-10);
  PDELab::FractionalStepParameter<double> method;
  PDELab::OneStepMethod<double, IGO, PDESOLVER, V, V> osm(method, igo, pdesolver);
  osm.setVerbosityLevel(0);
  double time = 0.0;
  const double dt = 0.05;
  for (int step = 0; step < 40; step++) {  // T = 2: the slowest mode has decayed by exp(-2 * 2 pi^2); the scheme is only
                                           // strongly A-stable (|R(inf)| ~ 0.7), the boundary layer needs the 40 steps
    V xnew(x);
    osm.apply(time, dt, x, g, xnew);
    x = xnew;
    time += dt;
  }
  const double err = l2_error_squared(gfs, x, problem, go0.handle(), degree);
  EXPECT(err <= 1e-6 && !std::isnan(err), name << ": l2 error squared after relaxation " << err << " <= 1e-6, "
                                               << osm.result().total.linear_solver_iterations << " linear iterations");
  // the constrained DOFs carry the interpolated boundary values
  V gi(gfs, 0.0);
  go0.B200_interpolate(g, gi);
  bool same = true;
  for (auto i : cc.dofs) same = same && x[i] == gi[i];
  EXPECT(same, name << ": Dirichlet DOFs hold the interpolated boundary values");
}

int main() {
  try {
    l2_blocked_ordering_case();
    instationary_dg_case();
    instationary_fem_case();
    dg_case<2, 1, PoissonProblem>(16, 3.0, 1e-6, "DG k=1 2D 16^2 (testconvectiondiffusiondg)", true);
    dg_case<2, 2, LayeredProblem>(6, 3.0, 0, "DG k=2 2D 6^2 layered diagonal A + c", false);
    dg_case<3, 2, LayeredProblem>(4, 3.0, 0, "DG k=2 3D 4^3 layered diagonal A + c (Kronecker kernel)", false);
    dg_case<3, 2, PoissonProblem>(8, 3.0, 1e-6, "DG k=2 3D 8^3 Poisson (Kronecker kernel)", true);
    dg_case<2, 2, RotatingProblem>(5, 3.0, 0, "DG k=2 2D 5^2 rotating b(x), c(x), A(x), bctype per face point", false);
    dg_case<3, 1, RotatingProblem>(3, 3.0, 0, "DG k=1 3D 3^3 rotating b(x), c(x), A(x), bctype per face point", false);
    fem_case<2, 2>(32, 1e-7, "Q2 2D 32^2 (testmatrixfree)");
    fem_case<2, 1>(16, 1e-4, "Q1 2D 16^2");
    fem_case<3, 2>(4, 1e-5, "Q2 3D 4^3");
  } catch (std::exception& e) {
    std::cerr << "exception: " << e.what() << std::endl;
    return 2;
  }
  std::cout << (failures ? "FAILED" : "ALL OK") << " (" << failures << " failures)" << std::endl;
  return failures ? 1 : 0;
}

// gridoperator.hh — C++ host mirror of PDELab's operator-evaluation interface over the C ABI of
// include/pdelab_b200.h (libpdelab_b200.so, hand-written sm_100a kernels).  Header-only, C++17.
//
// What it mirrors (paths relative to /root/reference/dune/pdelab/):
//   GridOperator<GFSU,GFSV,LOP,MB,DF,RF,JF,CU,CV>        gridoperator/gridoperator.hh:30-244
//   GridOperatorTraits                                  gridoperator/common/gridoperatorutilities.hh:31-80
//   ConvectionDiffusionDG<Param,FEM> (ctor arguments)   localoperator/convectiondiffusiondg.hh:85-102
//   ConvectionDiffusionFEM<Param,FEM>                   localoperator/convectiondiffusionfem.hh:38-61
//   ConvectionDiffusionBoundaryConditions / parameter-class call-backs
//                                                       localoperator/convectiondiffusionparameter.hh:111-209
//   Backend::Vector / Backend::native                   backend/istl/vector.hh, backend/interface.hh
//   ISTL::BCRSMatrixBackend, BCRSMatrix(go)             backend/istl/bcrsmatrixbackend.hh, bcrsmatrix.hh:78-82
//   OnTheFlyOperator                                    backend/istl/seqistlsolverbackend.hh:44-100
//   constraints(), interpolate(), set_nonconstrained_dofs()
//                                                       constraints/common/constraints.hh:588-687,796-802,
//                                                       gridfunctionspace/interpolate.hh
// Same member names, argument meaning and error behaviour (exceptions; B200::Exception plays the
// role of Dune::Exception), so a PDELab program switches by changing the namespace of these types.
// The user's parameter class keeps the reference's call-back interface; the call-backs are sampled
// once, at exactly the points the reference evaluates them, into the arrays the C ABI takes.
//
// The structured grid, finite element maps and function spaces are light descriptors (the CUDA
// path derives all index maps in closed form); with the DUNE core modules present a maintainer
// wires the real types in as shown in INTEGRATION.md.
//
// There is no CPU implementation behind this header: every compute member forwards to the CUDA
// library and throws if it reports an error (e.g. "no CUDA device available").
#ifndef PDELAB_B200_HOST_GRIDOPERATOR_HH
#define PDELAB_B200_HOST_GRIDOPERATOR_HH

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <memory>
#include <string>
#include <type_traits>
#include <tuple>
#include <vector>

#include "../../include/pdelab_b200.h"

namespace Dune {
namespace PDELab {
namespace B200 {

// ---- error convention -------------------------------------------------------------------------
// The reference throws Dune::Exception (DUNE_THROW, e.g. gridoperator.hh:195,203).
class Exception : public std::exception {
 public:
  explicit Exception(std::string msg) : msg_(std::move(msg)) {}
  const char* what() const noexcept override { return msg_.c_str(); }

 private:
  std::string msg_;
};

inline void check(int rc, const char* where) {
  if (rc != 0) throw Exception(std::string(where) + ": " + pdb200_last_error());
}

// ---- dense helpers (stand-ins for Dune::FieldVector / FieldMatrix) ------------------------------
template <class T, int n>
struct FieldVector : std::array<T, n> {
  FieldVector() { this->fill(T(0)); }
  FieldVector(T v) { this->fill(v); }  // NOLINT: Dune::FieldVector is implicitly constructible from a scalar
  FieldVector(std::initializer_list<T> l) {
    this->fill(T(0));
    std::copy(l.begin(), l.end(), this->begin());
  }
  static constexpr int dimension = n;
  T two_norm2() const {
    T s = 0;
    for (T v : *this) s += v * v;
    return s;
  }
  T two_norm() const { return std::sqrt(two_norm2()); }
  FieldVector& operator-=(const FieldVector& o) {
    for (int i = 0; i < n; i++) (*this)[i] -= o[i];
    return *this;
  }
  FieldVector& operator+=(const FieldVector& o) {
    for (int i = 0; i < n; i++) (*this)[i] += o[i];
    return *this;
  }
  T operator*(const FieldVector& o) const {
    T s = 0;
    for (int i = 0; i < n; i++) s += (*this)[i] * o[i];
    return s;
  }
};
template <class T, int n, int m>
struct FieldMatrix : std::array<FieldVector<T, m>, n> {
  FieldMatrix() = default;
  FieldMatrix(T v) {  // NOLINT
    for (auto& r : *this) r.fill(v);
  }
};

// ---- structured grid --------------------------------------------------------------------------
template <int dim>
class YaspGridView;

// Dune::YaspGrid<dim>(L, N): equidistant axis-aligned grid on [0,L] (dune-grid, used by every
// reference test of this path, e.g. test/testconvectiondiffusiondg.cc:52-58)
template <int dim_>
class YaspGrid {
 public:
  static constexpr int dimension = dim_;
  using ctype = double;
  using LeafGridView = YaspGridView<dim_>;
  YaspGrid(const FieldVector<double, dim_>& L, const std::array<int, dim_>& N) : upper_(L), cells_(N) {}
  YaspGrid(const FieldVector<double, dim_>& lower, const FieldVector<double, dim_>& upper, const std::array<int, dim_>& N)
      : lower_(lower), upper_(upper), cells_(N) {}
  LeafGridView leafGridView() const;
  const FieldVector<double, dim_>& lower() const { return lower_; }
  const FieldVector<double, dim_>& upper() const { return upper_; }
  const std::array<int, dim_>& cells() const { return cells_; }
  // outer sides that are processor boundaries of an overlapping partition
  std::array<std::array<int, 2>, 3> side_kind{{{0, 0}, {0, 0}, {0, 0}}};

 private:
  FieldVector<double, dim_> lower_{}, upper_;
  std::array<int, dim_> cells_;
};

// geometry of an axis-aligned cell or face: the closed forms of SURVEY.md Appendix A
template <int dim, int mydim>
struct BoxGeometry {
  FieldVector<double, dim> lo, ext;  // lower corner, edge lengths (ext[dir] = 0 for a face)
  int normal_dir = -1;               // face: direction of the normal
  FieldVector<double, dim> global(const FieldVector<double, mydim>& x) const {
    FieldVector<double, dim> g = lo;
    int t = 0;
    for (int d = 0; d < dim; d++)
      if (d != normal_dir) g[d] += ext[d] * x[t++];
    return g;
  }
  FieldVector<double, dim> center() const { return global(FieldVector<double, mydim>(0.5)); }
  double volume() const {
    double v = 1;
    for (int d = 0; d < dim; d++)
      if (d != normal_dir) v *= ext[d];
    return v;
  }
};

template <int dim>
struct Cell {
  static constexpr int dimension = dim;
  std::array<int, dim> coord;
  long long index;
  BoxGeometry<dim, dim> geo;
  const BoxGeometry<dim, dim>& geometry() const { return geo; }
};

template <int dim>
struct Intersection {
  Cell<dim> inside_;
  int dir, side;  // indexInInside = 2*dir + side (YaspGrid)
  BoxGeometry<dim, dim - 1> geo;
  const Cell<dim>& inside() const { return inside_; }
  int indexInInside() const { return 2 * dir + side; }
  bool boundary() const { return true; }
  const BoxGeometry<dim, dim - 1>& geometry() const { return geo; }
  // reference coordinates of a face point inside the cell (geometryInInside().global(x))
  struct InInside {
    int dir, side;
    FieldVector<double, dim> global(const FieldVector<double, dim - 1>& x) const {
      FieldVector<double, dim> g;
      int t = 0;
      for (int d = 0; d < dim; d++) g[d] = d == dir ? double(side) : x[t++];
      return g;
    }
  };
  InInside geometryInInside() const { return InInside{dir, side}; }
  FieldVector<double, dim> centerUnitOuterNormal() const {
    FieldVector<double, dim> n;
    n[dir] = side ? 1.0 : -1.0;
    return n;
  }
  FieldVector<double, dim> unitOuterNormal(const FieldVector<double, dim - 1>&) const { return centerUnitOuterNormal(); }
};

template <int dim_>
class YaspGridView {
 public:
  static constexpr int dimension = dim_;
  using ctype = double;
  using Grid = YaspGrid<dim_>;
  explicit YaspGridView(const Grid& g) : grid_(&g) {}
  const Grid& grid() const { return *grid_; }
  long long size(int codim) const {
    if (codim != 0) throw Exception("YaspGridView::size: only codim 0 is counted by the descriptor");
    long long n = 1;
    for (int v : grid_->cells()) n *= v;
    return n;
  }
  double h(int d) const { return (grid_->upper()[d] - grid_->lower()[d]) / grid_->cells()[d]; }
  Cell<dim_> cell(long long index) const {
    Cell<dim_> c;
    c.index = index;
    long long r = index;
    for (int d = 0; d < dim_; d++) {
      c.coord[d] = int(r % grid_->cells()[d]);
      r /= grid_->cells()[d];
      c.geo.ext[d] = h(d);
      c.geo.lo[d] = grid_->lower()[d] + h(d) * c.coord[d];
    }
    return c;
  }
  // boundary face with the numbering of pdelab_b200.h (direction-major, tangential lexicographic)
  Intersection<dim_> boundaryFace(int dir, int side, long long tang) const {
    std::array<int, dim_> cc{};
    long long r = tang;
    for (int d = 0; d < dim_; d++)
      if (d != dir) {
        cc[d] = int(r % grid_->cells()[d]);
        r /= grid_->cells()[d];
      }
    cc[dir] = side ? grid_->cells()[dir] - 1 : 0;
    long long idx = 0, stride = 1;
    for (int d = 0; d < dim_; d++) {
      idx += stride * cc[d];
      stride *= grid_->cells()[d];
    }
    Intersection<dim_> is;
    is.inside_ = cell(idx);
    is.dir = dir;
    is.side = side;
    is.geo.lo = is.inside_.geo.lo;
    is.geo.ext = is.inside_.geo.ext;
    is.geo.normal_dir = dir;
    if (side) is.geo.lo[dir] += is.inside_.geo.ext[dir];
    return is;
  }

 private:
  const Grid* grid_;
};
template <int dim_>
inline YaspGridView<dim_> YaspGrid<dim_>::leafGridView() const {
  return YaspGridView<dim_>(*this);
}

// ---- finite element maps and function spaces ------------------------------------------------------
// QkDGBasisPolynomial (finiteelementmap/qkdg.hh:15; l2orthonormal needs dune-localfunctions' OPB machinery: not mirrored)
enum class QkDGBasisPolynomial { lagrange = PDB200_BASIS_LAGRANGE, legendre = PDB200_BASIS_LEGENDRE, lobatto = PDB200_BASIS_LOBATTO };
// QkDGLocalFiniteElementMap<D,R,k,d,p> (finiteelementmap/qkdg.hh:17-200): Lagrange (default), Legendre, Gauss-Lobatto
template <class D, class R, int k, int d, QkDGBasisPolynomial p = QkDGBasisPolynomial::lagrange>
struct QkDGLocalFiniteElementMap {
  static constexpr int degree = k, dimension = d, space = PDB200_SPACE_QKDG, basis = (int)p;
  static constexpr QkDGBasisPolynomial polynomial() { return p; }
  static constexpr std::size_t maxLocalSize() {
    std::size_t n = 1;
    for (int i = 0; i < d; i++) n *= k + 1;
    return n;
  }
};
// QkLocalFiniteElementMap<GV,D,R,k> (finiteelementmap/qkfem.hh:17-78)
template <class GV, class D, class R, int k>
struct QkLocalFiniteElementMap {
  static constexpr int degree = k, dimension = GV::dimension, space = PDB200_SPACE_QK, basis = PDB200_BASIS_LAGRANGE;
  explicit QkLocalFiniteElementMap(const GV&) {}
  QkLocalFiniteElementMap() = default;
  static constexpr std::size_t maxLocalSize() {
    std::size_t n = 1;
    for (int i = 0; i < GV::dimension; i++) n *= k + 1;
    return n;
  }
};

struct NoConstraints {};
struct ConformingDirichletConstraints {};  // constraints/conforming.hh:36-139
struct P0ParallelConstraints {};           // constraints/p0.hh (selected through YaspGrid::side_kind)

namespace ISTL {
enum class Blocking { none, fixed };
// flat and fixed-block vectors are the same bytes (test/test-blocked-istl-ordering.cc:65-72)
template <Blocking blocking = Blocking::none, std::size_t block_size = 1>
struct VectorBackend {};
struct BCRSMatrixBackend {
  explicit BCRSMatrixBackend(std::size_t entries_per_row = 0) : entries_per_row_(entries_per_row) {}
  std::size_t avg_entries_per_row() const { return entries_per_row_; }
  std::size_t entries_per_row_;
};
}  // namespace ISTL

template <class GV, class FEM, class CON = NoConstraints, class VBE = ISTL::VectorBackend<>>
class GridFunctionSpace {
 public:
  using Traits = GridFunctionSpace;
  using GridViewType = GV;
  using GridView = GV;
  using FiniteElementMapType = FEM;
  using ConstraintsType = CON;
  using Backend = VBE;
  using SizeType = std::size_t;
  // GridFunctionSpace::ConstraintsContainer<E>::Type: the constrained DOFs (all with empty
  // linear combination = Dirichlet), constraints/common/constraintstransformation.hh
  template <class E>
  struct ConstraintsContainer {
    struct Type {
      std::vector<std::uint64_t> dofs;  // ascending
      void clear() { dofs.clear(); }
      std::size_t size() const { return dofs.size(); }
      bool containsNonDirichletConstraints() const { return false; }
    };
  };
  GridFunctionSpace(const GV& gv, const FEM& fem) : gv_(gv), fem_(fem) {}
  const GV& gridView() const { return gv_; }
  const FEM& finiteElementMap() const { return fem_; }
  void name(const std::string& n) { name_ = n; }
  const std::string& name() const { return name_; }
  void update() {}
  SizeType maxLocalSize() const { return FEM::maxLocalSize(); }
  SizeType size() const {
    SizeType n = 1;
    if (FEM::space == PDB200_SPACE_QKDG) {
      n = FEM::maxLocalSize();
      for (int v : gv_.grid().cells()) n *= v;
    } else {
      for (int v : gv_.grid().cells()) n *= FEM::degree * v + 1;
    }
    return n;
  }
  SizeType globalSize() const { return size(); }

 private:
  GV gv_;
  FEM fem_;
  std::string name_;
};

// ---- vectors ------------------------------------------------------------------------------------
namespace Backend {
// Backend::Vector<GFS,E>: one contiguous E[N] in container order (backend/istl/vector.hh)
template <class GFS, class E>
class Vector {
 public:
  using ElementType = E;
  using Container = std::vector<E>;
  using GridFunctionSpace = GFS;
  explicit Vector(const GFS& gfs, E v = E(0)) : gfs_(&gfs), data_(gfs.size(), v) {}
  Vector& operator=(E v) {
    std::fill(data_.begin(), data_.end(), v);
    return *this;
  }
  std::size_t N() const { return data_.size(); }
  std::size_t flatsize() const { return data_.size(); }
  E* data() { return data_.data(); }
  const E* data() const { return data_.data(); }
  E& operator[](std::size_t i) { return data_[i]; }
  const E& operator[](std::size_t i) const { return data_[i]; }
  Vector& operator+=(const Vector& o) {
    for (std::size_t i = 0; i < data_.size(); i++) data_[i] += o.data_[i];
    return *this;
  }
  Vector& operator-=(const Vector& o) {
    for (std::size_t i = 0; i < data_.size(); i++) data_[i] -= o.data_[i];
    return *this;
  }
  Vector& operator*=(E a) {
    for (E& v : data_) v *= a;
    return *this;
  }
  Vector& axpy(E a, const Vector& x) {
    for (std::size_t i = 0; i < data_.size(); i++) data_[i] += a * x.data_[i];
    return *this;
  }
  E dot(const Vector& o) const {
    E s = 0;
    for (std::size_t i = 0; i < data_.size(); i++) s += data_[i] * o.data_[i];
    return s;
  }
  E two_norm() const { return std::sqrt(dot(*this)); }
  E infinity_norm() const {
    E s = 0;
    for (E v : data_) s = std::max(s, std::abs(v));
    return s;
  }
  const GFS& gridFunctionSpace() const { return *gfs_; }
  Container& native() { return data_; }
  const Container& native() const { return data_; }

 private:
  const GFS* gfs_;
  Container data_;
};
template <class V>
auto native(V& v) -> decltype(v.native()) {
  return v.native();
}
}  // namespace Backend

// ---- boundary-condition and DG enums (same names as the reference) -------------------------------
struct ConvectionDiffusionBoundaryConditions {
  enum Type { Dirichlet = 1, Neumann = -1, Outflow = -2, None = -3 };  // convectiondiffusionparameter.hh:113
};
struct ConvectionDiffusionDGMethod {
  enum Type { NIPG, SIPG, IIPG };  // convectiondiffusiondg.hh:31
};
struct ConvectionDiffusionDGWeights {
  enum Type { weightsOn, weightsOff };  // convectiondiffusiondg.hh:36
};

// Traits the user's parameter class is written against (convectiondiffusionparameter.hh:39-105)
template <class GV, class RF>
struct ConvectionDiffusionParameterTraits {
  using GridViewType = GV;
  static constexpr int dimDomain = GV::dimension;
  using DomainFieldType = double;
  using DomainType = FieldVector<double, GV::dimension>;
  using IntersectionDomainType = FieldVector<double, GV::dimension - 1>;
  using RangeFieldType = RF;
  using RangeType = FieldVector<RF, GV::dimension>;
  using PermTensorType = FieldMatrix<RF, GV::dimension, GV::dimension>;
  using ElementType = Cell<GV::dimension>;
  using IntersectionType = Intersection<GV::dimension>;
};

// ConvectionDiffusionModelProblem (convectiondiffusionparameter.hh:124-209): the defaults a user's
// parameter class inherits — A = I, b = 0, c = 0, f = 0, Dirichlet everywhere, g = j = o = 0.
template <class GV, class RF>
class ConvectionDiffusionModelProblem {
 public:
  using BCType = ConvectionDiffusionBoundaryConditions::Type;
  using Traits = ConvectionDiffusionParameterTraits<GV, RF>;
  static constexpr bool permeabilityIsConstantPerCell() { return true; }
  template <class E, class X>
  typename Traits::PermTensorType A(const E&, const X&) const {
    typename Traits::PermTensorType I(RF(0));
    for (int i = 0; i < Traits::dimDomain; i++) I[i][i] = 1.0;
    return I;
  }
  template <class E, class X>
  typename Traits::RangeType b(const E&, const X&) const {
    return typename Traits::RangeType(RF(0));
  }
  template <class E, class X>
  RF c(const E&, const X&) const {
    return 0.0;
  }
  template <class E, class X>
  RF f(const E&, const X&) const {
    return 0.0;
  }
  template <class I, class X>
  BCType bctype(const I&, const X&) const {
    return ConvectionDiffusionBoundaryConditions::Dirichlet;
  }
  template <class E, class X>
  RF g(const E&, const X&) const {
    return 0.0;
  }
  template <class I, class X>
  RF j(const I&, const X&) const {
    return 0.0;
  }
  template <class I, class X>
  RF o(const I&, const X&) const {
    return 0.0;
  }
  void setTime(RF) {}
};

// ---- local operators: carry the constructor arguments of the reference ----------------------------
template <class Param, class FEM>
class ConvectionDiffusionDG {
 public:
  static constexpr bool isDG = true;
  using ParameterType = Param;
  // convectiondiffusiondg.hh:85-102 (same defaults)
  ConvectionDiffusionDG(Param& param, ConvectionDiffusionDGMethod::Type method = ConvectionDiffusionDGMethod::SIPG,
                        ConvectionDiffusionDGWeights::Type weights = ConvectionDiffusionDGWeights::weightsOn,
                        double alpha = 1.0, int intorderadd = 0)
      : param_(&param), method(method), weights(weights), alpha(alpha), intorderadd(intorderadd) {}
  static constexpr bool isLinear = true;
  Param& parameters() const { return *param_; }
  void setTime(double t) { param_->setTime(t); }
  Param* param_;
  ConvectionDiffusionDGMethod::Type method;
  ConvectionDiffusionDGWeights::Type weights;
  double alpha;
  int intorderadd;
};

template <class Param, class FEM>
class ConvectionDiffusionFEM {
 public:
  static constexpr bool isDG = false;
  using ParameterType = Param;
  explicit ConvectionDiffusionFEM(Param& param, int intorderadd = 0) : param_(&param), intorderadd(intorderadd) {}
  static constexpr bool isLinear = true;
  Param& parameters() const { return *param_; }
  void setTime(double t) { param_->setTime(t); }
  Param* param_;
  int intorderadd;
  // members the DG operator has; unused by the conforming path
  ConvectionDiffusionDGMethod::Type method = ConvectionDiffusionDGMethod::SIPG;
  ConvectionDiffusionDGWeights::Type weights = ConvectionDiffusionDGWeights::weightsOn;
  double alpha = 0.0;
};

// L2 (localoperator/l2.hh:25-230): the mass operator  alpha_volume = scaling * int u v  (no skeleton or
// boundary terms, no constraints of its own) — the local operator of test/test-blocked-istl-ordering.cc.
// It is the convection-diffusion operator with A = 0, b = 0, c = scaling and boundary type None, so the
// same kernels evaluate it (for QkDG the face coefficients vanish with A).  residual / jacobian_apply
// only: the reference's L2 pattern is block-diagonal (doPatternVolume only), which the assembled path of
// this library does not produce.
namespace detail {
template <class GV, class RF>
class L2Parameter : public ConvectionDiffusionModelProblem<GV, RF> {
 public:
  using Traits = ConvectionDiffusionParameterTraits<GV, RF>;
  using BCType = ConvectionDiffusionBoundaryConditions::Type;
  explicit L2Parameter(RF scaling) : scaling_(scaling) {}
  template <class E, class X>
  typename Traits::PermTensorType A(const E&, const X&) const {
    return typename Traits::PermTensorType(RF(0));
  }
  template <class E, class X>
  RF c(const E&, const X&) const {
    return scaling_;
  }
  template <class I, class X>
  BCType bctype(const I&, const X&) const {
    return ConvectionDiffusionBoundaryConditions::None;
  }

 private:
  RF scaling_;
};
}  // namespace detail

template <class GV, class FEM, class RF = double>
class L2 {
 public:
  using ParameterType = detail::L2Parameter<GV, RF>;
  static constexpr bool isLinear = true;
  // l2.hh:241-244.  The reference's intorderadd only raises the quadrature order of an integrand that the k+1
  // Gauss points of this library already integrate exactly (degree 2k per direction), so it is accepted and
  // not used: same mass matrix to rounding.
  explicit L2(int /*intorderadd*/ = 0, double scaling = 1.0) : param_(scaling), intorderadd(0) {}
  ParameterType& parameters() const { return param_; }
  void setTime(double) {}
  mutable ParameterType param_;
  int intorderadd;
  ConvectionDiffusionDGMethod::Type method = ConvectionDiffusionDGMethod::SIPG;
  ConvectionDiffusionDGWeights::Type weights = ConvectionDiffusionDGWeights::weightsOn;
  double alpha = 0.0;
};

// ConvectionDiffusionBoundaryConditionAdapter (convectiondiffusionparameter.hh:217-248)
template <class Param>
struct ConvectionDiffusionBoundaryConditionAdapter {
  explicit ConvectionDiffusionBoundaryConditionAdapter(const Param& p) : param(&p) {}
  template <class I, class X>
  bool isDirichlet(const I& is, const X& x) const {
    return param->bctype(is, x) == ConvectionDiffusionBoundaryConditions::Dirichlet;
  }
  const Param* param;
};
// ConvectionDiffusionDirichletExtensionAdapter (convectiondiffusionparameter.hh:255-300): u = g
template <class Param>
struct ConvectionDiffusionDirichletExtensionAdapter {
  template <class GV>
  ConvectionDiffusionDirichletExtensionAdapter(const GV&, Param& p) : param(&p) {}
  template <class E, class X>
  double evaluate(const E& e, const X& x) const {
    return param->g(e, x);
  }
  void setTime(double t) { param->setTime(t); }  // convectiondiffusionparameter.hh:291-294
  Param* param;
};

// ---- sampling of the parameter call-backs into the arrays of the C ABI ---------------------------
namespace detail {

template <int dim>
struct Sampled {
  int a_mode = PDB200_A_IDENTITY;
  int pointwise = 0;  // PDB200_POINTWISE_* bits: which of A, b, c, bctype are in the point-wise layout
  std::vector<double> A, b, c, f, g, j, o;
  std::vector<std::int8_t> bctype;
  bool has_b = false, has_c = false, has_f = false, has_g = false, has_j = false, has_o = false;
};

// The call-backs are sampled exactly where the reference evaluates them (include/pdelab_b200.h, layout (2)): A, b and
// c at every volume quadrature point (convectiondiffusiondg.hh:143-146,178,181; convectiondiffusionfem.hh:97-100,
// 127-129), A and b at the quadrature points of every face in the local coordinates of the cell (:370-371,426,755,
// 797), the QkDG boundary type per face quadrature point (:763).  A field that turns out constant on every cell
// (every boundary face) is handed over in the cell-wise layout — the same numbers, and the Kronecker kernels apply.
// A is only sampled per point when permeabilityIsConstantPerCell() is false, like the reference (:127,143).
template <class GV, class Param>
Sampled<GV::dimension> sample_parameters(const GV& gv, Param& param, int degree, int intorderadd, bool dg) {
  constexpr int dim = GV::dimension;
  Sampled<dim> S;
  const bool a_per_cell = param.permeabilityIsConstantPerCell();
  const int m = (2 * degree + intorderadd) / 2 + 1;  // convectiondiffusiondg.hh:139
  std::vector<double> xq(m), wq(m);
  check(pdb200_gauss_legendre(m, xq.data(), wq.data()), "pdb200_gauss_legendre");
  const long long ncells = gv.size(0);
  int nq = 1, nfq = 1;
  for (int d = 0; d < dim; d++) nq *= m;
  for (int d = 1; d < dim; d++) nfq *= m;
  const int NP = nq + 2 * dim * nfq;
  // local coordinates of the NP sample points of a cell
  std::vector<FieldVector<double, dim>> xp(NP);
  for (int q = 0; q < nq; q++) {
    int r = q;
    for (int d = 0; d < dim; d++) {
      xp[q][d] = xq[r % m];
      r /= m;
    }
  }
  for (int dir = 0; dir < dim; dir++)
    for (int side = 0; side < 2; side++)
      for (int q = 0; q < nfq; q++) {
        auto& x = xp[nq + (2 * dir + side) * nfq + q];
        int r = q;
        for (int d = 0; d < dim; d++) {
          if (d == dir) {
            x[d] = side;
          } else {
            x[d] = xq[r % m];
            r /= m;
          }
        }
      }
  const int na = a_per_cell ? 1 : NP;
  std::vector<double> Afull((std::size_t)ncells * na * dim * dim), bfull((std::size_t)ncells * NP * dim),
      cfull((std::size_t)ncells * nq);
  S.f.resize(ncells * nq);
  bool diag = true, scalar = true, ident = true, a_const = true, b_const = true, c_const = true;
  const FieldVector<double, dim> centre(0.5);
  for (long long e = 0; e < ncells; e++) {
    const auto cell = gv.cell(e);
    for (int pt = 0; pt < na; pt++) {
      const auto A = a_per_cell ? param.A(cell, centre) : param.A(cell, xp[pt]);
      double* out = &Afull[((std::size_t)e * na + pt) * dim * dim];
      for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++) {
          out[i * dim + j] = A[i][j];
          if (i != j && A[i][j] != 0.0) diag = false;
          if (i == j && A[i][i] != A[0][0]) scalar = false;
          if (i == j && A[i][i] != 1.0) ident = false;
          if (pt > 0 && out[i * dim + j] != Afull[(std::size_t)e * na * dim * dim + i * dim + j]) a_const = false;
        }
    }
    for (int pt = 0; pt < NP; pt++) {
      const auto b = param.b(cell, xp[pt]);
      for (int i = 0; i < dim; i++) {
        bfull[((std::size_t)e * NP + pt) * dim + i] = b[i];
        S.has_b |= b[i] != 0.0;
        if (b[i] != bfull[(std::size_t)e * NP * dim + i]) b_const = false;
      }
    }
    for (int q = 0; q < nq; q++) {
      const double c = param.c(cell, xp[q]);
      cfull[(std::size_t)e * nq + q] = c;
      S.has_c |= c != 0.0;
      if (c != cfull[(std::size_t)e * nq]) c_const = false;
      const double f = param.f(cell, xp[q]);
      S.f[e * nq + q] = f;
      S.has_f |= f != 0.0;
    }
  }
  // compress A to the cheapest layout that represents it exactly (selects the Kronecker kernel)
  const bool a_pw = !a_per_cell && !a_const;
  const int nae = a_pw ? NP : 1;  // entries per cell that are kept
  {
    const std::size_t ne = (std::size_t)ncells * nae;
    auto src = [&](std::size_t k) { return &Afull[(a_pw ? k : k * na) * dim * dim]; };  // cell-wise: the first sample
    if (diag && ident) {
      S.a_mode = PDB200_A_IDENTITY;
    } else if (diag && scalar) {
      S.a_mode = PDB200_A_SCALAR;
      S.A.resize(ne);
      for (std::size_t k = 0; k < ne; k++) S.A[k] = src(k)[0];
    } else if (diag) {
      S.a_mode = PDB200_A_DIAGONAL;
      S.A.resize(ne * dim);
      for (std::size_t k = 0; k < ne; k++)
        for (int i = 0; i < dim; i++) S.A[k * dim + i] = src(k)[i * dim + i];
    } else {
      S.a_mode = PDB200_A_FULL;
      S.A.resize(ne * dim * dim);
      for (std::size_t k = 0; k < ne; k++)
        for (int i = 0; i < dim * dim; i++) S.A[k * dim * dim + i] = src(k)[i];
    }
    if (a_pw && S.a_mode != PDB200_A_IDENTITY) S.pointwise |= PDB200_POINTWISE_A;
  }
  if (b_const) {
    S.b.resize((std::size_t)ncells * dim);
    for (long long e = 0; e < ncells; e++)
      for (int i = 0; i < dim; i++) S.b[e * dim + i] = bfull[(std::size_t)e * NP * dim + i];
  } else {
    S.b.swap(bfull);
    S.pointwise |= PDB200_POINTWISE_B;
  }
  if (c_const) {
    S.c.resize(ncells);
    for (long long e = 0; e < ncells; e++) S.c[e] = cfull[(std::size_t)e * nq];
  } else {
    S.c.swap(cfull);
    S.pointwise |= PDB200_POINTWISE_C;
  }
  // boundary faces in the numbering of pdelab_b200.h
  long long nbf = 0;
  for (int d = 0; d < dim; d++) nbf += 2 * (ncells / gv.grid().cells()[d]);
  S.bctype.resize(nbf);
  std::vector<std::int8_t> bcq(dg ? nbf * nfq : 0);  // QkDG: the type at every face quadrature point (:763)
  bool bc_const = true;
  S.g.assign(nbf * nfq, 0.0);
  S.j.assign(nbf * nfq, 0.0);
  S.o.assign(nbf * nfq, 0.0);
  long long bf = 0;
  const FieldVector<double, dim - 1> fcentre(0.5);
  for (int d = 0; d < dim; d++)
    for (int side = 0; side < 2; side++) {
      const long long nt = ncells / gv.grid().cells()[d];
      for (long long t = 0; t < nt; t++, bf++) {
        const auto is = gv.boundaryFace(d, side, t);
        // ConvectionDiffusionFEM and the conforming constraints: the type at the face centre
        // (convectiondiffusionfem.hh:226-229, constraints/conforming.hh:63-70)
        S.bctype[bf] = (std::int8_t)param.bctype(is, fcentre);
        for (int q = 0; q < nfq; q++) {
          FieldVector<double, dim - 1> xf;
          int r = q;
          for (int i = 0; i < dim - 1; i++) {
            xf[i] = xq[r % m];
            r /= m;
          }
          auto bc = param.bctype(is, fcentre);
          if (dg) {
            bc = param.bctype(is, xf);
            bcq[bf * nfq + q] = (std::int8_t)bc;
            if (q == 0) S.bctype[bf] = (std::int8_t)bc;  // the cell-wise layout, used if the type is constant on every face
            if (bcq[bf * nfq + q] != bcq[bf * nfq]) bc_const = false;
          }
          const auto xin = is.geometryInInside().global(xf);
          if (bc == ConvectionDiffusionBoundaryConditions::Dirichlet) {
            S.g[bf * nfq + q] = param.g(is.inside(), xin);  // convectiondiffusiondg.hh:842
            S.has_g |= S.g[bf * nfq + q] != 0.0;
          } else if (bc == ConvectionDiffusionBoundaryConditions::Neumann) {
            S.j[bf * nfq + q] = param.j(is, xf);  // :778
            S.has_j |= S.j[bf * nfq + q] != 0.0;
          } else if (bc == ConvectionDiffusionBoundaryConditions::Outflow) {
            S.o[bf * nfq + q] = param.o(is, xf);  // :813
            S.has_o |= S.o[bf * nfq + q] != 0.0;
          }
        }
      }
    }
  if (dg && !bc_const) {
    S.bctype.swap(bcq);
    S.pointwise |= PDB200_POINTWISE_BCTYPE;
  }
  return S;
}

}  // namespace detail

// ---- the assembled Jacobian container -----------------------------------------------------------
template <class GO>
class BCRSMatrixContainer;

// ---- GridOperator -------------------------------------------------------------------------------
template <class GFSU, class GFSV, class LOP, class MB, class DF, class RF, class JF,
          class CU = typename GFSU::template ConstraintsContainer<RF>::Type,
          class CV = typename GFSV::template ConstraintsContainer<RF>::Type>
class GridOperator {
 public:
  using GV = typename GFSU::GridView;
  static constexpr int dim = GV::dimension;
  using FEM = typename GFSU::FiniteElementMapType;
  // GridOperatorTraits (gridoperator/common/gridoperatorutilities.hh:31-80)
  struct Traits {
    using TrialGridFunctionSpace = GFSU;
    using TestGridFunctionSpace = GFSV;
    using TrialGridFunctionSpaceConstraints = CU;
    using TestGridFunctionSpaceConstraints = CV;
    using MatrixBackend = MB;
    using DomainField = DF;
    using RangeField = RF;
    using JacobianField = JF;
    using Domain = Backend::Vector<GFSU, DF>;
    using Range = Backend::Vector<GFSV, RF>;
    using Jacobian = BCRSMatrixContainer<GridOperator>;
    using LocalOperator = LOP;
  };
  using Domain = typename Traits::Domain;
  using Range = typename Traits::Range;
  using Jacobian = typename Traits::Jacobian;
  struct Pattern {  // scalar CSR, columns ascending (bcrsmatrixbackend.hh:90-121)
    std::vector<std::uint64_t> rowptr, colidx;
  };
  template <class T>
  struct MatrixContainer {
    using Type = Jacobian;
  };

  // local assembler facade: what OnTheFlyOperator / StationaryLinearProblemSolver touch
  // (gridoperator/default/localassembler.hh)
  struct LocalAssembler {
    LOP* lop;
    const CU* cu;
    const CV* cv;
    LOP& localOperator() const { return *lop; }
    const CU& trialConstraints() const { return *cu; }
    const CV& testConstraints() const { return *cv; }
    static constexpr bool isLinear() { return true; }
    GridOperator* go = nullptr;
    void setTime(double t) { go->setTime(t); }  // localassembler.hh: lop.setTime(t); the call-backs are re-sampled
    void setWeight(double w) {
      if (w != 1.0) throw Exception("pdelab_b200: engine weights other than 1 are not supported");
    }
    // default/localassembler.hh:259-283: whether this operator's engines run the pre- / post-processing steps of a
    // joint assembly (set by setupGridOperators).  The flags are kept for interface parity: the device path of a joint
    // assembly is the fused stage operator of OneStepGridOperator, which applies the constraints once, at the end.
    bool doPreProcessing() const { return pre_; }
    void preProcessing(bool v) { pre_ = v; }
    bool doPostProcessing() const { return post_; }
    void postProcessing(bool v) { post_ = v; }
    bool pre_ = true, post_ = true;
  };
  // global assembler facade (gridoperator/default/assembler.hh:35-83): what callers reach through go.assembler()
  struct Assembler {
    const GFSU* gfsu;
    const GFSV* gfsv;
    const GFSU& trialGridFunctionSpace() const { return *gfsu; }
    const GFSV& testGridFunctionSpace() const { return *gfsv; }
  };

  // gridoperator.hh:76-82
  GridOperator(const GFSU& gfsu, const CU& cu, const GFSV& gfsv, const CV& cv, LOP& lop, const MB& mb = MB())
      : gfsu_(gfsu), gfsv_(gfsv), lop_(lop), mb_(mb), cu_(&cu), cv_(&cv), la_{&lop, &cu, &cv}, as_{&gfsu, &gfsv} {
    la_.go = this;
    init();
  }
  // gridoperator.hh:85-89 (empty constraints)
  GridOperator(const GFSU& gfsu, const GFSV& gfsv, LOP& lop, const MB& mb = MB())
      : gfsu_(gfsu), gfsv_(gfsv), lop_(lop), mb_(mb), cu_(&empty_cu_), cv_(&empty_cv_), la_{&lop, &empty_cu_, &empty_cv_},
        as_{&gfsu, &gfsv} {
    la_.go = this;
    init();
  }
  GridOperator(const GridOperator&) = delete;
  GridOperator& operator=(const GridOperator&) = delete;
  ~GridOperator() {
    if (h_) pdb200_destroy(h_);
  }

  const GFSU& trialGridFunctionSpace() const { return gfsu_; }
  const GFSV& testGridFunctionSpace() const { return gfsv_; }
  typename GFSU::SizeType globalSizeU() const { return num_dofs(); }  // gridoperator.hh:104-107
  typename GFSV::SizeType globalSizeV() const { return num_dofs(); }
  LocalAssembler& localAssembler() const { return la_; }
  Assembler& assembler() { return as_; }  // gridoperator.hh:115-117
  const Assembler& assembler() const { return as_; }
  const MB& matrixBackend() const { return mb_; }
  // gridoperator.hh:122-151: operators of a joint assembly — the first one pre-processes, the last one post-processes
  template <class GridOperatorTuple>
  static void setupGridOperators(GridOperatorTuple tuple) {
    constexpr std::size_t size = std::tuple_size<GridOperatorTuple>::value;
    std::size_t index = 0;
    std::apply(
        [&](auto&... go) {
          ((go.localAssembler().preProcessing(index == 0), go.localAssembler().postProcessing(index == size - 1), ++index), ...);
        },
        tuple);
  }
  pdb200_handle handle() const { return h_; }  // for stream control / device-pointer calls

  // re-sample the parameter call-backs (the reference re-evaluates them on every assembly)
  void update() {
    if (h_) pdb200_destroy(h_);
    h_ = nullptr;
    init();
  }

  // LocalAssembler::setTime -> lop.setTime(t) -> param.setTime(t) (gridoperator/default/localassembler.hh): the
  // reference re-evaluates the call-backs on every assembly; here they are re-sampled when the time changes and
  // the device arrays are refreshed in place.  setTimeDependent(false) skips the re-sampling.
  void setTimeDependent(bool v) { time_dependent_ = v; }
  void setTime(double t) {
    lop_.setTime(t);
    if (!time_dependent_ || (time_set_ && t == time_)) return;
    time_ = t;
    time_set_ = true;
    auto S = detail::sample_parameters(gfsu_.gridView(), lop_.parameters(), FEM::degree, lop_.intorderadd,
                                       FEM::space == PDB200_SPACE_QKDG);
    const bool same_layout = S.a_mode == S_.a_mode && S.pointwise == S_.pointwise && S.has_b == S_.has_b && S.has_c == S_.has_c && S.has_f == S_.has_f &&
                             S.has_g == S_.has_g && S.has_j == S_.has_j && S.has_o == S_.has_o && S.bctype == S_.bctype;
    if (!same_layout) {  // a field switched on or off, or the boundary types moved: new operator (new constraint set)
      update();
      return;
    }
    S_ = std::move(S);
    pdb200_problem q = p_;
    q.A = S_.A.empty() ? nullptr : S_.A.data();
    q.b = S_.has_b ? S_.b.data() : nullptr;
    q.c = S_.has_c ? S_.c.data() : nullptr;
    q.f = S_.has_f ? S_.f.data() : nullptr;
    q.bctype = nullptr;
    q.g = S_.has_g ? S_.g.data() : nullptr;
    q.j = S_.has_j ? S_.j.data() : nullptr;
    q.o = S_.has_o ? S_.o.data() : nullptr;
    check(pdb200_update_coefficients(h_, &q), "setTime");
    p_.A = q.A, p_.b = q.b, p_.c = q.c, p_.f = q.f, p_.g = q.g, p_.j = q.j, p_.o = q.o;
    p_.bctype = S_.bctype.data();
  }
  // StationaryLinearProblemSolver::apply as one device-resident call (dispatch point shared with OneStepGridOperator)
  void solveStationary(int solver, int precond, bool matrix_free, double* x, double reduction, double min_defect,
                       unsigned maxiter, pdb200_solve_result* s) const {
    check(pdb200_solve_stationary(h_, solver, precond, matrix_free ? 1 : 0, x, reduction, min_defect, maxiter, s),
          "StationaryLinearProblemSolver::apply");
  }

  // gridoperator.hh:168-173
  void fill_pattern(Pattern& p) const {
    std::uint64_t nr = 0, nnz = 0;
    check(pdb200_pattern_size(h_, &nr, &nnz), "fill_pattern");
    p.rowptr.assign(nr + 1, 0);
    p.colidx.assign(nnz, 0);
    check(pdb200_pattern(h_, p.rowptr.data(), p.colidx.data()), "fill_pattern");
  }
  // gridoperator.hh:176-181:  r += R(x)
  void residual(const Domain& x, Range& r) const { check(pdb200_residual(h_, x.data(), r.data()), "residual"); }
  // gridoperator.hh:184-189:  A += dR/dx
  void jacobian(const Domain& x, Jacobian& a) const {
    check(pdb200_jacobian(h_, x.data(), a.values().data(), PDB200_LAYOUT_CSR), "jacobian");
  }
  // gridoperator.hh:192-197:  y += J z (linear problems)
  void jacobian_apply(const Domain& z, Range& y) const {
    check(pdb200_jacobian_apply(h_, z.data(), y.data()), "jacobian_apply");
  }
  // gridoperator.hh:200-205: throws for a linear local operator, like the reference
  void jacobian_apply(const Domain& u, const Domain& z, Range& y) const {
    check(pdb200_jacobian_apply_nonlinear(h_, u.data(), z.data(), y.data()), "jacobian_apply");
  }
  // y = J x with the zeroing fused (OnTheFlyOperator::apply); raw pointers may be device memory
  void onthefly_apply(const double* x, double* y) const { check(pdb200_onthefly_apply(h_, x, y), "apply"); }
  void residual(const double* x, double* r) const { check(pdb200_residual(h_, x, r), "residual"); }
  void jacobian_apply(const double* z, double* y) const { check(pdb200_jacobian_apply(h_, z, y), "jacobian_apply"); }

  void make_consistent(Jacobian&) const {}  // sequential / overlapping: nothing to add (gridoperator.hh:207-212)

  // gridoperator.hh:144-165: interpolate f into x on the constrained DOFs' space (Lagrange nodes),
  // then copy the unconstrained entries of xold
  template <class F>
  void interpolate(const Domain& xold, F& f, Domain& x) const {
    B200_interpolate(f, x);
    const auto con = constrained();
    std::vector<char> is_con(x.N(), 0);
    for (auto i : con) is_con[i] = 1;
    for (std::size_t i = 0; i < x.N(); i++)
      if (!is_con[i]) x[i] = xold[i];
  }

  // Lagrange interpolation of f (evaluate(cell, xlocal)) at the nodes j/k of every cell
  template <class F>
  void B200_interpolate(F& f, Domain& x) const {
    if (FEM::basis != PDB200_BASIS_LAGRANGE)
      throw Exception("interpolate: nodal interpolation at j/k is the Lagrange basis' (use an L2 projection for the "
                      "Legendre / Gauss-Lobatto QkDG bases)");
    const auto& gv = gfsu_.gridView();
    constexpr int k = FEM::degree;
    const int n = (int)FEM::maxLocalSize();
    std::vector<std::uint64_t> idx(n);
    for (long long e = 0; e < gv.size(0); e++) {
      const auto cell = gv.cell(e);
      check(pdb200_cell_dof_indices(h_, (std::uint64_t)e, idx.data()), "interpolate");
      for (int i = 0; i < n; i++) {
        FieldVector<double, dim> xl;
        int r = i;
        for (int d = 0; d < dim; d++) {
          xl[d] = double(r % (k + 1)) / k;
          r /= k + 1;
        }
        x[idx[i]] = f.evaluate(cell, xl);
      }
    }
  }

  std::vector<std::uint64_t> constrained() const {
    std::uint64_t n = 0;
    check(pdb200_constrained_dofs(h_, &n, nullptr), "constraints");
    std::vector<std::uint64_t> idx(n);
    if (n) check(pdb200_constrained_dofs(h_, &n, idx.data()), "constraints");
    return idx;
  }
  std::string lastKernel() const { return pdb200_last_kernel(h_); }
  // the problem description handed to the C ABI (tests hand the same struct to the CPU oracle)
  const pdb200_problem& problem() const { return p_; }

 private:
  std::size_t num_dofs() const {
    std::uint64_t n = 0;
    check(pdb200_num_dofs(h_, &n), "globalSize");
    return (std::size_t)n;
  }
  void init() {
    static_assert(std::is_same<GFSU, GFSV>::value, "Galerkin: trial and test space coincide on this path");
    const auto& gv = gfsu_.gridView();
    S_ = detail::sample_parameters(gv, lop_.parameters(), FEM::degree, lop_.intorderadd, FEM::space == PDB200_SPACE_QKDG);
    auto& S = S_;
    pdb200_problem& p = p_;
    p = pdb200_problem{};
    p.dim = dim;
    for (int d = 0; d < 3; d++) {
      p.cells[d] = d < dim ? gv.grid().cells()[d] : 1;
      p.lower[d] = d < dim ? gv.grid().lower()[d] : 0.0;
      p.upper[d] = d < dim ? gv.grid().upper()[d] : 1.0;
      for (int s = 0; s < 2; s++) p.side_kind[d][s] = gv.grid().side_kind[d][s];
    }
    p.space = FEM::space;
    p.degree = FEM::degree;
    p.basis = FEM::basis;
    p.dg_method = lop_.method == ConvectionDiffusionDGMethod::SIPG   ? PDB200_DG_SIPG
                  : lop_.method == ConvectionDiffusionDGMethod::NIPG ? PDB200_DG_NIPG
                                                                     : PDB200_DG_IIPG;
    p.dg_weights = lop_.weights == ConvectionDiffusionDGWeights::weightsOn ? PDB200_DG_WEIGHTS_ON : PDB200_DG_WEIGHTS_OFF;
    p.dg_alpha = lop_.alpha;
    p.intorderadd = lop_.intorderadd;
    p.a_mode = S.a_mode;
    p.pointwise = S.pointwise;
    p.A = S.A.empty() ? nullptr : S.A.data();
    p.b = S.has_b ? S.b.data() : nullptr;
    p.c = S.has_c ? S.c.data() : nullptr;
    p.f = S.has_f ? S.f.data() : nullptr;
    p.bctype = S.bctype.data();
    p.g = S.has_g ? S.g.data() : nullptr;
    p.j = S.has_j ? S.j.data() : nullptr;
    p.o = S.has_o ? S.o.data() : nullptr;
    p.device = 0;
    p.kernel = PDB200_KERNEL_AUTO;
    check(pdb200_create(&p, &h_), "GridOperator");
  }

  const GFSU& gfsu_;
  const GFSV& gfsv_;
  LOP& lop_;
  MB mb_;
  CU empty_cu_{};
  CV empty_cv_{};
  const CU* cu_;
  const CV* cv_;
  mutable LocalAssembler la_;
  Assembler as_;
  detail::Sampled<dim> S_;  // sampled call-backs (kept alive: p_ points into them)
  pdb200_problem p_{};
  pdb200_handle h_ = nullptr;
  bool time_dependent_ = true, time_set_ = false;
  double time_ = 0.0;
};

// FastDGGridOperator (gridoperator/fastdg.hh:37-230): the reference's assembler variant that skips the
// LFSIndexCache for DG spaces with Blocking::fixed and aliases vector blocks.  Every kernel of this library
// already addresses DG vectors as cell * n + i, so it is the same operator: same interface, same results.
template <class GFSU, class GFSV, class LOP, class MB, class DF, class RF, class JF,
          class CU = typename GFSU::template ConstraintsContainer<RF>::Type,
          class CV = typename GFSV::template ConstraintsContainer<RF>::Type>
using FastDGGridOperator = GridOperator<GFSU, GFSV, LOP, MB, DF, RF, JF, CU, CV>;

// Backend::Matrix / ISTL::BCRSMatrixContainer constructed from the grid operator
// (backend/istl/bcrsmatrix.hh:78-82 -> MB::buildPattern -> go.fill_pattern)
template <class GO>
class BCRSMatrixContainer {
 public:
  using ElementType = double;
  template <class AnyGO>  // GridOperator or OneStepGridOperator over it (bcrsmatrix.hh:78-82)
  explicit BCRSMatrixContainer(const AnyGO& go) {
    go.fill_pattern(p_);
    v_.assign(p_.colidx.size(), 0.0);
  }
  BCRSMatrixContainer& operator=(double s) {
    std::fill(v_.begin(), v_.end(), s);
    return *this;
  }
  std::size_t N() const { return p_.rowptr.size() - 1; }
  std::size_t M() const { return N(); }
  std::size_t nonzeroes() const { return v_.size(); }
  const std::vector<std::uint64_t>& rowptr() const { return p_.rowptr; }
  const std::vector<std::uint64_t>& colidx() const { return p_.colidx; }
  std::vector<double>& values() { return v_; }
  const std::vector<double>& values() const { return v_; }
  // entry access like BCRSMatrix::operator()(ri, ci) (backend/istl/bcrsmatrix.hh:212-215)
  double operator()(std::size_t i, std::size_t j) const {
    auto b = p_.colidx.begin() + p_.rowptr[i], e = p_.colidx.begin() + p_.rowptr[i + 1];
    auto it = std::lower_bound(b, e, (std::uint64_t)j);
    if (it == e || *it != j) throw Exception("BCRSMatrix: entry not in pattern");
    return v_[it - p_.colidx.begin()];
  }
  // y = A x (dune-istl BCRSMatrix::mv), host loop: the container lives in host memory here
  template <class X, class Y>
  void mv(const X& x, Y& y) const {
    for (std::size_t i = 0; i < N(); i++) {
      double s = 0;
      for (std::uint64_t k = p_.rowptr[i]; k < p_.rowptr[i + 1]; k++) s += v_[k] * x[p_.colidx[k]];
      y[i] = s;
    }
  }

 private:
  typename GO::Pattern p_;
  std::vector<double> v_;
};

// OnTheFlyOperator (backend/istl/seqistlsolverbackend.hh:44-100): y = J x / y += alpha J x
template <class X, class Y, class GO>
class OnTheFlyOperator {
 public:
  using domain_type = X;
  using range_type = Y;
  using field_type = double;
  explicit OnTheFlyOperator(const GO& go) : go_(go) {}
  void apply(const X& x, Y& y) const { go_.onthefly_apply(x.data(), y.data()); }  // :66-76
  void applyscaleadd(field_type alpha, const X& x, Y& y) const {               // :78-90
    Y t(y.gridFunctionSpace(), 0.0);
    go_.jacobian_apply(x, t);
    y.axpy(alpha, t);
  }

 private:
  const GO& go_;
};

// ---- linear solver back-ends and the stationary problem solver -----------------------------------
// LinearSolverResult / LinearResultStorage (backend/solver.hh:28-75)
template <class RF>
struct LinearSolverResult {
  bool converged = false;
  unsigned int iterations = 0;
  double elapsed = 0.0;
  RF reduction = 0.0;
  RF conv_rate = 0.0;
  void clear() { *this = LinearSolverResult(); }
};

namespace detail {
// what every sequential back-end of this mirror is: a (solver, preconditioner) pair run on the device
template <int SOLVER, int PRECOND, bool MATRIX_FREE>
class DeviceBackend {
 public:
  static constexpr int solver = SOLVER, precond = PRECOND;
  static constexpr bool matrix_free = MATRIX_FREE;
  explicit DeviceBackend(unsigned maxiter = 5000, int verbose = 1) : maxiter_(maxiter), verbose_(verbose) {}
  // SequentialNorm (backend/istl/seqistlsolverbackend.hh / backend/solver.hh): two-norm
  template <class V>
  double norm(const V& v) const {
    return std::sqrt(v.dot(v));
  }
  const LinearSolverResult<double>& result() const { return res; }
  unsigned maxiter() const { return maxiter_; }
  void store(const pdb200_solve_result& r) {
    res.converged = r.converged != 0;
    res.iterations = r.iterations;
    res.elapsed = r.elapsed;
    res.reduction = r.reduction;
    res.conv_rate = r.conv_rate;
    if (verbose_ > 0)
      std::printf("=== device Krylov: %u iterations, reduction %.3e, %.4f s\n", r.iterations, r.reduction, r.elapsed);
  }

 protected:
  LinearSolverResult<double> res;
  unsigned maxiter_;
  int verbose_;
};
}  // namespace detail

namespace detail {
template <class GO, int SOLVER, int PRECOND>
class MatrixFreeBackend : public DeviceBackend<SOLVER, PRECOND, true> {
  using Base = DeviceBackend<SOLVER, PRECOND, true>;

 public:
  explicit MatrixFreeBackend(const GO& go, unsigned maxiter = 5000, int verbose = 1) : Base(maxiter, verbose), go_(go) {}
  // apply(z, r, reduction): solve J z = r, r := final defect (seqistlsolverbackend.hh:181-192)
  template <class V, class W>
  void apply(V& z, W& r, double reduction) {
    pdb200_solve_result s;
    check(pdb200_solve(go_.handle(), SOLVER, PRECOND, nullptr, PDB200_LAYOUT_CSR, z.data(), r.data(), reduction,
                       this->maxiter_, &s),
          "ISTLBackend_SEQ_MatrixFree::apply");
    this->store(s);
  }
  void setLinearizationPoint(const typename GO::Domain&) {}  // linear operators only

 private:
  const GO& go_;
};
}  // namespace detail
// ISTLBackend_SEQ_MatrixFree_BCGS_Richardson (seqistlsolverbackend.hh:157-203,1039-1050)
template <class GO>
using ISTLBackend_SEQ_MatrixFree_BCGS_Richardson = detail::MatrixFreeBackend<GO, PDB200_SOLVER_BICGSTAB, PDB200_PRECOND_NONE>;
template <class GO>
using ISTLBackend_SEQ_MatrixFree_CG_Richardson = detail::MatrixFreeBackend<GO, PDB200_SOLVER_CG, PDB200_PRECOND_NONE>;
// ISTLBackend_SEQ_MatrixFree_Base<GO, PrecGO, Solver> (backend/istl/matrixfree/backends.hh:62-143) with PrecGO built
// on AssembledBlockJacobiPreconditionerLocalOperator (assembledblockjacobipreconditioner.hh:96-230): the
// preconditioner grid operator is not a separate object here, the library applies D^-1 matrix-free
template <class GO>
using ISTLBackend_SEQ_MatrixFree_BCGS_BlockJacobi = detail::MatrixFreeBackend<GO, PDB200_SOLVER_BICGSTAB, PDB200_PRECOND_BLOCK_JACOBI>;
template <class GO>
using ISTLBackend_SEQ_MatrixFree_CG_BlockJacobi = detail::MatrixFreeBackend<GO, PDB200_SOLVER_CG, PDB200_PRECOND_BLOCK_JACOBI>;
// the same back-end with PrecGO built on BlockSORPreconditionerLocalOperator (backend/istl/matrixfree/
// blocksorpreconditioner.hh:36-301; backends.hh:79-88 requires the FastDG grid operator for it): one matrix-free block
// SOR sweep in index-set order per application; CG gets the symmetric (forward + backward) sweep
template <class GO>
using ISTLBackend_SEQ_MatrixFree_BCGS_BlockSOR = detail::MatrixFreeBackend<GO, PDB200_SOLVER_BICGSTAB, PDB200_PRECOND_BLOCK_SOR>;
template <class GO>
using ISTLBackend_SEQ_MatrixFree_CG_BlockSSOR = detail::MatrixFreeBackend<GO, PDB200_SOLVER_CG, PDB200_PRECOND_BLOCK_SSOR>;

namespace detail {
template <class GO, int SOLVER, int PRECOND>
class AssembledBackend : public DeviceBackend<SOLVER, PRECOND, false> {
  using Base = DeviceBackend<SOLVER, PRECOND, false>;

 public:
  // the reference's assembled back-ends are constructed without the grid operator (:208-224); the
  // device path needs its handle, so it is passed here
  explicit AssembledBackend(const GO& go, unsigned maxiter = 5000, int verbose = 1) : Base(maxiter, verbose), go_(go) {}
  // apply(A, z, r, reduction) (:226-248)
  template <class M, class V, class W>
  void apply(M& A, V& z, W& r, double reduction) {
    pdb200_solve_result s;
    check(pdb200_solve(go_.handle(), SOLVER, PRECOND, A.values().data(), PDB200_LAYOUT_CSR, z.data(), r.data(),
                       reduction, this->maxiter_, &s),
          "ISTLBackend_SEQ::apply");
    this->store(s);
  }

 private:
  const GO& go_;
};
}  // namespace detail
// ISTLBackend_SEQ_BCGS_Jac / _CG_Jac / _BCGS_Richardson (seqistlsolverbackend.hh:385-416,538-553)
template <class GO>
using ISTLBackend_SEQ_BCGS_Jac = detail::AssembledBackend<GO, PDB200_SOLVER_BICGSTAB, PDB200_PRECOND_JACOBI>;
template <class GO>
using ISTLBackend_SEQ_CG_Jac = detail::AssembledBackend<GO, PDB200_SOLVER_CG, PDB200_PRECOND_JACOBI>;
template <class GO>
using ISTLBackend_SEQ_BCGS_Richardson = detail::AssembledBackend<GO, PDB200_SOLVER_BICGSTAB, PDB200_PRECOND_NONE>;

// StationaryLinearProblemSolverResult (stationary/linearproblem.hh:20-45)
template <class RF>
struct StationaryLinearProblemSolverResult : LinearSolverResult<RF> {
  RF first_defect = 0.0, defect = 0.0;
  double assembler_time = 0.0, linear_solver_time = 0.0;
  int linear_solver_iterations = 0;
};

// StationaryLinearProblemSolver (stationary/linearproblem.hh:57-320): apply() assembles (unless the
// back-end is matrix-free), evaluates r = R(x), solves J z = r to max(reduction, min_defect/|r|) and
// updates x -= z — as ONE device-resident call (pdb200_solve_stationary).
template <class GO, class LS, class V>
class StationaryLinearProblemSolver {
 public:
  using Result = StationaryLinearProblemSolverResult<double>;
  StationaryLinearProblemSolver(const GO& go, LS& ls, V& x, double reduction, double min_defect = 1e-99, int verbose = 1)
      : go_(go), ls_(ls), x_(&x), reduction_(reduction), min_defect_(min_defect), verbose_(verbose) {}
  // linearproblem.hh:106-118: the solution vector is handed to apply(x)
  StationaryLinearProblemSolver(const GO& go, LS& ls, double reduction, double min_defect = 1e-99, int verbose = 1)
      : go_(go), ls_(ls), x_(nullptr), reduction_(reduction), min_defect_(min_defect), verbose_(verbose) {}
  void apply(V& x, bool reuse_matrix = false) {  // linearproblem.hh:180-186
    x_ = &x;
    apply(reuse_matrix);
  }
  void apply(bool /*reuse_matrix*/ = false) {
    if (!x_) throw Exception("StationaryLinearProblemSolver: no solution vector");
    pdb200_solve_result s;
    go_.solveStationary(LS::solver, LS::precond, LS::matrix_free, x_->data(), reduction_, min_defect_, ls_.maxiter(), &s);
    ls_.store(s);
    static_cast<LinearSolverResult<double>&>(res_) = ls_.result();
    res_.first_defect = s.first_defect;
    res_.defect = s.defect;
    res_.linear_solver_time = s.elapsed;
    res_.linear_solver_iterations = (int)s.iterations;
  }
  const Result& result() const { return res_; }

 private:
  const GO& go_;
  LS& ls_;
  V* x_;
  double reduction_, min_defect_;
  int verbose_;
  Result res_;
};

// constraints(bctype, gfs, cc) (constraints/common/constraints.hh:588-687): the constrained set is
// derived by the library from the per-face boundary types; it needs an operator handle, so the
// grid operator offers it — this overload fills cc from a grid operator built on the same spaces.
template <class GO, class CC>
void constraints(const GO& go, CC& cc) {
  cc.dofs = go.constrained();
}
// set_nonconstrained_dofs(cc, v, x) (constraints/common/constraints.hh:796-802)
template <class CC, class V>
void set_nonconstrained_dofs(const CC& cc, double v, V& x) {
  std::vector<char> is_con(x.N(), 0);
  for (auto i : cc.dofs) is_con[i] = 1;
  for (std::size_t i = 0; i < x.N(); i++)
    if (!is_con[i]) x[i] = v;
}

}  // namespace B200
}  // namespace PDELab
}  // namespace Dune

#endif  // PDELAB_B200_HOST_GRIDOPERATOR_HH

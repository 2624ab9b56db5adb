// onestep.hh — C++ host mirror of PDELab's instationary layer over the C ABI (pdb200_onestep_*):
//   TimeSteppingParameterInterface and the method tables   dune/pdelab/instationary/onestepparameter.hh:43-698
//   OneStepGridOperator<GO0, GO1, implicit>                 dune/pdelab/gridoperator/onestep.hh:30-308
//   OneStepMethod<T, IGOS, PDESOLVER, TrlV, TstV>           dune/pdelab/instationary/implicitonestep.hh:37-439
// Same class names, template parameters, member names and argument meaning as the reference, in namespace
// Dune::PDELab::B200.  A stage is one fused operator on the device (csrc/onestep.cu); the linear solver
// back-ends and StationaryLinearProblemSolver of gridoperator.hh work on a OneStepGridOperator unchanged.
#ifndef PDELAB_B200_HOST_ONESTEP_HH
#define PDELAB_B200_HOST_ONESTEP_HH

#include <iomanip>
#include <iostream>
#include <memory>

#include "gridoperator.hh"

namespace Dune {
namespace PDELab {
namespace B200 {

// ---- instationary/onestepparameter.hh ----------------------------------------------------------------
template <class R>
class TimeSteppingParameterInterface {
 public:
  using RealType = R;
  virtual bool implicit() const = 0;
  virtual unsigned s() const = 0;
  virtual R a(int r, int i) const = 0;  // r in 1..s, i in 0..r
  virtual R b(int r, int i) const = 0;
  virtual R d(int r) const = 0;
  virtual std::string name() const = 0;
  virtual ~TimeSteppingParameterInterface() {}
};

namespace detail {
// common storage of the tables: S stages, (S+1) columns
template <class R, int S>
class TableParameter : public TimeSteppingParameterInterface<R> {
 public:
  unsigned s() const override { return S; }
  R a(int r, int i) const override { return A[r - 1][i]; }
  R b(int r, int i) const override { return B[r - 1][i]; }
  R d(int i) const override { return D[i]; }

 protected:
  R D[S + 1] = {};
  R A[S][S + 1] = {};
  R B[S][S + 1] = {};
};
}  // namespace detail

// onestepparameter.hh:88-151
template <class R>
class OneStepThetaParameter : public detail::TableParameter<R, 1> {
 public:
  explicit OneStepThetaParameter(R theta_) : theta(theta_) {
    this->D[0] = 0.0, this->D[1] = 1.0;
    this->A[0][0] = -1.0, this->A[0][1] = 1.0;
    this->B[0][0] = 1.0 - theta, this->B[0][1] = theta;
  }
  bool implicit() const override { return theta > 0.0; }
  std::string name() const override { return "one step theta"; }

 private:
  R theta;
};
template <class R>
class ExplicitEulerParameter : public OneStepThetaParameter<R> {
 public:
  ExplicitEulerParameter() : OneStepThetaParameter<R>(0.0) {}
  std::string name() const override { return "explicit Euler"; }
};
template <class R>
class ImplicitEulerParameter : public OneStepThetaParameter<R> {
 public:
  ImplicitEulerParameter() : OneStepThetaParameter<R>(1.0) {}
  std::string name() const override { return "implicit Euler"; }
};
// onestepparameter.hh:213-280
template <class R>
class HeunParameter : public detail::TableParameter<R, 2> {
 public:
  HeunParameter() {
    const R D_[3] = {0.0, 1.0, 1.0}, A_[2][3] = {{-1.0, 1.0, 0.0}, {-0.5, -0.5, 1.0}}, B_[2][3] = {{1.0, 0.0, 0.0}, {0.0, 0.5, 0.0}};
    for (int i = 0; i < 3; i++) this->D[i] = D_[i];
    for (int r = 0; r < 2; r++)
      for (int i = 0; i < 3; i++) this->A[r][i] = A_[r][i], this->B[r][i] = B_[r][i];
  }
  bool implicit() const override { return false; }
  std::string name() const override { return "Heun"; }
};
// onestepparameter.hh:286-357
template <class R>
class Shu3Parameter : public detail::TableParameter<R, 3> {
 public:
  Shu3Parameter() {
    const R D_[4] = {0.0, 1.0, 0.5, 1.0};
    const R A_[3][4] = {{-1.0, 1.0, 0.0, 0.0}, {-0.75, -0.25, 1.0, 0.0}, {-1.0 / 3.0, 0.0, -2.0 / 3.0, 1.0}};
    const R B_[3][4] = {{1.0, 0.0, 0.0, 0.0}, {0.0, 0.25, 0.0, 0.0}, {0.0, 0.0, 2.0 / 3.0, 0.0}};
    for (int i = 0; i < 4; i++) this->D[i] = D_[i];
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < 4; i++) this->A[r][i] = A_[r][i], this->B[r][i] = B_[r][i];
  }
  bool implicit() const override { return false; }
  std::string name() const override { return "Shu's third order method"; }
};
// onestepparameter.hh:363-436
template <class R>
class RK4Parameter : public detail::TableParameter<R, 4> {
 public:
  RK4Parameter() {
    const R D_[5] = {0.0, 0.5, 0.5, 1.0, 1.0};
    const R B_[4][5] = {{0.5, 0, 0, 0, 0}, {0, 0.5, 0, 0, 0}, {0, 0, 1.0, 0, 0}, {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0, 0}};
    for (int i = 0; i < 5; i++) this->D[i] = D_[i];
    for (int r = 0; r < 4; r++) {
      this->A[r][0] = -1.0, this->A[r][r + 1] = 1.0;
      for (int i = 0; i < 5; i++) this->B[r][i] = B_[r][i];
    }
  }
  bool implicit() const override { return false; }
  std::string name() const override { return "RK4"; }
};
// onestepparameter.hh:444-510
template <class R>
class Alexander2Parameter : public detail::TableParameter<R, 2> {
 public:
  Alexander2Parameter() {
    const R alpha = 1.0 - 0.5 * std::sqrt(2.0);
    this->D[0] = 0.0, this->D[1] = alpha, this->D[2] = 1.0;
    this->A[0][0] = -1.0, this->A[0][1] = 1.0;
    this->A[1][0] = -1.0, this->A[1][2] = 1.0;
    this->B[0][1] = alpha;
    this->B[1][1] = 1.0 - alpha, this->B[1][2] = alpha;
  }
  bool implicit() const override { return true; }
  std::string name() const override { return "Alexander (order 2)"; }
};
// onestepparameter.hh:521-598
template <class R>
class FractionalStepParameter : public detail::TableParameter<R, 3> {
 public:
  FractionalStepParameter() {
    const R theta = 1.0 - 0.5 * std::sqrt(2.0), thetap = 1.0 - 2.0 * theta, alpha = 2.0 - std::sqrt(2.0), beta = 1.0 - alpha;
    this->D[0] = 0.0, this->D[1] = theta, this->D[2] = 1.0 - theta, this->D[3] = 1.0;
    this->A[0][0] = -1.0, this->A[0][1] = 1.0;
    this->A[1][1] = -1.0, this->A[1][2] = 1.0;
    this->A[2][2] = -1.0, this->A[2][3] = 1.0;
    this->B[0][0] = beta * theta, this->B[0][1] = alpha * theta;
    this->B[1][1] = alpha * thetap, this->B[1][2] = alpha * theta;
    this->B[2][2] = beta * theta, this->B[2][3] = alpha * theta;
  }
  bool implicit() const override { return true; }
  std::string name() const override { return "Fractional step theta"; }
};
// onestepparameter.hh:604-698
template <class R>
class Alexander3Parameter : public detail::TableParameter<R, 3> {
 public:
  Alexander3Parameter() {
    R alpha = 0.4358665215;
    for (int i = 1; i <= 10; i++)  // Newton iteration for alpha (:613-621)
      alpha = alpha - (alpha * (alpha * alpha - 3.0 * (alpha - 0.5)) - 1.0 / 6.0) / (3.0 * alpha * (alpha - 2.0) + 1.5);
    const R tau2 = (1.0 + alpha) * 0.5;
    const R b1 = -(6.0 * alpha * alpha - 16.0 * alpha + 1.0) * 0.25;
    const R b2 = (6 * alpha * alpha - 20.0 * alpha + 5.0) * 0.25;
    this->D[0] = 0.0, this->D[1] = alpha, this->D[2] = tau2, this->D[3] = 1.0;
    for (int r = 0; r < 3; r++) this->A[r][0] = -1.0, this->A[r][r + 1] = 1.0;
    this->B[0][1] = alpha;
    this->B[1][1] = tau2 - alpha, this->B[1][2] = alpha;
    this->B[2][1] = b1, this->B[2][2] = b2, this->B[2][3] = alpha;
  }
  bool implicit() const override { return true; }
  std::string name() const override { return "Alexander (claims order 3)"; }
};

// ---- gridoperator/onestep.hh --------------------------------------------------------------------------
template <class GO0, class GO1, bool implicit = true>
class OneStepGridOperator {
 public:
  using Pattern = typename GO0::Pattern;
  using Traits = typename GO0::Traits;
  using Domain = typename GO0::Domain;
  using Range = typename GO0::Range;
  using Jacobian = typename GO0::Jacobian;
  template <class MFT>
  struct MatrixContainer {
    using Type = Jacobian;
  };
  using Real = double;
  using OneStepParameters = TimeSteppingParameterInterface<Real>;

  // the facade OnTheFlyOperator / StationaryLinearProblemSolver / OneStepMethod touch (onestep/localassembler.hh)
  struct LocalAssembler {
    OneStepGridOperator* igo;
    static constexpr bool isLinear() { return true; }
    Real timeAtStage(int stage) const { return igo->timeAtStage(stage); }
    const typename Traits::TrialGridFunctionSpaceConstraints& trialConstraints() const {
      return igo->go0_.localAssembler().trialConstraints();
    }
  };

  // onestep.hh:66-76
  OneStepGridOperator(GO0& go0, GO1& go1) : go0_(go0), go1_(go1), la_{this} {
    GO0::setupGridOperators(std::tie(go0_, go1_));  // onestep.hh:73: go0 pre-processes, go1 post-processes
    create();
    if (!implicit) dt_mode_ = PDB200_ONESTEP_DO_NOT_ASSEMBLE_DT;
  }
  OneStepGridOperator(const OneStepGridOperator&) = delete;
  OneStepGridOperator& operator=(const OneStepGridOperator&) = delete;
  ~OneStepGridOperator() {
    if (os_) pdb200_onestep_destroy(os_);
  }

  // onestep.hh:78-91
  void divideMassTermByDeltaT() {
    if (!implicit) throw Exception("This function should not be called in explicit mode");
    dt_mode_ = PDB200_ONESTEP_DIVIDE_OPERATOR1_BY_DT;
    if (method_) check(pdb200_onestep_set_dt_mode(os_, dt_mode_), "divideMassTermByDeltaT");
  }
  void multiplySpatialTermByDeltaT() {
    if (!implicit) throw Exception("This function should not be called in explicit mode");
    dt_mode_ = PDB200_ONESTEP_MULTIPLY_OPERATOR0_BY_DT;
    if (method_) check(pdb200_onestep_set_dt_mode(os_, dt_mode_), "multiplySpatialTermByDeltaT");
  }

  const typename Traits::TrialGridFunctionSpace& trialGridFunctionSpace() const { return go0_.trialGridFunctionSpace(); }
  const typename Traits::TestGridFunctionSpace& testGridFunctionSpace() const { return go0_.testGridFunctionSpace(); }
  std::size_t globalSizeU() const { return go0_.globalSizeU(); }
  std::size_t globalSizeV() const { return go0_.globalSizeV(); }
  LocalAssembler& localAssembler() const { return la_; }
  const typename Traits::MatrixBackend& matrixBackend() const { return go0_.matrixBackend(); }

  // onestep.hh:113-128: the pattern of the spatial operator (the mass couplings are a subset of it)
  void fill_pattern(Pattern& p) const { go0_.fill_pattern(p); }

  // onestep.hh:245-254
  void setMethod(const OneStepParameters& method) {
    method_ = &method;
    const int s = (int)method.s();
    std::vector<double> a((std::size_t)s * (s + 1), 0.0), b(a), d(s + 1, 0.0);
    for (int r = 1; r <= s; r++)
      for (int i = 0; i <= r; i++) a[(r - 1) * (s + 1) + i] = method.a(r, i), b[(r - 1) * (s + 1) + i] = method.b(r, i);
    for (int i = 0; i <= s; i++) d[i] = method.d(i);
    check(pdb200_onestep_set_method(os_, s, a.data(), b.data(), d.data(), method.implicit() ? 1 : 0), "setMethod");
    if (method.implicit()) check(pdb200_onestep_set_dt_mode(os_, dt_mode_), "setMethod");
  }
  void preStep(const OneStepParameters& method, Real time, Real dt) {
    setMethod(method);
    time_ = time, dt_ = dt;
    check(pdb200_onestep_pre_step(os_, time, dt), "preStep");
  }
  Real timeAtStage(int stage) const {
    double t = 0;
    check(pdb200_onestep_time_at_stage(os_, stage, &t), "timeAtStage");
    return t;
  }

  // onestep.hh:130-139 with the per-stage times of prestageengine.hh:208-211
  void preStage(unsigned stage, const std::vector<Domain*>& x) {
    if (!implicit) throw Exception("This function should not be called in explicit mode");
    if (x.size() < stage) throw Exception("preStage: the solutions of stages 0..r-1 are needed");
    check(pdb200_onestep_pre_stage_begin(os_, (int)stage), "preStage");
    for (unsigned i = 0; i < stage; i++) {
      setTime(timeAtStage((int)i));
      check(pdb200_onestep_pre_stage_add(os_, (int)i, x[i]->data()), "preStage");
    }
    setTime(timeAtStage((int)stage));  // residualengine.hh:155-156: the stage itself lives at t + d_r dt
  }
  // explicit_jacobian_residual (onestep.hh:161-178) + the mass solve of ExplicitOneStepMethod::apply
  // (instationary/explicitonestep.hh:365-407) as one device call: x_r = -M^-1 sum_i (a_ri M x_i + b_ri dt R0(x_i))
  void explicit_stage(unsigned stage, const std::vector<Domain*>& x, Domain& xr, double reduction) {
    if (implicit) throw Exception("This function should not be called in implicit mode");
    if (x.size() < stage) throw Exception("explicit stage: the solutions of stages 0..r-1 are needed");
    // split per earlier stage: R0(x_i) is evaluated with the coefficients at t + d_i dt (prestageengine.hh:208-211)
    check(pdb200_onestep_explicit_stage_begin(os_, (int)stage), "explicit_jacobian_residual");
    for (unsigned i = 0; i < stage; i++) {
      setTime(timeAtStage((int)i));
      check(pdb200_onestep_explicit_stage_add(os_, (int)i, x[i]->data()), "explicit_jacobian_residual");
    }
    check(pdb200_onestep_explicit_stage_finish(os_, xr.data(), reduction), "explicit_jacobian_residual");
  }
  // onestep.hh:141-149
  void residual(const Domain& x, Range& r) const {
    if (!implicit) throw Exception("This function should not be called in explicit mode");
    check(pdb200_onestep_residual(os_, x.data(), r.data()), "residual");
  }
  // onestep.hh:151-159
  void jacobian(const Domain& x, Jacobian& a) const {
    if (!implicit) throw Exception("This function should not be called in explicit mode");
    check(pdb200_onestep_jacobian(os_, x.data(), a.values().data(), PDB200_LAYOUT_CSR), "jacobian");
  }
  // onestep.hh:180-185
  void jacobian_apply(const Domain& update, Range& result) const {
    check(pdb200_onestep_jacobian_apply(os_, update.data(), result.data()), "jacobian_apply");
  }
  // onestep.hh:187-192: both operators are linear
  void jacobian_apply(const Domain&, const Domain&, Range&) const {
    throw Exception("Your trying to use a non linear jacobian apply for a linear problem.");
  }
  void onthefly_apply(const double* x, double* y) const { check(pdb200_onestep_onthefly_apply(os_, x, y), "apply"); }
  void jacobian_apply(const double* z, double* y) const { check(pdb200_onestep_jacobian_apply(os_, z, y), "jacobian_apply"); }
  void residual(const double* x, double* r) const { check(pdb200_onestep_residual(os_, x, r), "residual"); }

  // onestep.hh:194-211
  template <class F, class X>
  void interpolate(unsigned stage, const X& xold, F& f, X& x) const {
    const Real t = timeAtStage((int)stage);
    f.setTime(t);
    go0_.localAssembler().setTime(t);
    go0_.interpolate(xold, f, x);  // includes copy_nonconstrained_dofs(xold -> x)
  }

  void postStep() {}
  void postStage() {}
  Real suggestTimestep(Real dt) const { return dt; }  // both local operators keep the suggested step (no CFL limit)
  void update() {
    go0_.update();
    go1_.update();
    recreate();
  }
  void make_consistent(Jacobian&) const {}

  // the fused operator of the current stage: what the linear solver back-ends bind to
  pdb200_handle handle() const {
    pdb200_handle st = nullptr;
    check(pdb200_onestep_stage_operator(os_, &st), "stage operator");
    return st;
  }
  pdb200_onestep_handle onestepHandle() const { return os_; }
  void solveStationary(int solver, int precond, bool matrix_free, double* x, double reduction, double min_defect,
                       unsigned maxiter, pdb200_solve_result* s) const {
    check(pdb200_onestep_solve_stationary(os_, solver, precond, matrix_free ? 1 : 0, x, reduction, min_defect, maxiter, s),
          "StationaryLinearProblemSolver::apply");
  }
  std::vector<std::uint64_t> constrained() const { return go0_.constrained(); }

 private:
  void create() {
    check(pdb200_onestep_create(go0_.handle(), go1_.handle(), &os_), "OneStepGridOperator");
    h0_ = go0_.handle(), h1_ = go1_.handle();
  }
  void recreate() {
    if (os_) pdb200_onestep_destroy(os_);
    os_ = nullptr;
    create();
    if (method_) {
      setMethod(*method_);
      check(pdb200_onestep_pre_step(os_, time_, dt_), "preStep");
    }
  }
  // la0.setTime / la1.setTime; a re-created operator (a coefficient field switched on or off over time) cannot keep
  // the constant part that is being assembled
  void setTime(Real t) {
    go0_.setTime(t);
    go1_.setTime(t);
    if (go0_.handle() != h0_ || go1_.handle() != h1_)
      throw Exception("OneStepGridOperator: the set of active coefficient fields or the boundary types changed over "
                      "time; call update() between time steps (the constraint set and the stage operator are rebuilt)");
  }

  GO0& go0_;
  GO1& go1_;
  mutable LocalAssembler la_;
  pdb200_onestep_handle os_ = nullptr;
  pdb200_handle h0_ = nullptr, h1_ = nullptr;
  const OneStepParameters* method_ = nullptr;
  int dt_mode_ = PDB200_ONESTEP_MULTIPLY_OPERATOR0_BY_DT;
  Real time_ = 0.0, dt_ = 1.0;
};

// OnTheFlyOperator on the one-step operator binds like on a GridOperator (igo.onthefly_apply / jacobian_apply).

// ---- instationary/implicitonestep.hh --------------------------------------------------------------------
struct OneStepMethodPartialResult {
  unsigned timesteps = 0;
  double assembler_time = 0.0, linear_solver_time = 0.0;
  int linear_solver_iterations = 0, nonlinear_solver_iterations = 0;
};
struct OneStepMethodResult {
  OneStepMethodPartialResult total, successful;
};

template <class T, class IGOS, class PDESOLVER, class TrlV, class TstV = TrlV>
class OneStepMethod {
 public:
  using Result = OneStepMethodResult;
  // implicitonestep.hh:67-78
  OneStepMethod(const TimeSteppingParameterInterface<T>& method, IGOS& igos, PDESOLVER& pdesolver)
      : method_(&method), igos_(igos), pdesolver_(pdesolver) {}
  void setVerbosityLevel(int level) { verbosity_ = level; }
  void setStepNumber(int newstep) { step_ = newstep; }
  const Result& result() const { return res_; }
  void setMethod(const TimeSteppingParameterInterface<T>& method) { method_ = &method; }

  // implicitonestep.hh:122-262
  T apply(T time, T dt, TrlV& xold, TrlV& xnew) {
    return run(time, dt, xold, xnew, [&](unsigned r, std::vector<TrlV*>& x) {
      if (r > 1) *(x[r]) = *(x[r - 1]);  // result of the last stage as initial guess
      else if (x[r] != &xnew) *(x[r]) = xnew;
    });
  }
  // implicitonestep.hh:264-400: constraints are interpolated from f at the start of each stage
  template <class F>
  T apply(T time, T dt, TrlV& xold, F& f, TrlV& xnew) {
    return run(time, dt, xold, xnew, [&](unsigned r, std::vector<TrlV*>& x) {
      const TrlV* init_guess = (r == 1) ? &xnew : x[r - 1];
      TrlV guess(*init_guess);
      igos_.interpolate(r, guess, f, *x[r]);
    });
  }

 private:
  template <class Init>
  T run(T time, T dt, TrlV& xold, TrlV& xnew, Init&& init) {
    OneStepMethodPartialResult step_result;
    std::vector<TrlV*> x(1, &xold);
    std::vector<std::unique_ptr<TrlV>> owned;
    if (verbosity_ >= 1)
      std::cout << "TIME STEP [" << method_->name() << "] " << std::setw(6) << step_ << " time (from): " << std::scientific
                << time << " dt: " << dt << " time (to): " << time + dt << std::endl;
    igos_.preStep(*method_, time, dt);
    for (unsigned r = 1; r <= method_->s(); ++r) {
      if (verbosity_ >= 2) std::cout << "STAGE " << r << " time (to): " << time + method_->d(r) * dt << "." << std::endl;
      igos_.preStage(r, x);
      if (r == method_->s()) {
        x.push_back(&xnew);
      } else {
        owned.emplace_back(new TrlV(igos_.trialGridFunctionSpace()));
        x.push_back(owned.back().get());
      }
      init(r, x);
      pdesolver_.apply(*x[r]);
      const auto& pderes = pdesolver_.result();
      step_result.linear_solver_time += pderes.linear_solver_time;
      step_result.linear_solver_iterations += pderes.linear_solver_iterations;
      step_result.nonlinear_solver_iterations += 1;
      igos_.postStage();
    }
    igos_.postStep();
    step_result.timesteps = 1;
    for (auto* p : {&res_.total, &res_.successful}) {
      p->timesteps += 1;
      p->linear_solver_time += step_result.linear_solver_time;
      p->linear_solver_iterations += step_result.linear_solver_iterations;
      p->nonlinear_solver_iterations += step_result.nonlinear_solver_iterations;
    }
    if (verbosity_ >= 1)
      std::cout << "::: timesteps      " << std::setw(6) << res_.successful.timesteps << "\n::: lin iterations "
                << std::setw(6) << res_.successful.linear_solver_iterations << std::endl;
    step_++;
    return dt;
  }

  const TimeSteppingParameterInterface<T>* method_;
  IGOS& igos_;
  PDESOLVER& pdesolver_;
  int verbosity_ = 1, step_ = 1;
  Result res_;
};

// ---- instationary/explicitonestep.hh ---------------------------------------------------------------------
// ExplicitOneStepMethod<T, IGOS, LS, TrlV, TstV, TC> (explicitonestep.hh:181-414) for QkDG spaces.  The linear solver
// LS of the reference only ever sees the block-diagonal mass matrix; here that solve is the exact block inverse inside
// pdb200_onestep_explicit_stage, so LS is accepted for interface parity and `reduction` is forwarded for the degrees
// without a closed-form inverse.  The time-step controller (CFL limit) is not part of this path: dt is used as given.
template <class T, class IGOS, class LS, class TrlV, class TstV = TrlV>
class ExplicitOneStepMethod {
 public:
  ExplicitOneStepMethod(const TimeSteppingParameterInterface<T>& method, IGOS& igos, LS& ls, double ls_reduction = 0.99)
      : method_(&method), igos_(igos), ls_(ls), reduction_(ls_reduction) {
    if (method.implicit()) throw Exception("explicit one step method called with implicit scheme");  // :226-228
  }
  void setVerbosityLevel(int level) { verbosity_ = level; }
  void setStepNumber(int newstep) { step_ = newstep; }
  void setReduction(const double& r) { reduction_ = r; }
  void setMethod(const TimeSteppingParameterInterface<T>& method) {
    if (method.implicit()) throw Exception("explicit one step method called with implicit scheme");
    method_ = &method;
  }
  // explicitonestep.hh:282-414
  T apply(T time, T dt, TrlV& xold, TrlV& xnew) {
    std::vector<TrlV*> x(1, &xold);
    std::vector<std::unique_ptr<TrlV>> owned;
    if (verbosity_ >= 1)
      std::cout << "TIME STEP [" << method_->name() << "] " << std::setw(6) << step_ << " time (from): " << std::scientific
                << time << " dt: " << dt << " time (to): " << time + dt << std::endl;
    igos_.preStep(*method_, time, dt);
    for (unsigned r = 1; r <= method_->s(); ++r) {
      if (r == method_->s()) {
        x.push_back(&xnew);
      } else {
        owned.emplace_back(new TrlV(igos_.trialGridFunctionSpace()));
        x.push_back(owned.back().get());
      }
      igos_.explicit_stage(r, x, *x[r], reduction_ >= 0.99 ? 1e-12 : reduction_);
      igos_.postStage();
    }
    igos_.postStep();
    step_++;
    return dt;
  }

 private:
  const TimeSteppingParameterInterface<T>* method_;
  IGOS& igos_;
  LS& ls_;
  double reduction_;
  int verbosity_ = 1, step_ = 1;
};

}  // namespace B200
}  // namespace PDELab
}  // namespace Dune

#endif  // PDELAB_B200_HOST_ONESTEP_HH

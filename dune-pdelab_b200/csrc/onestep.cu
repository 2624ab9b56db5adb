// onestep.cu — OneStepGridOperator on the device (gridoperator/onestep.hh:30-308 and the engines in
// gridoperator/onestep/*.hh): the operator of one Runge-Kutta / fractional-step stage
//
//     r  +=  b_rr * dt * R0(x; t + d_r dt)  +  R1(x)  +  const_residual,
//     const_residual = sum_{s<r}  b_rs * dt * R0(x_s; t + d_s dt)  +  a_rs * R1(x_s)            (preStage)
//
// with R0 the spatial operator (ConvectionDiffusionDG / ConvectionDiffusionFEM) and R1 the temporal one (L2).
//
// B200 design: the reference runs the two local assemblers side by side with engine weights
// (onestep/residualengine.hh:135-139, prestageengine.hh:205-230).  Both operators of this path are
// convection-diffusion-reaction forms that are homogeneous of degree one in their coefficient fields
// (A, b, c, f, j, o) — the SIPG face weights omega_s = delta_n / (delta_s + delta_n) are degree zero, the
// penalty is degree one — so the weighted sum  w0 R0 + w1 R1  IS the convection-diffusion operator with the
// coefficient fields  w0 (A, b, c, f, j, o)_0 + w1 (A, b, c, f, j, o)_1  and the Dirichlet data g of the
// spatial operator.  A stage therefore costs ONE pass of the Kronecker kernels over the vectors (same HBM
// traffic as a stationary apply, the mass term rides in the reaction slot) instead of two assemblies; only
// the per-cell coefficient arrays (1/n of a vector) are rewritten when the weights change.  A negative w0
// (some a_rs, b_rs of the pre-stage sums are negative) would flip the upwind direction of b, so the sign is
// pulled out:  w0 R0 + w1 R1 = sgn(w0) (|w0| R0 + sgn(w0) w1 R1).
//
// With weightsOff the penalty is independent of A; there the weight goes into the penalty constant alpha.
//
// Deviation (documented in DESIGN.md): rounding differs from the reference's  w0 * (...) + w1 * (...)  by a few
// ulp, and the 1e-20 regularisation of the harmonic weights (convectiondiffusiondg.hh:330-332) acts on the
// scaled permeability — relative effect 1e-20 / (w0 delta).

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "operator.h"

struct pdb200_onestep {
  pdb200_operator* go0 = nullptr;  // spatial operator (not owned)
  pdb200_operator* go1 = nullptr;  // temporal operator (not owned)
  pdb200_operator* stage = nullptr;  // the fused stage operator (owned)
  std::vector<void*> owned;        // temporaries used to describe the stage operator
  double* const_residual = nullptr;
  double* tmp = nullptr;
  double *hx = nullptr, *hr = nullptr;  // staging of host vectors
  // method (TimeSteppingParameterInterface, instationary/onestepparameter.hh:43-84)
  int s = 0;
  std::vector<double> a, b, d;  // a, b: s x (s+1) row-major (row r-1 <-> stage r), d: s+1
  bool implicit_method = true;
  int dt_mode = PDB200_ONESTEP_MULTIPLY_OPERATOR0_BY_DT;
  double time = 0.0, dt = 1.0, dt_factor0 = 1.0, dt_factor1 = 1.0;
  int stage_no = 0;
  // what the stage operator currently holds
  bool combined_valid = false;
  double cw0 = 0.0, cw1 = 0.0;
  uint64_t cv0 = 0, cv1 = 0;
  uint64_t launches = 0;

  ~pdb200_onestep() {
    if (go0) cudaSetDevice(go0->device);
    for (void* p : owned) cudaFree(p);
    if (const_residual) cudaFree(const_residual);
    if (tmp) cudaFree(tmp);
    if (hx) cudaFree(hx);
    if (hr) cudaFree(hr);
    if (stage) pdb200_destroy(stage);
  }
};

namespace {

#define OS_C(call)                                                    \
  do {                                                                \
    if ((call) != 0) throw Error(std::string(pdb200_last_error()));   \
  } while (0)

bool is_dev(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

size_t a_len(const DevParams& P, int mode) {
  switch (mode) {
    case PDB200_A_IDENTITY: return 0;
    case PDB200_A_SCALAR: return (size_t)P.ncells;
    case PDB200_A_DIAGONAL: return (size_t)P.ncells * P.dim;
    default: return (size_t)P.ncells * P.dim * P.dim;
  }
}

// out[i] = w0 * x0[i] + w1 * x1[i]; a null input stands for the constant c0 / c1 (0 for an absent field)
__global__ void combine_kernel(double* __restrict__ out, const double* __restrict__ x0, const double* __restrict__ x1,
                               double w0, double w1, double c0, double c1, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v0 = x0 ? x0[i] : c0, v1 = x1 ? x1[i] : c1;
  out[i] = w0 * v0 + w1 * v1;
}

// A of the stage operator in the layout `mode` from an operator's own tensor field (any cheaper layout)
__global__ void combine_tensor_kernel(double* __restrict__ out, int mode, int dim, const double* __restrict__ A0, int mode0,
                                      double w0, const double* __restrict__ A1, int mode1, double w1, long long ncells) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= ncells) return;
  auto entry = [&](const double* A, int m, int i, int j) -> double {
    switch (m) {
      case PDB200_A_IDENTITY: return i == j ? 1.0 : 0.0;
      case PDB200_A_SCALAR: return i == j ? A[e] : 0.0;
      case PDB200_A_DIAGONAL: return i == j ? A[e * dim + i] : 0.0;
      default: return A[(e * dim + i) * dim + j];
    }
  };
  if (mode == PDB200_A_SCALAR) {
    out[e] = w0 * entry(A0, mode0, 0, 0) + w1 * entry(A1, mode1, 0, 0);
  } else if (mode == PDB200_A_DIAGONAL) {
    for (int i = 0; i < dim; i++) out[e * dim + i] = w0 * entry(A0, mode0, i, i) + w1 * entry(A1, mode1, i, i);
  } else {
    for (int i = 0; i < dim; i++)
      for (int j = 0; j < dim; j++) out[(e * dim + i) * dim + j] = w0 * entry(A0, mode0, i, j) + w1 * entry(A1, mode1, i, j);
  }
}

__global__ void scale_kernel(double* __restrict__ y, double a, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] *= a;
}

__global__ void axpy_kernel(double* __restrict__ y, const double* __restrict__ x, double a, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}

unsigned grid_for(long long n) { return (unsigned)((n + 255) / 256); }

// the temporal operator must not bring boundary terms of its own: the stage operator integrates the boundary
// conditions of the spatial operator.  L2 (localoperator/l2.hh) has none; its mirror carries type None and A = 0.
void check_compatible(const pdb200_operator* g0, const pdb200_operator* g1) {
  const DevParams &P0 = g0->P, &P1 = g1->P;
  if (g0->device != g1->device) throw Error("OneStepGridOperator: both operators must live on the same device");
  if (P0.dim != P1.dim || P0.k != P1.k || P0.dg != P1.dg || P0.basis != P1.basis || P0.ncells != P1.ncells ||
      P0.ndofs != P1.ndofs)
    throw Error("OneStepGridOperator: the two grid operators need the same grid and function space "
                "(gridoperator/onestep/localassembler.hh:65-84)");
  for (int d = 0; d < 3; d++) {
    if (P0.N[d] != P1.N[d] || P0.h[d] != P1.h[d]) throw Error("OneStepGridOperator: grids differ");
    for (int s = 0; s < 2; s++)
      if (P0.side_kind[d][s] != P1.side_kind[d][s]) throw Error("OneStepGridOperator: partitions differ");
  }
  if (P0.m < P1.m)
    throw Error("OneStepGridOperator: the spatial operator's quadrature must be at least as fine as the temporal one's");
  if (P1.b || P1.g || P1.j || P1.o || (P1.f && P1.m != P0.m))
    throw Error("OneStepGridOperator: the temporal operator may only carry A, c and f (on the same quadrature)");
  if (P0.pw || P1.pw)
    throw Error("OneStepGridOperator: the stage operator combines cell-wise coefficient fields; the point-wise layouts "
                "(pdb200_problem::pointwise) are not supported here");
  if (P0.dg) {
    // The fused operator w0 go0 + w1 go1 is evaluated with go0's face terms on the COMBINED tensor: with a diffusion
    // tensor in go1 the harmonic weights and the penalty (non-linear in A) and go0's Dirichlet / Neumann terms would see
    // it too, which is not w0 R0 + w1 R1.  A QkDG temporal operator must be a pure reaction form without face terms
    // (L2, localoperator/l2.hh): A == 0 and no boundary integrals.
    if (P1.a_mode != PDB200_A_IDENTITY && P1.A) {
      std::vector<double> a1(a_len(P1, P1.a_mode));
      PDB_CUDA(cudaMemcpy(a1.data(), P1.A, a1.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (double v : a1)
        if (v != 0.0) throw Error("OneStepGridOperator: on QkDG spaces the temporal operator must be an L2 mass operator (A == 0)");
    } else if (P1.a_mode == PDB200_A_IDENTITY) {
      throw Error("OneStepGridOperator: on QkDG spaces the temporal operator must be an L2 mass operator (A == 0, not the identity)");
    }
    bool none = P1.bctype != nullptr;
    if (none) {
      long long nbf1 = 0;
      for (int d = 0; d < P1.dim; d++) nbf1 += 2 * (P1.ncells / P1.N[d]);
      std::vector<int8_t> bt((size_t)nbf1);
      PDB_CUDA(cudaMemcpy(bt.data(), P1.bctype, bt.size(), cudaMemcpyDeviceToHost));
      for (int8_t v : bt) none = none && v == PDB200_BC_NONE;
    }
    bool has_domain_side = false;
    for (int d = 0; d < P1.dim; d++)
      for (int sd = 0; sd < 2; sd++) has_domain_side |= P1.side_kind[d][sd] == PDB200_SIDE_DOMAIN;
    if (has_domain_side && !none)
      throw Error("OneStepGridOperator: on QkDG spaces the temporal operator must carry boundary type None on every face "
                  "(an L2 mass operator has no boundary integrals)");
  }
}

void build_stage_operator(pdb200_onestep* os) {
  pdb200_operator *g0 = os->go0, *g1 = os->go1;
  const DevParams &P0 = g0->P, &P1 = g1->P;
  PDB_CUDA(cudaSetDevice(g0->device));
  pdb200_problem p;
  std::memset(&p, 0, sizeof(p));
  p.dim = P0.dim;
  for (int d = 0; d < 3; d++) {
    p.cells[d] = P0.N[d];
    p.lower[d] = g0->lower[d];
    p.upper[d] = g0->upper[d];
    for (int s = 0; s < 2; s++) p.side_kind[d][s] = P0.side_kind[d][s];
  }
  p.space = P0.dg ? PDB200_SPACE_QKDG : PDB200_SPACE_QK;
  p.degree = P0.k;
  p.dg_method = P0.theta == -1.0 ? PDB200_DG_SIPG : (P0.theta == 0.0 ? PDB200_DG_IIPG : PDB200_DG_NIPG);
  p.dg_weights = P0.weights_on ? PDB200_DG_WEIGHTS_ON : PDB200_DG_WEIGHTS_OFF;
  p.dg_alpha = P0.alpha;
  p.intorderadd = (P0.m - 1) * 2 - 2 * P0.k;  // m = (2k + intorderadd)/2 + 1, intorderadd in {0, 1} -> even representative
  if (p.intorderadd < 0) p.intorderadd = 0;
  // tensor layout: the richer of the two, at least scalar (the weight has to live somewhere)
  int mode = P0.a_mode > P1.a_mode ? P0.a_mode : P1.a_mode;
  if (mode == PDB200_A_IDENTITY) mode = PDB200_A_SCALAR;
  p.a_mode = mode;
  auto dalloc = [&](size_t count) -> double* {
    double* q = nullptr;
    PDB_CUDA(cudaMalloc(&q, count * sizeof(double)));
    PDB_CUDA(cudaMemset(q, 0, count * sizeof(double)));
    os->owned.push_back(q);
    return q;
  };
  long long nbf = 0;
  for (int d = 0; d < P0.dim; d++) nbf += 2 * (P0.ncells / P0.N[d]);
  p.A = dalloc(a_len(P0, mode));
  p.b = P0.b ? dalloc((size_t)P0.ncells * P0.dim) : nullptr;
  p.c = dalloc((size_t)P0.ncells);
  p.f = (P0.f || P1.f) ? dalloc((size_t)P0.ncells * P0.nq) : nullptr;
  p.bctype = P0.bctype;  // device pointer, copied by pdb200_create
  p.g = P0.g;            // Dirichlet data is not weighted (it multiplies A- and b-dependent factors)
  p.j = P0.j ? dalloc((size_t)nbf * P0.nfq) : nullptr;
  p.o = P0.o ? dalloc((size_t)nbf * P0.nfq) : nullptr;
  p.device = g0->device;
  p.kernel = g0->kernel_choice;
  p.basis = P0.basis;
  pdb200_handle st = nullptr;
  OS_C(pdb200_create(&p, &st));
  os->stage = st;
  // pdb200_create copied every array: the descriptions are no longer needed
  for (void* q : os->owned) cudaFree(q);
  os->owned.clear();
  if (st->P.m != P0.m) throw Error("OneStepGridOperator: internal error (quadrature of the stage operator)");
  st->stream = g0->stream;
  os->combined_valid = false;
}

// stage operator <- w0 * go0 + w1 * go1 (coefficient fields only; cached per (w0, w1, coefficient versions))
void combine(pdb200_onestep* os, double w0, double w1) {
  pdb200_operator *g0 = os->go0, *g1 = os->go1, *st = os->stage;
  if (os->combined_valid && os->cw0 == w0 && os->cw1 == w1 && os->cv0 == g0->coeff_version && os->cv1 == g1->coeff_version)
    return;
  const DevParams &P0 = g0->P, &P1 = g1->P, &PS = st->P;
  if (w0 < 0.0 && P0.b) throw Error("OneStepGridOperator: internal error (negative weight on an upwinded operator)");
  cudaStream_t s = st->stream;
  auto lin = [&](const double* out, const double* x0, const double* x1, double c0, double c1, long long n) {
    if (!out || n == 0) return;
    combine_kernel<<<grid_for(n), 256, 0, s>>>(const_cast<double*>(out), x0, x1, w0, w1, c0, c1, n);
    os->launches++;
  };
  combine_tensor_kernel<<<grid_for(PS.ncells), 256, 0, s>>>(const_cast<double*>(PS.A), PS.a_mode, PS.dim, P0.A, P0.a_mode, w0,
                                                             P1.A, P1.a_mode, w1, PS.ncells);
  os->launches++;
  long long nbf = 0;
  for (int d = 0; d < PS.dim; d++) nbf += 2 * (PS.ncells / PS.N[d]);
  lin(PS.b, P0.b, nullptr, 0.0, 0.0, PS.ncells * PS.dim);
  lin(PS.c, P0.c, P1.c, 0.0, 0.0, PS.ncells);
  lin(PS.f, P0.f, P1.f, 0.0, 0.0, PS.ncells * PS.nq);
  lin(PS.j, P0.j, nullptr, 0.0, 0.0, nbf * PS.nfq);
  lin(PS.o, P0.o, nullptr, 0.0, 0.0, nbf * PS.nfq);
  if (P0.g && PS.g && os->cv0 != g0->coeff_version)
    PDB_CUDA(cudaMemcpyAsync(const_cast<double*>(PS.g), P0.g, (size_t)nbf * PS.nfq * sizeof(double), cudaMemcpyDeviceToDevice, s));
  PDB_CUDA(cudaGetLastError());
  if (P0.dg && !P0.weights_on) {
    // weightsOff: the penalty  alpha / h_F * k (k + d - 1)  does not contain A (harmonic_average = 1,
    // convectiondiffusiondg.hh:334-338), so the weight of the spatial operator goes into alpha itself.  The
    // Kronecker / matrix plans bake alpha in: they are rebuilt on the next launch.
    const double alpha = w0 * P0.alpha;
    if (alpha != st->P.alpha) {
      PDB_CUDA(cudaStreamSynchronize(s));
      st->P.alpha = alpha;
      dg_fast_plan_destroy(st->fast);
      st->fast = nullptr;
      dg_kron_plan_destroy(st->kron);
      st->kron = nullptr;
      matrix_plan_destroy(st->matrix);
      st->matrix = nullptr;
      dg_blockjac_destroy(st->blockjac);
      st->blockjac = nullptr;
    }
  }
  // cached quantities of the stage operator depend on its coefficients
  st->r0_valid = false;
  if (pdb_uses_cached_r0(st) && !P1.f) {
    // R(0) is linear in (A, b, f, j, o) for fixed Dirichlet data and the temporal operator has none:
    // R_stage(0) = w0 R0(0), one scaling pass instead of a reference-order evaluation per weight change
    pdb_ensure_r0(g0);
    if (!st->r0) PDB_CUDA(cudaMalloc(&st->r0, (size_t)PS.ndofs * sizeof(double)));
    combine_kernel<<<grid_for(PS.ndofs), 256, 0, s>>>(st->r0, g0->r0, nullptr, w0, 0.0, 0.0, 0.0, PS.ndofs);
    os->launches++;
    PDB_CUDA(cudaGetLastError());
    st->r0_valid = true;
  }
  st->coeff_version++;
  dg_blockjac_invalidate(st->blockjac);
  fem_plan_invalidate(st->fem);
  os->combined_valid = true;
  os->cw0 = w0;
  os->cw1 = w1;
  os->cv0 = g0->coeff_version;
  os->cv1 = g1->coeff_version;
}

double coef(const std::vector<double>& m, int s, int r, int i) { return m[(size_t)(r - 1) * (s + 1) + i]; }

void need_method(const pdb200_onestep* os) {
  if (os->s <= 0) throw Error("OneStepGridOperator: no time-stepping method set (setMethod / preStep)");
}
void need_stage(const pdb200_onestep* os) {
  need_method(os);
  if (os->stage_no < 1 || os->stage_no > os->s) throw Error("OneStepGridOperator: no stage selected (preStage)");
}
void need_implicit(const pdb200_onestep* os, const char* what) {
  if (!os->implicit_method)
    throw Error(std::string("This function should not be called in explicit mode (") + what + ", gridoperator/onestep.hh)");
}

// weights of the stage operator itself (onestep/residualengine.hh:135-158, jacobianengine.hh:92-96,
// jacobianapplyengine.hh): la0 <- b_rr * dt_factor0 (skipped if |b_rr| <= 1e-6), la1 <- dt_factor1
void stage_weights(const pdb200_onestep* os, double* w0, double* w1) {
  const double b_rr = coef(os->b, os->s, os->stage_no, os->stage_no);
  const bool implicit = std::fabs(b_rr) > 1e-6;
  *w0 = implicit ? b_rr * os->dt_factor0 : 0.0;
  *w1 = os->dt_factor1;
  if (*w0 < 0.0 && os->go0->P.b)
    throw Error("OneStepGridOperator: a negative diagonal coefficient b_rr with a convective spatial operator is not supported");
}

void combine_stage(pdb200_onestep* os) {
  double w0, w1;
  stage_weights(os, &w0, &w1);
  combine(os, w0, w1);
}

// device view of a caller's vector (host vectors are staged, like every entry point of the C ABI)
struct Staged {
  double* dev = nullptr;
  double* host = nullptr;
  size_t bytes = 0;
  cudaStream_t s = nullptr;
  Staged(pdb200_onestep* os, double** slot, const double* p, bool copy_in) {
    bytes = (size_t)os->go0->P.ndofs * sizeof(double);
    s = os->stage->stream;
    if (is_dev(p)) {
      dev = const_cast<double*>(p);
      return;
    }
    if (!*slot) PDB_CUDA(cudaMalloc(slot, bytes));
    dev = *slot;
    host = const_cast<double*>(p);
    if (copy_in) PDB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s));
  }
  void copy_back() {
    if (!host) return;
    PDB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s));
    PDB_CUDA(cudaStreamSynchronize(s));
  }
};

}  // namespace

#define OS_TRY try {
#define OS_CATCH                     \
  }                                  \
  catch (const std::exception& e) {  \
    pdb_set_last_error(e.what());    \
    return 1;                        \
  }                                  \
  return 0;
#define OS_CHECK(os) \
  if (!(os)) throw Error("null one-step operator handle")

extern "C" {

int pdb200_onestep_create(pdb200_handle go0, pdb200_handle go1, pdb200_onestep_handle* out) {
  OS_TRY
  if (!go0 || !go1 || !out) throw Error("null argument");
  *out = nullptr;
  check_compatible(go0, go1);
  std::unique_ptr<pdb200_onestep> os(new pdb200_onestep);
  os->go0 = go0;
  os->go1 = go1;
  build_stage_operator(os.get());
  PDB_CUDA(cudaMalloc(&os->const_residual, (size_t)go0->P.ndofs * sizeof(double)));
  PDB_CUDA(cudaMemset(os->const_residual, 0, (size_t)go0->P.ndofs * sizeof(double)));
  *out = os.release();
  OS_CATCH
}

int pdb200_onestep_destroy(pdb200_onestep_handle os) {
  OS_TRY
  delete os;
  OS_CATCH
}

int pdb200_onestep_set_method(pdb200_onestep_handle os, int stages, const double* a, const double* b, const double* d,
                              int implicit) {
  OS_TRY
  OS_CHECK(os);
  if (stages < 1 || !a || !b || !d) throw Error("setMethod: need s >= 1 and the arrays a, b (s x (s+1)) and d (s+1)");
  os->s = stages;
  os->a.assign(a, a + (size_t)stages * (stages + 1));
  os->b.assign(b, b + (size_t)stages * (stages + 1));
  os->d.assign(d, d + stages + 1);
  os->implicit_method = implicit != 0;
  // OneStepGridOperator's constructor: explicit methods never assemble dt (onestep.hh:74-75).  In the reference
  // `implicit` is a template parameter of the grid operator; here one handle may see both kinds of method, so the
  // user's mode (dt_mode) is kept and the EFFECTIVE mode is derived from the method in pre_step.
  os->stage_no = 0;
  OS_CATCH
}

int pdb200_onestep_set_dt_mode(pdb200_onestep_handle os, int mode) {
  OS_TRY
  OS_CHECK(os);
  if (mode < 0 || mode > 2) throw Error("Unknown mode for assembling of time step size!");  // localassembler.hh:122-125
  if (!os->implicit_method && mode != PDB200_ONESTEP_DO_NOT_ASSEMBLE_DT)
    throw Error("This function should not be called in explicit mode");  // onestep.hh:78-91
  os->dt_mode = mode;
  OS_CATCH
}

int pdb200_onestep_pre_step(pdb200_onestep_handle os, double time, double dt) {
  OS_TRY
  OS_CHECK(os);
  need_method(os);
  if (!(dt > 0.0)) throw Error("preStep: dt must be positive");
  os->time = time;
  os->dt = dt;
  // onestep/localassembler.hh:101-130
  const int mode = os->implicit_method ? os->dt_mode : PDB200_ONESTEP_DO_NOT_ASSEMBLE_DT;
  if (mode == PDB200_ONESTEP_DIVIDE_OPERATOR1_BY_DT) {
    os->dt_factor0 = 1.0;
    os->dt_factor1 = 1.0 / dt;
  } else if (mode == PDB200_ONESTEP_MULTIPLY_OPERATOR0_BY_DT) {
    os->dt_factor0 = dt;
    os->dt_factor1 = 1.0;
  } else {
    os->dt_factor0 = 1.0;
    os->dt_factor1 = 1.0;
  }
  os->stage_no = 0;
  OS_CATCH
}

int pdb200_onestep_time_at_stage(pdb200_onestep_handle os, int stage, double* t) {
  OS_TRY
  OS_CHECK(os);
  need_method(os);
  if (stage < 0 || stage > os->s || !t) throw Error("timeAtStage: stage out of range");
  *t = os->time + os->d[stage] * os->dt;  // localassembler.hh:148-156
  OS_CATCH
}

int pdb200_onestep_pre_stage_begin(pdb200_onestep_handle os, int stage) {
  OS_TRY
  OS_CHECK(os);
  need_method(os);
  if (stage < 1 || stage > os->s) throw Error("preStage: stage must be in 1..s");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  os->stage_no = stage;
  os->stage->stream = os->go0->stream;
  // prestageengine.hh:166-172
  PDB_CUDA(cudaMemsetAsync(os->const_residual, 0, (size_t)os->go0->P.ndofs * sizeof(double), os->stage->stream));
  OS_CATCH
}

// const_residual += w0 R0(x) + w1 R1(x) through the fused stage operator (sign of w0 pulled out, see the header)
static void add_weighted_residual(pdb200_onestep* os, double w0, double w1, const double* x) {
  if (w0 == 0.0 && w1 == 0.0) return;
  const double sgn = w0 < 0.0 ? -1.0 : 1.0;
  combine(os, sgn * w0, sgn * w1);
  Staged X(os, &os->hx, x, true);
  const long long n = os->go0->P.ndofs;
  cudaStream_t s = os->stage->stream;
  if (sgn > 0.0) {
    OS_C(pdb200_residual(os->stage, X.dev, os->const_residual));
  } else {
    if (!os->tmp) PDB_CUDA(cudaMalloc(&os->tmp, (size_t)n * sizeof(double)));
    PDB_CUDA(cudaMemsetAsync(os->tmp, 0, (size_t)n * sizeof(double), s));
    OS_C(pdb200_residual(os->stage, X.dev, os->tmp));
    axpy_kernel<<<grid_for(n), 256, 0, s>>>(os->const_residual, os->tmp, -1.0, n);
    os->launches++;
    PDB_CUDA(cudaGetLastError());
  }
  if (X.host) PDB_CUDA(cudaStreamSynchronize(s));
}

int pdb200_onestep_pre_stage_add(pdb200_onestep_handle os, int i, const double* x) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  if (i < 0 || i >= os->stage_no || !x) throw Error("preStage: the solutions of stages 0..r-1 are needed");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  // prestageengine.hh:180-186, 205-230
  const double a = coef(os->a, os->s, os->stage_no, i), b = coef(os->b, os->s, os->stage_no, i);
  const bool do0 = std::fabs(b) > 1e-6, do1 = std::fabs(a) > 1e-6;
  add_weighted_residual(os, do0 ? b * os->dt_factor0 : 0.0, do1 ? a * os->dt_factor1 : 0.0, x);
  OS_CATCH
}

// One stage of an EXPLICIT method (ExplicitOneStepMethod::apply, instationary/explicitonestep.hh:332-414 with
// OneStepGridOperator::explicit_jacobian_residual, gridoperator/onestep.hh:161-178 and
// onestep/jacobianresidualengine.hh:290-400):  D = -M,  alpha = sum_i a_ri R1(x_i),  beta = sum_i b_ri R0(x_i),
// alpha += dt beta,  solve D x_r = alpha,  i.e.  x_r = -M^-1 (sum_{i<r} a_ri M x_i + b_ri dt R0(x_i)).
// The mass matrix of a QkDG space is block diagonal: for k <= 2 the solve is the exact block inverse by fast
// diagonalisation (one kernel, no Krylov loop), otherwise a CG on the mass operator to `reduction`.
// The stage is split like preStage so that a host with time-dependent coefficients can re-sample them at
// t + d_i dt before add(i): the jacobian-residual engine delegates to the pre-stage engine, which sets
// la0.setTime(time + d[s] dt) for every earlier stage s (jacobianresidualengine.hh, prestageengine.hh:208-211).
int pdb200_onestep_explicit_stage_begin(pdb200_onestep_handle os, int stage) {
  OS_TRY
  OS_CHECK(os);
  need_method(os);
  if (os->implicit_method) throw Error("explicit one step method called with implicit scheme");  // explicitonestep.hh:226-228
  if (stage < 1 || stage > os->s) throw Error("explicit stage: stage must be in 1..s");
  pdb200_operator* g0 = os->go0;
  if (!g0->P.dg)
    throw Error("explicit one-step methods need a block-diagonal mass matrix (QkDG spaces)");
  PDB_CUDA(cudaSetDevice(g0->device));
  os->stage_no = stage;
  os->stage->stream = g0->stream;
  PDB_CUDA(cudaMemsetAsync(os->const_residual, 0, (size_t)g0->P.ndofs * sizeof(double), os->stage->stream));
  OS_CATCH
}

int pdb200_onestep_explicit_stage_add(pdb200_onestep_handle os, int i, const double* x) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  if (os->implicit_method) throw Error("explicit one step method called with implicit scheme");
  if (i < 0 || i >= os->stage_no || !x) throw Error("explicit stage: the solutions of stages 0..r-1 are needed");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  const double a = coef(os->a, os->s, os->stage_no, i), b = coef(os->b, os->s, os->stage_no, i);
  add_weighted_residual(os, std::fabs(b) > 1e-6 ? b * os->dt : 0.0, std::fabs(a) > 1e-6 ? a : 0.0, x);
  OS_CATCH
}

int pdb200_onestep_explicit_stage_finish(pdb200_onestep_handle os, double* xr, double reduction) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  if (os->implicit_method) throw Error("explicit one step method called with implicit scheme");
  if (!xr) throw Error("explicit stage: null result vector");
  pdb200_operator *g0 = os->go0, *g1 = os->go1;
  PDB_CUDA(cudaSetDevice(g0->device));
  cudaStream_t s = os->stage->stream;
  const long long n = g0->P.ndofs;
  Staged XR(os, &os->hr, xr, true);
  g1->stream = s;
  if (dg_blockjac_supported(g1->P)) {
    OS_C(pdb200_block_jacobi_apply(g1, os->const_residual, XR.dev));   // M^-1 (...)
    scale_kernel<<<grid_for(n), 256, 0, s>>>(XR.dev, -1.0, n);
    os->launches++;
    PDB_CUDA(cudaGetLastError());
  } else {
    scale_kernel<<<grid_for(n), 256, 0, s>>>(os->const_residual, -1.0, n);
    os->launches++;
    PDB_CUDA(cudaGetLastError());
    pdb200_solve_result res;
    OS_C(pdb200_solve(g1, PDB200_SOLVER_CG, PDB200_PRECOND_NONE, nullptr, PDB200_LAYOUT_CSR, XR.dev, os->const_residual,
                      reduction > 0.0 ? reduction : 1e-12, 5000, &res));
    if (!res.converged) throw Error("explicit stage: the mass-matrix solve did not converge");
  }
  XR.copy_back();
  PDB_CUDA(cudaStreamSynchronize(s));
  OS_CATCH
}

int pdb200_onestep_explicit_stage(pdb200_onestep_handle os, int stage, const double* const* x, double* xr, double reduction) {
  if (!x || !xr) {
    pdb_set_last_error("explicit stage: stage must be in 1..s and the vectors given");
    return 1;
  }
  if (int rc = pdb200_onestep_explicit_stage_begin(os, stage)) return rc;
  for (int i = 0; i < stage; i++)
    if (int rc = pdb200_onestep_explicit_stage_add(os, i, x[i])) return rc;
  return pdb200_onestep_explicit_stage_finish(os, xr, reduction);
}

int pdb200_onestep_pre_stage(pdb200_onestep_handle os, int stage, const double* const* x) {
  if (int rc = pdb200_onestep_pre_stage_begin(os, stage)) return rc;
  if (!x) {
    pdb_set_last_error("preStage: null solution list");
    return 1;
  }
  for (int i = 0; i < stage; i++)
    if (int rc = pdb200_onestep_pre_stage_add(os, i, x[i])) return rc;
  return 0;
}

int pdb200_onestep_const_residual(pdb200_onestep_handle os, double* out) {
  OS_TRY
  OS_CHECK(os);
  PDB_CUDA(cudaSetDevice(os->go0->device));
  PDB_CUDA(cudaMemcpyAsync(out, os->const_residual, (size_t)os->go0->P.ndofs * sizeof(double), cudaMemcpyDefault,
                           os->stage->stream));
  PDB_CUDA(cudaStreamSynchronize(os->stage->stream));
  OS_CATCH
}

int pdb200_onestep_stage_operator(pdb200_onestep_handle os, pdb200_handle* stage) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  need_implicit(os, "stage operator");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  combine_stage(os);
  *stage = os->stage;
  OS_CATCH
}

int pdb200_onestep_residual(pdb200_onestep_handle os, const double* x, double* r) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  need_implicit(os, "residual");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  combine_stage(os);
  Staged X(os, &os->hx, x, true), R(os, &os->hr, r, true);
  const long long n = os->go0->P.ndofs;
  cudaStream_t s = os->stage->stream;
  // residualengine.hh:162-176: assemble, add the constant part, then constrain.  The stage operator zeroes the
  // constrained rows and the constant part holds zeros there (it went through the same post-processing).
  OS_C(pdb200_residual(os->stage, X.dev, R.dev));
  axpy_kernel<<<grid_for(n), 256, 0, s>>>(R.dev, os->const_residual, 1.0, n);
  os->launches++;
  PDB_CUDA(cudaGetLastError());
  R.copy_back();
  if (X.host && !R.host) PDB_CUDA(cudaStreamSynchronize(s));
  OS_CATCH
}

int pdb200_onestep_jacobian_apply(pdb200_onestep_handle os, const double* z, double* y) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  PDB_CUDA(cudaSetDevice(os->go0->device));
  combine_stage(os);
  OS_C(pdb200_jacobian_apply(os->stage, z, y));
  OS_CATCH
}

int pdb200_onestep_onthefly_apply(pdb200_onestep_handle os, const double* x, double* y) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  PDB_CUDA(cudaSetDevice(os->go0->device));
  combine_stage(os);
  OS_C(pdb200_onthefly_apply(os->stage, x, y));
  OS_CATCH
}

int pdb200_onestep_jacobian(pdb200_onestep_handle os, const double* x, double* values, int layout) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  need_implicit(os, "jacobian");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  combine_stage(os);
  OS_C(pdb200_jacobian(os->stage, x, values, layout));
  OS_CATCH
}

int pdb200_onestep_solve_stationary(pdb200_onestep_handle os, int solver, int precond, int matrix_free, double* x,
                                    double reduction, double min_defect, uint32_t maxiter, pdb200_solve_result* res) {
  OS_TRY
  OS_CHECK(os);
  need_stage(os);
  need_implicit(os, "StationaryLinearProblemSolver::apply");
  if (!x || !res) throw Error("pdb200_onestep_solve_stationary: null argument");
  PDB_CUDA(cudaSetDevice(os->go0->device));
  combine_stage(os);
  pdb200_operator* st = os->stage;
  const long long n = st->P.ndofs;
  cudaStream_t s = st->stream;
  if (!st->krylov) st->krylov = krylov_create();
  Staged X(os, &os->hx, x, true);
  double *values = nullptr, *r = nullptr, *z = nullptr;
  struct Free3 {
    double *&a, *&b, *&c;
    ~Free3() {
      if (a) cudaFree(a);
      if (b) cudaFree(b);
      if (c) cudaFree(c);
    }
  } guard{values, r, z};
  if (!matrix_free) {  // *_jacobian = 0; igo.jacobian(x, *_jacobian)  (linearproblem.hh:221-226 on onestep.hh:151-159)
    uint64_t nrows = 0, nnz = 0;
    OS_C(pdb200_pattern_size(st, &nrows, &nnz));
    PDB_CUDA(cudaMalloc(&values, nnz * sizeof(double)));
    OS_C(pdb200_jacobian_fresh(st, X.dev, values, PDB200_LAYOUT_CSR));
  }
  PDB_CUDA(cudaMalloc(&r, (size_t)n * sizeof(double)));
  PDB_CUDA(cudaMalloc(&z, (size_t)n * sizeof(double)));
  PDB_CUDA(cudaMemsetAsync(r, 0, (size_t)n * sizeof(double), s));
  PDB_CUDA(cudaMemsetAsync(z, 0, (size_t)n * sizeof(double), s));
  // r = 0; igo.residual(x, r)  (linearproblem.hh:203, 244-246)
  OS_C(pdb200_residual(st, X.dev, r));
  axpy_kernel<<<grid_for(n), 256, 0, s>>>(r, os->const_residual, 1.0, n);
  os->launches++;
  PDB_CUDA(cudaGetLastError());
  const double defect = krylov_two_norm(st->krylov, n, r, s);
  const double red = defect > 0.0 ? std::max(reduction, min_defect / defect) : reduction;  // :212-214
  OS_C(pdb200_solve(st, solver, precond, values, PDB200_LAYOUT_CSR, z, r, red, maxiter, res));
  res->first_defect = defect;
  res->defect = defect * res->reduction;
  krylov_axpy(n, -1.0, z, X.dev, s);  // x -= z  (:289)
  os->launches++;
  X.copy_back();
  PDB_CUDA(cudaStreamSynchronize(s));
  OS_CATCH
}

int pdb200_onestep_launch_count(pdb200_onestep_handle os, uint64_t* n) {
  OS_TRY
  OS_CHECK(os);
  *n = os->launches + os->stage->launches;
  OS_CATCH
}

}  // extern "C"

// operator.cu — the C ABI of include/pdelab_b200.h: operator handle, coefficient upload, DOF
// numbering queries, host/device pointer staging and kernel dispatch.
//
// Mirrors Dune::PDELab::GridOperator (gridoperator/gridoperator.hh:30-244): the handle owns what
// the reference's GridOperator references (function-space sizes, constraints, local-operator
// parameters) and exposes residual / jacobian_apply / jacobian / fill_pattern.

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "operator.h"

namespace {
thread_local std::string g_last_error;
}
void pdb_set_last_error(const std::string& s) { g_last_error = s; }  // for onestep.cu

// R(0) of the affine DG residual with the reference-order kernel, cached per coefficient set (used by the
// Kronecker kernels' residual form and, scaled, by the one-step stage operator)
void pdb_ensure_r0(pdb200_operator* op) {
  if (op->r0_valid) return;
  const DevParams& P = op->P;
  if (!op->r0) PDB_CUDA(cudaMalloc(&op->r0, (size_t)P.ndofs * sizeof(double)));
  double* zero = nullptr;
  PDB_CUDA(cudaMalloc(&zero, (size_t)P.ndofs * sizeof(double)));
  PDB_CUDA(cudaMemsetAsync(zero, 0, (size_t)P.ndofs * sizeof(double), op->stream));
  launch_dg_generic(P, zero, op->r0, /*residual=*/true, /*overwrite=*/true, op->errflag, op->stream);
  PDB_CUDA(cudaStreamSynchronize(op->stream));
  PDB_CUDA(cudaFree(zero));
  op->r0_valid = true;
  op->launches += 1;
}
bool pdb_uses_cached_r0(const pdb200_operator* op) {
  const DevParams& P = op->P;
  return P.dg && op->kernel_choice != PDB200_KERNEL_GENERIC &&
         (dg_fast_supported(P) || dg_kron_supported(P) || dg_small_supported(P));
}


namespace {

bool is_device_pointer(const void* p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

template <typename T>
const T* upload(pdb200_operator* op, const T* src, size_t count) {
  if (!src || count == 0) return nullptr;
  T* d = nullptr;
  PDB_CUDA(cudaMalloc(&d, count * sizeof(T)));
  op->owned.push_back(d);
  PDB_CUDA(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyDefault));
  return d;
}

// array lengths in the cell-wise or point-wise layout (header of pdelab_b200.h)
size_t a_count(const DevParams& P) {
  const size_t e = (size_t)P.ncells * ((P.pw & PDB200_POINTWISE_A) ? P.np : 1);
  switch (P.a_mode) {
    case PDB200_A_IDENTITY: return 0;
    case PDB200_A_SCALAR: return e;
    case PDB200_A_DIAGONAL: return e * P.dim;
    default: return e * P.dim * P.dim;
  }
}
size_t b_count(const DevParams& P) { return (size_t)P.ncells * ((P.pw & PDB200_POINTWISE_B) ? P.np : 1) * P.dim; }
size_t c_count(const DevParams& P) { return (size_t)P.ncells * ((P.pw & PDB200_POINTWISE_C) ? P.nq : 1); }

long long num_bfaces(const DevParams& P) {
  long long n = 0;
  for (int d = 0; d < P.dim; d++) n += 2 * (P.ncells / P.N[d]);
  return n;
}

void ensure_device(pdb200_operator* op) { PDB_CUDA(cudaSetDevice(op->device)); }

void check_errflag(pdb200_operator* op) {
  int flag = 0;
  PDB_CUDA(cudaMemcpyAsync(&flag, op->errflag, sizeof(int), cudaMemcpyDeviceToHost, op->stream));
  PDB_CUDA(cudaStreamSynchronize(op->stream));
  if (flag) {
    PDB_CUDA(cudaMemsetAsync(op->errflag, 0, sizeof(int), op->stream));
    throw Error("Outflow boundary condition on inflow!");  // convectiondiffusiondg.hh:802-806
  }
}

enum class Mode { Residual, JacobianApply, OnTheFly };

// run one vector kernel on device pointers
void run_vector_device(pdb200_operator* op, const double* x, double* y, Mode mode, int part = PDB200_PART_ALL,
                       cudaStream_t stream_override = nullptr) {
  const DevParams& P = op->P;
  struct StreamGuard {  // the BOUNDARY part of apply_p2p runs on the side stream
    pdb200_operator* op;
    cudaStream_t saved;
    ~StreamGuard() { op->stream = saved; }
  } guard{op, op->stream};
  if (stream_override) op->stream = stream_override;
  const bool residual = mode == Mode::Residual;
  const bool overwrite = mode == Mode::OnTheFly;
  bool use_fast0 = P.dg && dg_fast_supported(P) && op->kernel_choice != PDB200_KERNEL_GENERIC;
  if (part != PDB200_PART_ALL && !use_fast0) {
    // kernels without a tile decomposition: everything runs in the BOUNDARY phase (after the exchange)
    if (part == PDB200_PART_INTERIOR) return;
    part = PDB200_PART_ALL;
  }
  if (!P.dg) {
    launch_fem_vector(op->fem, P, x, y, residual, overwrite, op->stream);
    const bool kron = kron_coefficients(P);
    op->last_kernel = kron ? (residual ? "fem_kron+r0" : "fem_kron") : (residual ? "fem_residual" : "fem_jacobian_apply");
    op->launches += 2;
    return;
  }
  bool use_fast = use_fast0;
  const bool use_kron = !use_fast && dg_kron_supported(P) && op->kernel_choice != PDB200_KERNEL_GENERIC;
  const bool use_small = !use_fast && !use_kron && dg_small_supported(P) && op->kernel_choice != PDB200_KERNEL_GENERIC;
  if (op->kernel_choice == PDB200_KERNEL_FAST && !use_fast && !use_kron && !use_small)
    throw Error("PDB200_KERNEL_FAST requested but the configuration has no fast kernel "
                "(needs QkDG with cell-wise constant diagonal A: dim=3 with k=2 and even cells[0] (b allowed), or with "
                "b=0: dim=3 with k in {1,3,4}, dim=2 with k in {1,2})");
  if (use_fast || use_kron || use_small) {
    const double* r0 = nullptr;
    if (residual) {
      // The operator is affine: R(x) = J x + R(0).  R(0) (source term lambda_volume,
      // convectiondiffusiondg.hh:1048-1075, and the boundary data g, j, o, :684-879) is evaluated
      // once per coefficient set with the reference-order kernel and cached.
      pdb_ensure_r0(op);
      r0 = op->r0;
    }
    if (use_fast) {
      if (!op->fast) op->fast = dg_fast_plan_create(P, op->K);
      op->launches += launch_dg_fast(op->fast, P, x, y, r0, overwrite, part, op->stream, op->errflag);
      op->last_kernel = residual ? "dg_fast_q2_3d+r0" : "dg_fast_q2_3d";
    } else if (use_kron) {
      if (!op->kron) op->kron = dg_kron_plan_create(P, op->K);
      op->launches += launch_dg_kron(op->kron, P, x, y, r0, overwrite, op->stream);
      op->last_kernel = residual ? "dg_kron_3d+r0" : "dg_kron_3d";
    } else {
      op->launches += launch_dg_small(P, op->K, x, y, r0, overwrite, op->stream, op->errflag);
      op->last_kernel = residual ? "dg_small+r0" : "dg_small";
    }
  } else {
    launch_dg_generic(P, x, y, residual, overwrite, op->errflag, op->stream);
    op->last_kernel = residual ? "dg_generic_residual" : "dg_generic_jacobian_apply";
    op->launches += 1;
  }
}


// y = J x on the overlapping partition: exchange of x's ghost layers hidden behind the interior tiles
void apply_p2p_device(pdb200_operator* h, double* x, double* y) {
  if (p2p_is_qk(h->p2p)) {
    // conforming Qk: the lattice planes travel direction by direction (corners in two / three hops), then every rank
    // evaluates complete rows for the closure of its owned cells; no tile split to hide the exchange behind
    h->launches += p2p_exchange(h->p2p, h->P, x, h->stream);
    run_vector_device(h, x, y, Mode::OnTheFly);
    return;
  }
  {
    // QkDG k = 2 in 3-D with tile-aligned owned extents: ONE launch per step — push blocks, interior tiles, then the
    // tiles next to a processor side, which read the neighbour's layer straight from the mailbox (no unpack pass, no
    // side stream, no event fork / join).  PDB200_P2P_FUSED=0 keeps the multi-launch schedule below.
    static const bool fused_on = [] {
      const char* e = getenv("PDB200_P2P_FUSED");
      return !(e && e[0] == '0');
    }();
    const DevParams& P = h->P;
    if (fused_on && P.dg && h->kernel_choice != PDB200_KERNEL_GENERIC && dg_fast_fused_supported(P)) {
      const FusedTable* table = p2p_fused_table(h->p2p, P);
      if (table) {
        if (!h->fast) h->fast = dg_fast_plan_create(P, h->K);
        h->launches += launch_dg_fast_fused(h->fast, P, x, y, table, p2p_fused_table_host(h->p2p), p2p_next_epoch(h->p2p),
                                            h->stream, h->errflag);
        h->last_kernel = "dg_fast_q2_3d+halo";
        return;
      }
    }
  }
  cudaStream_t side = p2p_stream(h->p2p);
  PDB_CUDA(cudaEventRecord(p2p_event(h->p2p, 0), h->stream));  // x is ready
  PDB_CUDA(cudaStreamWaitEvent(side, p2p_event(h->p2p, 0), 0));
  h->launches += p2p_push(h->p2p, h->P, x, side);
  h->launches += p2p_wait_unpack(h->p2p, h->P, x, side);
  // the boundary tiles follow the unpack on the (high-priority) side stream, so they fill in
  // beside the last interior tiles instead of waiting for them; the outputs are disjoint
  run_vector_device(h, x, y, Mode::OnTheFly, PDB200_PART_BOUNDARY, side);
  PDB_CUDA(cudaEventRecord(p2p_event(h->p2p, 1), side));
  run_vector_device(h, x, y, Mode::OnTheFly, PDB200_PART_INTERIOR);
  PDB_CUDA(cudaStreamWaitEvent(h->stream, p2p_event(h->p2p, 1), 0));
}

// y = J x with HOST vectors through the fast kernel: the vector is cut into windows of tile layers
// along z; window c is computed as soon as its input layers (and one layer of window c+1) have
// arrived, and its result travels back while the next window is computed.  PCIe is full duplex,
// so the call costs about max(H2D, D2H) instead of H2D + kernel + D2H.
bool run_onthefly_host_pipelined(pdb200_operator* op, const double* x, double* y) {
  const DevParams& P = op->P;
  if (!(P.dg && dg_fast_supported(P) && op->kernel_choice != PDB200_KERNEL_GENERIC)) return false;
  const int nzt = dg_fast_ztiles(P);
  const int nwin = std::min(16, nzt);
  if (nwin < 2) return false;
  const size_t bytes = (size_t)P.ndofs * sizeof(double);
  if (!op->dx) PDB_CUDA(cudaMalloc(&op->dx, bytes));
  if (!op->dy) PDB_CUDA(cudaMalloc(&op->dy, bytes));
  if (!op->fast) op->fast = dg_fast_plan_create(P, op->K);
  if (!op->h2d_stream) {
    PDB_CUDA(cudaStreamCreateWithFlags(&op->h2d_stream, cudaStreamNonBlocking));
    PDB_CUDA(cudaStreamCreateWithFlags(&op->d2h_stream, cudaStreamNonBlocking));
    for (auto& e : op->pipe_ev) PDB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const size_t layer = (size_t)P.N[0] * P.N[1] * P.n;  // doubles per cell layer
  cudaEvent_t* ev_in = op->pipe_ev;                     // [16]
  cudaEvent_t* ev_out = op->pipe_ev + 16;               // [16]
  cudaEvent_t ev_start = op->pipe_ev[32];
  PDB_CUDA(cudaEventRecord(ev_start, op->stream));      // order after earlier work on the handle's stream
  PDB_CUDA(cudaStreamWaitEvent(op->h2d_stream, ev_start, 0));
  PDB_CUDA(cudaStreamWaitEvent(op->d2h_stream, ev_start, 0));
  int zin0[16], zin1[16], tlo[16], thi[16];
  for (int c = 0; c < nwin; c++) {
    tlo[c] = (int)((long long)nzt * c / nwin);
    thi[c] = (int)((long long)nzt * (c + 1) / nwin);
    dg_fast_ztile_layers(P, tlo[c], thi[c], &zin0[c], &zin1[c]);
  }
  for (int c = 0; c < nwin; c++) {  // all uploads are queued up front: the copy engine streams them back to back
    PDB_CUDA(cudaMemcpyAsync(op->dx + zin0[c] * layer, x + zin0[c] * layer, (zin1[c] - zin0[c]) * layer * sizeof(double),
                             cudaMemcpyHostToDevice, op->h2d_stream));
    PDB_CUDA(cudaEventRecord(ev_in[c], op->h2d_stream));
  }
  for (int c = 0; c < nwin; c++) {
    PDB_CUDA(cudaStreamWaitEvent(op->stream, ev_in[std::min(c + 1, nwin - 1)], 0));  // needs one layer of the next window
    op->launches += launch_dg_fast(op->fast, P, op->dx, op->dy, nullptr, true, PDB200_PART_ALL, op->stream, op->errflag, tlo[c], thi[c]);
    PDB_CUDA(cudaEventRecord(ev_out[c], op->stream));
    PDB_CUDA(cudaStreamWaitEvent(op->d2h_stream, ev_out[c], 0));
    PDB_CUDA(cudaMemcpyAsync(y + zin0[c] * layer, op->dy + zin0[c] * layer, (zin1[c] - zin0[c]) * layer * sizeof(double),
                             cudaMemcpyDeviceToHost, op->d2h_stream));
  }
  op->last_kernel = "dg_fast_q2_3d";
  PDB_CUDA(cudaStreamSynchronize(op->d2h_stream));
  PDB_CUDA(cudaStreamSynchronize(op->stream));
  return true;
}

void run_vector(pdb200_operator* op, const double* x, double* y, Mode mode) {
  ensure_device(op);
  const DevParams& P = op->P;
  const bool xd = is_device_pointer(x), yd = is_device_pointer(y);
  const size_t bytes = (size_t)P.ndofs * sizeof(double);
  if (!xd && !yd && mode == Mode::OnTheFly && run_onthefly_host_pipelined(op, x, y)) return;
  const double* xdev = x;
  double* ydev = y;
  if (!xd) {
    if (!op->dx) PDB_CUDA(cudaMalloc(&op->dx, bytes));
    PDB_CUDA(cudaMemcpyAsync(op->dx, x, bytes, cudaMemcpyHostToDevice, op->stream));
    xdev = op->dx;
  }
  if (!yd) {
    if (!op->dy) PDB_CUDA(cudaMalloc(&op->dy, bytes));
    if (mode != Mode::OnTheFly) PDB_CUDA(cudaMemcpyAsync(op->dy, y, bytes, cudaMemcpyHostToDevice, op->stream));
    ydev = op->dy;
  }
  run_vector_device(op, xdev, ydev, mode);
  if (!yd) {
    PDB_CUDA(cudaMemcpyAsync(y, op->dy, bytes, cudaMemcpyDeviceToHost, op->stream));
  }
  if (!xd || !yd) {
    // host-visible results: synchronous on return like the reference (SURVEY.md §8b "Threading")
    if (P.dg && (P.bctype != nullptr) && (P.b != nullptr))
      check_errflag(op);
    else
      PDB_CUDA(cudaStreamSynchronize(op->stream));
  }
}

}  // namespace

#define PDB_TRY try {
#define PDB_CATCH                    \
  }                                  \
  catch (const std::exception& e) {  \
    g_last_error = e.what();         \
    return 1;                        \
  }                                  \
  return 0;
#define PDB_CHECK_HANDLE(h) \
  if (!(h)) throw Error("null operator handle")

extern "C" {

const char* pdb200_last_error(void) { return g_last_error.c_str(); }
const char* pdb200_version(void) { return "pdelab_b200 0.1 (sm_100a)"; }

int pdb200_create(const pdb200_problem* p, pdb200_handle* out) {
  PDB_TRY
  if (!p || !out) throw Error("null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw Error("no CUDA device available: pdelab_b200 has no CPU fallback");
  if (p->device < 0 || p->device >= ndev) throw Error("invalid CUDA device ordinal");
  if (p->dim != 2 && p->dim != 3) throw Error("dim must be 2 or 3");
  if (p->space != PDB200_SPACE_QKDG && p->space != PDB200_SPACE_QK) throw Error("unknown finite element space");
  if (p->degree < 1 || p->degree > MAX_K) throw Error("degree must be in 1..4");
  if (p->space == PDB200_SPACE_QK && p->degree > 2)
    throw Error("conforming Qk is available for k in {1,2} (finiteelementmap/qkfem.hh:17-78)");
  if (p->intorderadd < 0 || p->intorderadd > 1)
    throw Error("intorderadd must be 0 or 1 (k+1 Gauss points per direction)");
  std::unique_ptr<pdb200_operator> op(new pdb200_operator);
  op->device = p->device;
  op->kernel_choice = p->kernel;
  PDB_CUDA(cudaSetDevice(op->device));
  DevParams& P = op->P;
  std::memset(&P, 0, sizeof(P));
  P.dim = p->dim;
  P.k = p->degree;
  P.n1 = P.k + 1;
  P.dg = p->space == PDB200_SPACE_QKDG;
  P.basis = p->basis;
  if (P.basis < PDB200_BASIS_LAGRANGE || P.basis > PDB200_BASIS_LOBATTO) throw Error("unknown QkDG basis");
  if (!P.dg && P.basis != PDB200_BASIS_LAGRANGE)
    throw Error("conforming Qk spaces use the Lagrange basis (finiteelementmap/qkfem.hh)");
  P.m = (2 * P.k + p->intorderadd) / 2 + 1;  // convectiondiffusiondg.hh:139, convectiondiffusionfem.hh:93
  P.n = P.nq = P.nfq = 1;
  P.ncells = 1;
  P.vol = 1.0;
  for (int d = 0; d < 3; d++) {
    P.N[d] = d < P.dim ? p->cells[d] : 1;
    op->lower[d] = p->lower[d];
    op->upper[d] = p->upper[d];
    if (P.N[d] < 1) throw Error("cells must be positive");
    P.h[d] = d < P.dim ? (p->upper[d] - p->lower[d]) / P.N[d] : 1.0;
    if (!(P.h[d] > 0.0)) throw Error("upper must be greater than lower");
    P.ih[d] = 1.0 / P.h[d];
    if (d < P.dim) {
      P.n *= P.n1;
      P.nq *= P.m;
      if (d > 0) P.nfq *= P.m;
      P.ncells *= P.N[d];
      P.vol *= P.h[d];
    }
    for (int s = 0; s < 2; s++) P.side_kind[d][s] = p->side_kind[d][s];
  }
  for (int d = 0; d < 3; d++) {
    P.area[d] = 1.0;
    for (int e = 0; e < P.dim; e++)
      if (e != d) P.area[d] *= P.h[e];
  }
  long long nbf = 0;
  for (int d = 0; d < P.dim; d++)
    for (int s = 0; s < 2; s++) {
      P.bf_off[d][s] = nbf;
      nbf += P.ncells / P.N[d];
    }
  P.ndofs = host_num_dofs(P);
  P.theta = 1.0;  // convectiondiffusiondg.hh:99-101
  if (p->dg_method == PDB200_DG_SIPG) P.theta = -1.0;
  if (p->dg_method == PDB200_DG_IIPG) P.theta = 0.0;
  P.alpha = p->dg_alpha;
  P.weights_on = p->dg_weights == PDB200_DG_WEIGHTS_ON;
  P.a_mode = p->a_mode;
  if (P.a_mode < 0 || P.a_mode > 3) throw Error("invalid a_mode");
  if (P.a_mode != PDB200_A_IDENTITY && !p->A) throw Error("a_mode needs the array A");
  host_fill_tables(P, op->K, op->xq, op->wq);
  P.np = P.nq + 2 * P.dim * P.nfq;
  // point-wise layouts only for arrays that are present (a bit without its array means nothing)
  P.pw = 0;
  if ((p->pointwise & PDB200_POINTWISE_A) && P.a_mode != PDB200_A_IDENTITY) P.pw |= PDB200_POINTWISE_A;
  if ((p->pointwise & PDB200_POINTWISE_B) && p->b) P.pw |= PDB200_POINTWISE_B;
  if ((p->pointwise & PDB200_POINTWISE_C) && p->c) P.pw |= PDB200_POINTWISE_C;
  if ((p->pointwise & PDB200_POINTWISE_BCTYPE) && p->bctype) {
    if (!P.dg)
      throw Error("ConvectionDiffusionFEM evaluates bctype at the face centre (convectiondiffusionfem.hh:226-229): "
                  "PDB200_POINTWISE_BCTYPE is a QkDG layout");
    P.pw |= PDB200_POINTWISE_BCTYPE;
  }
  if (p->pointwise & ~(PDB200_POINTWISE_A | PDB200_POINTWISE_B | PDB200_POINTWISE_C | PDB200_POINTWISE_BCTYPE))
    throw Error("unknown bits in pdb200_problem::pointwise");
  P.A = upload(op.get(), p->A, a_count(P));
  P.b = upload(op.get(), p->b, b_count(P));
  P.c = upload(op.get(), p->c, c_count(P));
  P.f = upload(op.get(), p->f, (size_t)P.ncells * P.nq);
  P.bctype = upload(op.get(), p->bctype, (size_t)nbf * ((P.pw & PDB200_POINTWISE_BCTYPE) ? P.nfq : 1));
  P.g = upload(op.get(), p->g, (size_t)nbf * P.nfq);
  P.j = upload(op.get(), p->j, (size_t)nbf * P.nfq);
  P.o = upload(op.get(), p->o, (size_t)nbf * P.nfq);
  PDB_CUDA(cudaMalloc(&op->errflag, sizeof(int)));
  PDB_CUDA(cudaMemset(op->errflag, 0, sizeof(int)));
  if (!P.dg) op->fem = fem_plan_create(P, p->bctype ? op->P.bctype : nullptr, op->K);
  *out = op.release();
  PDB_CATCH
}

int pdb200_destroy(pdb200_handle h) {
  PDB_TRY
  delete h;
  PDB_CATCH
}

int pdb200_update_coefficients(pdb200_handle h, const pdb200_problem* p) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  DevParams& P = h->P;
  const long long nbf = num_bfaces(P);
  auto upd = [&](const void* dst, const void* src, size_t bytes, const char* name) {
    if (!src) return;
    if (!dst) throw Error(std::string("update_coefficients: array was not given at create time: ") + name);
    PDB_CUDA(cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyDefault, h->stream));
  };
  upd(P.A, p->A, a_count(P) * 8, "A");
  upd(P.b, p->b, b_count(P) * 8, "b");
  upd(P.c, p->c, c_count(P) * 8, "c");
  upd(P.f, p->f, (size_t)P.ncells * P.nq * 8, "f");
  upd(P.g, p->g, (size_t)nbf * P.nfq * 8, "g");
  upd(P.j, p->j, (size_t)nbf * P.nfq * 8, "j");
  upd(P.o, p->o, (size_t)nbf * P.nfq * 8, "o");
  if (p->bctype) throw Error("update_coefficients: bctype changes the constraint set; create a new operator");
  h->r0_valid = false;
  h->coeff_version++;
  dg_blockjac_invalidate(h->blockjac);
  fem_plan_invalidate(h->fem);
  PDB_CUDA(cudaStreamSynchronize(h->stream));
  PDB_CATCH
}

int pdb200_num_dofs(pdb200_handle h, uint64_t* n) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  *n = (uint64_t)h->P.ndofs;
  PDB_CATCH
}
int pdb200_local_size(pdb200_handle h, uint32_t* n) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  *n = (uint32_t)h->P.n;
  PDB_CATCH
}
int pdb200_num_boundary_faces(pdb200_handle h, uint64_t* n) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  *n = (uint64_t)num_bfaces(h->P);
  PDB_CATCH
}
int pdb200_boundary_face_offset(pdb200_handle h, int dir, int side, uint64_t* first) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  if (dir < 0 || dir >= h->P.dim || side < 0 || side > 1) throw Error("invalid (dir, side)");
  *first = (uint64_t)h->P.bf_off[dir][side];
  PDB_CATCH
}
int pdb200_quadrature_size(pdb200_handle h, uint32_t* m) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  *m = (uint32_t)h->P.m;
  PDB_CATCH
}
int pdb200_quadrature(pdb200_handle h, double* points, double* weights) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  for (int i = 0; i < h->P.m; i++) {
    points[i] = h->xq[i];
    weights[i] = h->wq[i];
  }
  PDB_CATCH
}

int pdb200_gauss_legendre(int m, double* points, double* weights) {
  PDB_TRY
  if (m < 1 || m > 16 || !points || !weights) throw Error("gauss_legendre: 1 <= m <= 16 and non-null outputs");
  std::vector<long double> x, w;
  host_gauss(m, x, w);
  for (int i = 0; i < m; i++) {
    points[i] = (double)x[i];
    weights[i] = (double)w[i];
  }
  PDB_CATCH
}

int pdb200_cell_dof_indices(pdb200_handle h, uint64_t cell, uint64_t* idx) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  if (cell >= (uint64_t)h->P.ncells) throw Error("cell index out of range");
  host_cell_dof_indices(h->P, (long long)cell, idx);
  PDB_CATCH
}

int pdb200_constrained_dofs(pdb200_handle h, uint64_t* count, uint64_t* idx) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  std::vector<int8_t> bct;
  if (h->P.bctype) {
    bct.resize(num_bfaces(h->P));
    PDB_CUDA(cudaMemcpy(bct.data(), h->P.bctype, bct.size(), cudaMemcpyDeviceToHost));
  }
  std::vector<uint64_t> list = host_constrained_dofs(h->P, bct.empty() ? nullptr : bct.data());
  *count = list.size();
  if (idx) std::memcpy(idx, list.data(), list.size() * sizeof(uint64_t));
  PDB_CATCH
}

int pdb200_residual(pdb200_handle h, const double* x, double* r) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  run_vector(h, x, r, Mode::Residual);
  PDB_CATCH
}

int pdb200_jacobian_apply(pdb200_handle h, const double* z, double* y) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  run_vector(h, z, y, Mode::JacobianApply);
  PDB_CATCH
}

int pdb200_onthefly_apply(pdb200_handle h, const double* x, double* y) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  run_vector(h, x, y, Mode::OnTheFly);
  PDB_CATCH
}

int pdb200_jacobian_apply_nonlinear(pdb200_handle h, const double*, const double*, double*) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  // gridoperator/gridoperator.hh:202-203
  throw Error("jacobian_apply(u,z,y) with linearisation point called for a linear local operator");
  PDB_CATCH
}

int pdb200_pattern_size(pdb200_handle h, uint64_t* nrows, uint64_t* nnz) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->matrix) h->matrix = matrix_plan_create(h->P, h->fem, h->stream);
  matrix_pattern_size(h->matrix, 0, nrows, nnz);
  PDB_CATCH
}
int pdb200_block_pattern_size(pdb200_handle h, uint64_t* nrows, uint64_t* nnz) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->P.dg) throw Error("block pattern is defined for QkDG spaces (Blocking::fixed)");
  if (!h->matrix) h->matrix = matrix_plan_create(h->P, h->fem, h->stream);
  matrix_pattern_size(h->matrix, 1, nrows, nnz);
  PDB_CATCH
}

static void pattern_out(pdb200_handle h, int layout, void* rowptr, void* colidx, bool col32) {
  ensure_device(h);
  if (!h->matrix) h->matrix = matrix_plan_create(h->P, h->fem, h->stream);
  h->launches += matrix_pattern_write(h->matrix, layout, rowptr, is_device_pointer(rowptr), colidx,
                                      is_device_pointer(colidx), col32, h->stream);
}

int pdb200_pattern(pdb200_handle h, uint64_t* rowptr, uint64_t* colidx) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  pattern_out(h, 0, rowptr, colidx, false);
  PDB_CATCH
}
int pdb200_pattern_i32(pdb200_handle h, uint64_t* rowptr, uint32_t* colidx) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  pattern_out(h, 0, rowptr, colidx, true);
  PDB_CATCH
}
int pdb200_block_pattern(pdb200_handle h, uint64_t* rowptr, uint64_t* colidx) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  if (!h->P.dg) throw Error("block pattern is defined for QkDG spaces (Blocking::fixed)");
  pattern_out(h, 1, rowptr, colidx, false);
  PDB_CATCH
}

static void jacobian_impl(pdb200_handle h, double* values, int layout, bool fresh) {
  ensure_device(h);
  if (layout != PDB200_LAYOUT_CSR && layout != PDB200_LAYOUT_BCSR) throw Error("unknown matrix layout");
  if (layout == PDB200_LAYOUT_BCSR && !h->P.dg) throw Error("BCSR layout is defined for QkDG spaces");
  if (!h->matrix) h->matrix = matrix_plan_create(h->P, h->fem, h->stream);
  h->launches += matrix_assemble(h->matrix, layout, values, is_device_pointer(values), fresh, h->errflag, h->stream);
  if (h->P.dg && h->P.bctype && h->P.b) check_errflag(h);
}

int pdb200_jacobian(pdb200_handle h, const double* x, double* values, int layout) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  (void)x;  // both local operators are linear: the Jacobian does not depend on x
  jacobian_impl(h, values, layout, false);
  PDB_CATCH
}

int pdb200_jacobian_fresh(pdb200_handle h, const double* x, double* values, int layout) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  (void)x;
  jacobian_impl(h, values, layout, true);
  PDB_CATCH
}

int pdb200_csr_mv(pdb200_handle h, const double* values, int layout, const double* x, double* y) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->matrix) h->matrix = matrix_plan_create(h->P, h->fem, h->stream);
  if (!is_device_pointer(values) || !is_device_pointer(x) || !is_device_pointer(y))
    throw Error("csr_mv expects device pointers");
  h->launches += matrix_mv(h->matrix, layout, values, x, y, h->stream);
  PDB_CATCH
}

namespace {

// device staging of a host vector for the solver entry points
struct Staged {
  double* dev = nullptr;
  double* host = nullptr;
  size_t bytes = 0;
  bool owned = false;
  Staged(double* p, size_t n, bool copy_in, cudaStream_t s) : host(p), bytes(n * sizeof(double)) {
    if (is_device_pointer(p)) {
      dev = p;
    } else {
      PDB_CUDA(cudaMalloc(&dev, bytes));
      owned = true;
      if (copy_in) PDB_CUDA(cudaMemcpyAsync(dev, p, bytes, cudaMemcpyHostToDevice, s));
    }
  }
  void copy_out(cudaStream_t s) {
    if (owned) PDB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s));
  }
  ~Staged() {
    if (owned) cudaFree(dev);
  }
};

void point_diagonal_device(pdb200_operator* h, double* d) {
  if (h->P.dg) {
    h->launches += launch_dg_diagonal(h->P, h->K, d, h->stream);
  } else {
    launch_fem_diagonal(h->fem, h->P, d, h->stream);
    h->launches += 2;
  }
}

// binds the operator (matrix-free or assembled) and the preconditioner, then runs the Krylov loop
void solve_device(pdb200_operator* h, int solver, int precond, const double* values, int layout, double* z, double* r,
                  double reduction, uint32_t maxiter, pdb200_solve_result* res, bool ovlp = false) {
  const DevParams& P = h->P;
  if (!h->krylov) h->krylov = krylov_create();
  KrylovOps ops;
  double* dinv = nullptr;
  struct Free {
    double*& p;
    ~Free() {
      if (p) cudaFree(p);
    }
  } free_dinv{dinv};
  double* staged_values = nullptr;
  Free free_values{staged_values};
  if (!values) {
    if (precond == PDB200_PRECOND_BLOCK_JACOBI) {
      if (!h->blockjac) h->blockjac = dg_blockjac_create(P, h->K);
      ops.prec = [h](const double* in, double* out) {
        h->launches += launch_dg_blockjac(h->blockjac, h->P, h->K, in, out, h->stream);
      };
    } else if (precond == PDB200_PRECOND_BLOCK_SOR || precond == PDB200_PRECOND_BLOCK_SSOR) {
      // BlockSORPreconditionerLocalOperator (backend/istl/matrixfree/blocksorpreconditioner.hh) through
      // GridOperatorPreconditioner::apply: v = 0, one forward sweep (SSOR: forward then backward, symmetric for CG)
      if (!h->blockjac) h->blockjac = dg_blockjac_create(P, h->K);
      const bool sym = precond == PDB200_PRECOND_BLOCK_SSOR;
      ops.prec = [h, sym](const double* in, double* out) {
        h->launches += launch_dg_blocksor(h->blockjac, h->P, h->K, in, out, h->relaxation, false, true, h->stream);
        if (sym) h->launches += launch_dg_blocksor(h->blockjac, h->P, h->K, in, out, h->relaxation, true, false, h->stream);
      };
    } else if (precond == PDB200_PRECOND_JACOBI) {  // point Jacobi on the matrix-free diagonal
      PDB_CUDA(cudaMalloc(&dinv, (size_t)P.ndofs * sizeof(double)));
      point_diagonal_device(h, dinv);
      krylov_invert(P.ndofs, dinv, h->stream);
      h->launches += 1;
      ops.dinv = dinv;
    } else if (precond != PDB200_PRECOND_NONE) {
      throw Error("pdb200_solve: unknown preconditioner");
    }
    ops.apply = [h](const double* in, double* out) { run_vector_device(h, in, out, Mode::OnTheFly); };
  } else {
    if (!h->matrix) h->matrix = matrix_plan_create(P, h->fem, h->stream);
    if (!is_device_pointer(values)) {  // host container (the C++ mirror's BCRSMatrixContainer): stage it
      uint64_t nr = 0, nnz = 0;
      matrix_pattern_size(h->matrix, layout, &nr, &nnz);
      const size_t count = layout == PDB200_LAYOUT_BCSR ? (size_t)nnz * P.n * P.n : (size_t)nnz;
      PDB_CUDA(cudaMalloc(&staged_values, count * sizeof(double)));
      PDB_CUDA(cudaMemcpyAsync(staged_values, values, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      values = staged_values;
    }
    ops.apply = [h, values, layout](const double* in, double* out) {
      h->launches += matrix_mv(h->matrix, layout, values, in, out, h->stream);
    };
    if (precond == PDB200_PRECOND_JACOBI) {
      if (layout != PDB200_LAYOUT_CSR) throw Error("pdb200_solve: Jacobi needs the scalar CSR layout");
      uint64_t nrows = 0, nnz = 0;
      matrix_pattern_size(h->matrix, PDB200_LAYOUT_CSR, &nrows, &nnz);
      uint64_t* rowptr = nullptr;
      uint32_t* colidx = nullptr;
      PDB_CUDA(cudaMalloc(&rowptr, (nrows + 1) * sizeof(uint64_t)));
      PDB_CUDA(cudaMalloc(&colidx, nnz * sizeof(uint32_t)));
      PDB_CUDA(cudaMalloc(&dinv, nrows * sizeof(double)));
      h->launches += matrix_pattern_write(h->matrix, PDB200_LAYOUT_CSR, rowptr, true, colidx, true, true, h->stream);
      krylov_diag_inverse((long long)nrows, rowptr, colidx, values, dinv, h->stream);
      h->launches += 1;
      PDB_CUDA(cudaStreamSynchronize(h->stream));
      cudaFree(rowptr);
      cudaFree(colidx);
      ops.dinv = dinv;
    } else if (precond != PDB200_PRECOND_NONE) {
      throw Error("pdb200_solve: unknown preconditioner");
    }
  }
  if (ovlp) {
    // OverlappingOperator + OverlappingScalarProduct (backend/istl/ovlpistlsolverbackend.hh:40-134).  Vectors are kept
    // in the unique representation (ghost layers zero) outside the operator: the apply makes its input consistent
    // (owner -> ghost copy over the NVLink mailboxes, hidden behind the interior tiles), evaluates the local rows,
    // zeroes the ghost rows of the result (set_constrained_dofs(cc, 0.0, y)) and drops the input's ghosts again, so
    // every inner product is the disjoint dot product without a mask; the global sums go through the peer mailboxes.
    if (!h->p2p || !h->comm)
      throw Error("pdb200_solve_ovlp: call pdb200_halo_p2p_create / _connect and pdb200_comm_create / _connect first");
    // Conforming Qk: "ghost" = every lattice point the rank does not own (the planes it receives, including the
    // interface plane towards a lower neighbour); the rows of owned points are complete because the ghost cell layer
    // supplies all adjacent cells and the boundary of the extended box is constrained.
    if (!values) {
      ops.apply = [h](const double* in, double* out) {
        apply_p2p_device(h, const_cast<double*>(in), out);
        h->launches += p2p_zero_ghosts(h->p2p, h->P, out, h->stream);
        h->launches += p2p_zero_ghosts(h->p2p, h->P, const_cast<double*>(in), h->stream);
      };
    } else {
      auto mv = ops.apply;
      ops.apply = [h, mv](const double* in, double* out) {
        h->launches += p2p_exchange(h->p2p, h->P, const_cast<double*>(in), h->stream);
        mv(in, out);
        h->launches += p2p_zero_ghosts(h->p2p, h->P, out, h->stream);
        h->launches += p2p_zero_ghosts(h->p2p, h->P, const_cast<double*>(in), h->stream);
      };
    }
    ops.allreduce = [h](double* P1, double* P2) {
      comm_allreduce_partials(h->comm, P1, P2, krylov_partial_count(), h->stream);
    };
    h->launches += p2p_zero_ghosts(h->p2p, P, r, h->stream);
    if (ops.dinv && !P.dg) h->launches += p2p_zero_ghosts(h->p2p, P, const_cast<double*>(ops.dinv), h->stream);
  }
  h->launches += krylov_solve(h->krylov, solver, P.ndofs, ops, z, r, reduction, maxiter, h->stream, res);
  if (ovlp) {
    // hand back a consistent solution (the reference's vectors are consistent on the whole overlap)
    h->launches += p2p_exchange(h->p2p, h->P, z, h->stream);
    PDB_CUDA(cudaStreamSynchronize(h->stream));
    p2p_check(h->p2p);
    comm_check(h->comm);
  }
}

}  // namespace

int pdb200_point_diagonal(pdb200_handle h, double* d) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!d) throw Error("pdb200_point_diagonal: null argument");
  Staged ds(d, (size_t)h->P.ndofs, false, h->stream);
  point_diagonal_device(h, ds.dev);
  ds.copy_out(h->stream);
  if (ds.owned) PDB_CUDA(cudaStreamSynchronize(h->stream));
  PDB_CATCH
}

int pdb200_block_jacobi_apply(pdb200_handle h, const double* r, double* z) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!r || !z) throw Error("pdb200_block_jacobi_apply: null argument");
  if (!h->blockjac) h->blockjac = dg_blockjac_create(h->P, h->K);
  Staged rs(const_cast<double*>(r), (size_t)h->P.ndofs, true, h->stream), zs(z, (size_t)h->P.ndofs, false, h->stream);
  h->launches += launch_dg_blockjac(h->blockjac, h->P, h->K, rs.dev, zs.dev, h->stream);
  zs.copy_out(h->stream);
  if (zs.owned || rs.owned) PDB_CUDA(cudaStreamSynchronize(h->stream));
  PDB_CATCH
}

int pdb200_block_diagonal_apply(pdb200_handle h, const double* z, double* y) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!z || !y) throw Error("pdb200_block_diagonal_apply: null argument");
  if (!h->blockjac) h->blockjac = dg_blockjac_create(h->P, h->K);
  Staged zs(const_cast<double*>(z), (size_t)h->P.ndofs, true, h->stream), ys(y, (size_t)h->P.ndofs, false, h->stream);
  h->launches += launch_dg_blockdiag(h->blockjac, h->P, h->K, zs.dev, ys.dev, 1, h->stream);
  ys.copy_out(h->stream);
  if (zs.owned || ys.owned) PDB_CUDA(cudaStreamSynchronize(h->stream));
  PDB_CATCH
}

int pdb200_block_offdiagonal_apply(pdb200_handle h, const double* z, double* y) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!z || !y) throw Error("pdb200_block_offdiagonal_apply: null argument");
  if (!h->blockjac) h->blockjac = dg_blockjac_create(h->P, h->K);
  Staged zs(const_cast<double*>(z), (size_t)h->P.ndofs, true, h->stream), ys(y, (size_t)h->P.ndofs, false, h->stream);
  run_vector_device(h, zs.dev, ys.dev, Mode::OnTheFly);                                          // y = J z
  h->launches += launch_dg_blockdiag(h->blockjac, h->P, h->K, zs.dev, ys.dev, 2, h->stream);    // y -= D z
  ys.copy_out(h->stream);
  if (zs.owned || ys.owned) PDB_CUDA(cudaStreamSynchronize(h->stream));
  PDB_CATCH
}

int pdb200_block_sor_apply(pdb200_handle h, const double* d, double* v, double omega, int flags) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!d || !v) throw Error("pdb200_block_sor_apply: null argument");
  if (!(omega > 0.0 && omega < 2.0)) throw Error("pdb200_block_sor_apply: omega must be in (0, 2)");
  if (!h->blockjac) h->blockjac = dg_blockjac_create(h->P, h->K);
  const bool backward = flags & PDB200_SOR_BACKWARD, keep = flags & PDB200_SOR_KEEP_ITERATE;
  Staged ds(const_cast<double*>(d), (size_t)h->P.ndofs, true, h->stream), vs(v, (size_t)h->P.ndofs, keep, h->stream);
  h->launches += launch_dg_blocksor(h->blockjac, h->P, h->K, ds.dev, vs.dev, omega, backward, !keep, h->stream);
  vs.copy_out(h->stream);
  if (ds.owned || vs.owned) PDB_CUDA(cudaStreamSynchronize(h->stream));
  PDB_CATCH
}

int pdb200_set_relaxation(pdb200_handle h, double omega) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  if (!(omega > 0.0 && omega < 2.0)) throw Error("pdb200_set_relaxation: omega must be in (0, 2)");
  h->relaxation = omega;
  PDB_CATCH
}

int pdb200_solve(pdb200_handle h, int solver, int precond, const double* values, int layout, double* z, double* r,
                 double reduction, uint32_t maxiter, pdb200_solve_result* res) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!z || !r || !res) throw Error("pdb200_solve: null argument");
  Staged zs(z, (size_t)h->P.ndofs, true, h->stream), rs(r, (size_t)h->P.ndofs, true, h->stream);
  solve_device(h, solver, precond, values, layout, zs.dev, rs.dev, reduction, maxiter, res);
  zs.copy_out(h->stream);
  rs.copy_out(h->stream);
  PDB_CUDA(cudaStreamSynchronize(h->stream));
  check_errflag(h);
  PDB_CATCH
}

int pdb200_solve_stationary(pdb200_handle h, int solver, int precond, int matrix_free, double* x, double reduction,
                            double min_defect, uint32_t maxiter, pdb200_solve_result* res) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!x || !res) throw Error("pdb200_solve_stationary: null argument");
  const DevParams& P = h->P;
  const size_t n = (size_t)P.ndofs;
  if (!h->krylov) h->krylov = krylov_create();
  Staged xs(x, n, true, h->stream);
  double *values = nullptr, *r = nullptr, *z = nullptr;
  struct Free3 {
    double *&a, *&b, *&c;
    ~Free3() {
      if (a) cudaFree(a);
      if (b) cudaFree(b);
      if (c) cudaFree(c);
    }
  } guard{values, r, z};
  if (!matrix_free) {  // *_jacobian = 0; go.jacobian(x, *_jacobian)  (linearproblem.hh:221-226)
    if (!h->matrix) h->matrix = matrix_plan_create(P, h->fem, h->stream);
    uint64_t nrows = 0, nnz = 0;
    matrix_pattern_size(h->matrix, PDB200_LAYOUT_CSR, &nrows, &nnz);
    PDB_CUDA(cudaMalloc(&values, nnz * sizeof(double)));
    h->launches += matrix_assemble(h->matrix, PDB200_LAYOUT_CSR, values, true, /*fresh=*/true, h->errflag, h->stream);
  }
  PDB_CUDA(cudaMalloc(&r, n * sizeof(double)));
  PDB_CUDA(cudaMalloc(&z, n * sizeof(double)));
  PDB_CUDA(cudaMemsetAsync(r, 0, n * sizeof(double), h->stream));
  PDB_CUDA(cudaMemsetAsync(z, 0, n * sizeof(double), h->stream));
  run_vector_device(h, xs.dev, r, Mode::Residual);  // residual is additive (:244-246)
  const double defect = krylov_two_norm(h->krylov, P.ndofs, r, h->stream);
  const double red = defect > 0.0 ? std::max(reduction, min_defect / defect) : reduction;
  solve_device(h, solver, precond, values, PDB200_LAYOUT_CSR, z, r, red, maxiter, res);
  res->first_defect = defect;
  res->defect = defect * res->reduction;
  krylov_axpy(P.ndofs, -1.0, z, xs.dev, h->stream);  // *_x -= z
  h->launches += 1;
  xs.copy_out(h->stream);
  PDB_CUDA(cudaStreamSynchronize(h->stream));
  check_errflag(h);
  PDB_CATCH
}

int pdb200_gather_dofs(pdb200_handle h, const double* x, const int64_t* idx, uint64_t n, double* buf) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  launch_gather(x, (const long long*)idx, (long long)n, buf, h->stream);
  h->launches += 1;
  PDB_CATCH
}
int pdb200_scatter_dofs(pdb200_handle h, const double* buf, const int64_t* idx, uint64_t n, double* x) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  launch_scatter(buf, (const long long*)idx, (long long)n, x, h->stream);
  h->launches += 1;
  PDB_CATCH
}

int pdb200_halo_layer_size(pdb200_handle h, int dir, uint64_t* ndoubles) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  if (!h->P.dg) throw Error("halo exchange helpers are implemented for QkDG spaces");
  if (dir < 0 || dir >= h->P.dim) throw Error("invalid direction");
  *ndoubles = (uint64_t)(h->P.ncells / h->P.N[dir]) * h->P.n;
  PDB_CATCH
}
int pdb200_halo_pack(pdb200_handle h, const double* x, int dir, int side, double* buf) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->P.dg) throw Error("halo exchange helpers are implemented for QkDG spaces");
  launch_halo_copy(h->P, const_cast<double*>(x), buf, dir, side, /*pack=*/true, h->stream);
  h->launches += 1;
  PDB_CATCH
}
int pdb200_halo_unpack(pdb200_handle h, double* x, int dir, int side, const double* buf) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->P.dg) throw Error("halo exchange helpers are implemented for QkDG spaces");
  launch_halo_copy(h->P, x, const_cast<double*>(buf), dir, side, /*pack=*/false, h->stream);
  h->launches += 1;
  PDB_CATCH
}

int pdb200_onthefly_apply_part(pdb200_handle h, const double* x, double* y, int part) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (part < PDB200_PART_ALL || part > PDB200_PART_BOUNDARY) throw Error("invalid part");
  if (!is_device_pointer(x) || !is_device_pointer(y)) throw Error("apply_part expects device pointers");
  run_vector_device(h, x, y, Mode::OnTheFly, part);
  PDB_CATCH
}

int pdb200_halo_p2p_create(pdb200_handle h, pdb200_ipc_handle* mine) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!mine) throw Error("null argument");
  if (h->p2p) throw Error("p2p mailbox already created");
  h->p2p = p2p_create(h->P, mine);
  PDB_CATCH
}
int pdb200_halo_p2p_connect(pdb200_handle h, int dir, int side, const pdb200_ipc_handle* neighbour) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->p2p) throw Error("call pdb200_halo_p2p_create first");
  if (!neighbour) throw Error("null argument");
  p2p_connect(h->p2p, h->P, dir, side, neighbour);
  PDB_CATCH
}
int pdb200_halo_exchange_p2p(pdb200_handle h, double* x) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->p2p) throw Error("call pdb200_halo_p2p_create / _connect first");
  if (!is_device_pointer(x)) throw Error("halo_exchange_p2p expects a device pointer");
  h->launches += p2p_exchange(h->p2p, h->P, x, h->stream);
  PDB_CATCH
}
int pdb200_onthefly_apply_p2p(pdb200_handle h, double* x, double* y) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->p2p) throw Error("call pdb200_halo_p2p_create / _connect first");
  if (!is_device_pointer(x) || !is_device_pointer(y)) throw Error("apply_p2p expects device pointers");
  apply_p2p_device(h, x, y);
  PDB_CATCH
}

int pdb200_comm_create(pdb200_handle h, int rank, int size, pdb200_ipc_handle* mine) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!mine) throw Error("null argument");
  if (h->comm) throw Error("reduction mailbox already created");
  h->comm = comm_create(rank, size, mine);
  PDB_CATCH
}
int pdb200_comm_connect(pdb200_handle h, int peer_rank, const pdb200_ipc_handle* peer) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->comm) throw Error("call pdb200_comm_create first");
  comm_connect(h->comm, peer_rank, peer);
  PDB_CATCH
}
int pdb200_comm_sum(pdb200_handle h, double* values, int count) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!h->comm) throw Error("call pdb200_comm_create / _connect first");
  if (count < 1 || count > 2 || !values) throw Error("pdb200_comm_sum: one or two values");
  const int nb = krylov_partial_count();
  double* buf = nullptr;
  PDB_CUDA(cudaMalloc(&buf, 2 * (size_t)nb * sizeof(double)));
  struct F {
    double* p;
    ~F() { cudaFree(p); }
  } guard{buf};
  PDB_CUDA(cudaMemsetAsync(buf, 0, 2 * (size_t)nb * sizeof(double), h->stream));
  PDB_CUDA(cudaMemcpyAsync(buf, values, sizeof(double), cudaMemcpyDefault, h->stream));
  if (count == 2) PDB_CUDA(cudaMemcpyAsync(buf + nb, values + 1, sizeof(double), cudaMemcpyDefault, h->stream));
  comm_allreduce_partials(h->comm, buf, count == 2 ? buf + nb : nullptr, nb, h->stream);
  h->launches += 1;
  PDB_CUDA(cudaMemcpyAsync(values, buf, sizeof(double), cudaMemcpyDefault, h->stream));
  if (count == 2) PDB_CUDA(cudaMemcpyAsync(values + 1, buf + nb, sizeof(double), cudaMemcpyDefault, h->stream));
  PDB_CUDA(cudaStreamSynchronize(h->stream));
  comm_check(h->comm);
  PDB_CATCH
}
int pdb200_solve_ovlp(pdb200_handle h, int solver, int precond, const double* values, int layout, double* z, double* r,
                      double reduction, uint32_t maxiter, pdb200_solve_result* res) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  if (!z || !r || !res) throw Error("pdb200_solve_ovlp: null argument");
  if (!is_device_pointer(z) || !is_device_pointer(r)) throw Error("pdb200_solve_ovlp expects device vectors");
  solve_device(h, solver, precond, values, layout, z, r, reduction, maxiter, res, /*ovlp=*/true);
  PDB_CATCH
}

int pdb200_set_stream(pdb200_handle h, void* stream) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  h->stream = (cudaStream_t)stream;
  PDB_CATCH
}
int pdb200_synchronize(pdb200_handle h) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  ensure_device(h);
  PDB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->p2p) p2p_check(h->p2p);
  PDB_CATCH
}
int pdb200_launch_count(pdb200_handle h, uint64_t* n) {
  PDB_TRY
  PDB_CHECK_HANDLE(h);
  *n = h->launches;
  PDB_CATCH
}
const char* pdb200_last_kernel(pdb200_handle h) { return h ? h->last_kernel : ""; }

}  // extern "C"

// krylov.cu — device-resident Krylov solvers for the operators of this library: the consumers of
// jacobian_apply / the assembled Jacobian.
//
// Restates, on the GPU, what the reference's sequential ISTL solver back-ends run on one core
// (paths relative to /root/reference/dune/pdelab/):
//   ISTLBackend_SEQ_MatrixFree_BCGS_Richardson   backend/istl/seqistlsolverbackend.hh:157-203,1039-1050
//     = OnTheFlyOperator (:44-100) + Dune::Richardson(1.0) + Dune::BiCGSTABSolver
//   ISTLBackend_SEQ_CG_Jac / ISTLBackend_SEQ_BCGS_Jac   :208-255,401-416,538-553
//     = MatrixAdapter + Dune::SeqJac (one step, w = 1) + Dune::CGSolver / BiCGSTABSolver
//   norm = SequentialNorm (two_norm), result = LinearSolverResult (backend/solver.hh:28-51)
// The iteration itself (dune-istl >= 2.10 solvers.hh, un-vendored: restated from its published
// algorithm) keeps dune-istl's order of updates, its half-iteration count for BiCGSTAB and its
// stopping rule  def < def0 * reduction || def < 1e-30, so iteration counts are comparable with the
// reference's own matrix-free test (test/matrixfree/matrix_free_linear.cc:390-393).
//
// Mapping to the machine: all vectors and all scalars (rho, alpha, omega, beta) stay on the device.
// Every vector pass is one fused kernel (update + the inner products the next step needs); inner
// products are reduced in two deterministic stages: NB block partials, then every block of the
// consuming kernel re-adds the NB partials in the same fixed order, so there is no scalar kernel, no
// atomics, and results are bit-reproducible.  The host sees one 8-byte norm per convergence test.

#include <chrono>
#include <cmath>

#include "common.cuh"
#include "krylov.h"

namespace pdb {

namespace {

constexpr int NB = 148 * 4;  // blocks of every vector kernel = block partials per inner product
constexpr int NT = 256;

struct Scalars {  // device-resident recurrence scalars
  double rho, rho_new, alpha, omega, rholast;
};

__device__ __forceinline__ double block_reduce(double v) {
  __shared__ double ws[NT / 32];
  __shared__ double total;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();  // protects ws / total against the previous use
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double w = threadIdx.x < NT / 32 ? ws[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) total = w;
  }
  __syncthreads();
  return total;
}

// sum of the NB block partials, same order in every block
__device__ __forceinline__ double sum_partials(const double* __restrict__ P) {
  double v = 0.0;
  for (int i = threadIdx.x; i < NB; i += NT) v += P[i];
  return block_reduce(v);
}

#define GRID_LOOP(i, n) for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < (n); i += (long long)NB * NT)

// r = b - A x has been formed in r.  rt = r, p = v = 0;  PA = |r|^2 (= rt.r)
__global__ void __launch_bounds__(NT) bcgs_init_kernel(long long n, const double* __restrict__ r, double* __restrict__ rt,
                                                       double* __restrict__ p, double* __restrict__ v,
                                                       double* __restrict__ PA, Scalars* S) {
  double acc = 0.0;
  GRID_LOOP(i, n) {
    const double ri = r[i];
    rt[i] = ri;
    p[i] = 0.0;
    v[i] = 0.0;
    acc = fma(ri, ri, acc);
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) {
    PA[blockIdx.x] = acc;
    if (blockIdx.x == 0) {
      S->rho = 1.0;
      S->alpha = 1.0;
      S->omega = 1.0;
    }
  }
}

// rho_new = sum(PR);  beta = (rho_new / rho)(alpha / omega);  p = r + beta (p - omega v);  y = W p
__global__ void __launch_bounds__(NT) bcgs_direction_kernel(long long n, const double* __restrict__ r, double* __restrict__ p,
                                                            const double* __restrict__ v, const double* __restrict__ dinv,
                                                            double* __restrict__ y, const double* __restrict__ PR,
                                                            Scalars* S) {
  const double rho_new = sum_partials(PR);
  const double omega = S->omega;
  const double beta = (rho_new / S->rho) * (S->alpha / omega);
  GRID_LOOP(i, n) {
    const double pi = fma(beta, fma(-omega, v[i], p[i]), r[i]);
    p[i] = pi;
    if (dinv) y[i] = dinv[i] * pi;
  }
  __syncthreads();
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) S->rho_new = rho_new;  // read by later kernels only
}

// PA = a.b  (and PB = a.a if two)
__global__ void __launch_bounds__(NT) dot_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b,
                                                 double* __restrict__ PA, double* __restrict__ PB) {
  double ab = 0.0, aa = 0.0;
  GRID_LOOP(i, n) {
    const double ai = a[i];
    ab = fma(ai, b[i], ab);
    if (PB) aa = fma(ai, ai, aa);
  }
  ab = block_reduce(ab);
  if (PB) aa = block_reduce(aa);
  if (threadIdx.x == 0) {
    PA[blockIdx.x] = ab;
    if (PB) PB[blockIdx.x] = aa;
  }
}

// alpha = rho_new / sum(PH);  x += alpha y;  r -= alpha v;  y2 = W r;  PN = |r|^2
// (y and y2 may be the same buffer: no __restrict__ on them)
__global__ void __launch_bounds__(NT) bcgs_half1_kernel(long long n, double* __restrict__ x, const double* y,
                                                        double* __restrict__ r, const double* __restrict__ v,
                                                        const double* __restrict__ dinv, double* y2,
                                                        const double* __restrict__ PH, double* __restrict__ PN,
                                                        Scalars* S) {
  const double alpha = S->rho_new / sum_partials(PH);
  double acc = 0.0;
  GRID_LOOP(i, n) {
    x[i] = fma(alpha, y[i], x[i]);
    const double ri = fma(-alpha, v[i], r[i]);
    r[i] = ri;
    if (dinv) y2[i] = dinv[i] * ri;
    acc = fma(ri, ri, acc);
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) {
    PN[blockIdx.x] = acc;
    if (blockIdx.x == 0) S->alpha = alpha;
  }
}

// omega = sum(PTR) / sum(PTT);  x += omega y;  r -= omega t;  PN = |r|^2;  PR = rt.r;  rho = rho_new
// (y may be r itself: no __restrict__ on them)
__global__ void __launch_bounds__(NT) bcgs_half2_kernel(long long n, double* __restrict__ x, const double* y,
                                                        double* r, const double* __restrict__ t,
                                                        const double* __restrict__ rt, const double* __restrict__ PTR,
                                                        const double* __restrict__ PTT, double* __restrict__ PN,
                                                        double* __restrict__ PR, Scalars* S) {
  const double tr = sum_partials(PTR);
  const double tt = sum_partials(PTT);
  const double omega = tr / tt;
  double nn = 0.0, rr = 0.0;
  GRID_LOOP(i, n) {
    const double ri_old = r[i];
    const double yi = y == r ? ri_old : y[i];
    x[i] = fma(omega, yi, x[i]);
    const double ri = fma(-omega, t[i], ri_old);
    r[i] = ri;
    nn = fma(ri, ri, nn);
    rr = fma(rt[i], ri, rr);
  }
  nn = block_reduce(nn);
  rr = block_reduce(rr);
  if (threadIdx.x == 0) {
    PN[blockIdx.x] = nn;
    PR[blockIdx.x] = rr;
    if (blockIdx.x == 0) {
      S->omega = omega;
      S->rho = S->rho_new;
    }
  }
}

// CG start: p = W r;  PN = |r|^2;  PZ = p.r
__global__ void __launch_bounds__(NT) cg_init_kernel(long long n, const double* __restrict__ r, const double* __restrict__ dinv,
                                                     double* __restrict__ p, double* __restrict__ PN,
                                                     double* __restrict__ PZ) {
  double nn = 0.0, rz = 0.0;
  GRID_LOOP(i, n) {
    const double ri = r[i];
    const double zi = dinv ? dinv[i] * ri : ri;
    p[i] = zi;
    nn = fma(ri, ri, nn);
    rz = fma(zi, ri, rz);
  }
  nn = block_reduce(nn);
  rz = block_reduce(rz);
  if (threadIdx.x == 0) {
    PN[blockIdx.x] = nn;
    PZ[blockIdx.x] = rz;
  }
}

// lambda = rholast / sum(PPQ);  x += lambda p;  r -= lambda q;  z = W r;  PN = |r|^2;  PZ = z.r
// first != 0: rholast = sum(PZ0) (the start kernel's p.r)
__global__ void __launch_bounds__(NT) cg_update_kernel(long long n, double* __restrict__ x, const double* __restrict__ p,
                                                       double* __restrict__ r, const double* __restrict__ q,
                                                       const double* __restrict__ dinv, double* __restrict__ z,
                                                       const double* __restrict__ PPQ, const double* __restrict__ PZ0,
                                                       double* __restrict__ PN, double* __restrict__ PZ, Scalars* S,
                                                       int first) {
  const double rholast = first ? sum_partials(PZ0) : S->rholast;
  const double lambda = rholast / sum_partials(PPQ);
  double nn = 0.0, rz = 0.0;
  GRID_LOOP(i, n) {
    x[i] = fma(lambda, p[i], x[i]);
    const double ri = fma(-lambda, q[i], r[i]);
    r[i] = ri;
    const double zi = dinv ? dinv[i] * ri : ri;
    if (dinv) z[i] = zi;
    nn = fma(ri, ri, nn);
    rz = fma(zi, ri, rz);
  }
  nn = block_reduce(nn);
  rz = block_reduce(rz);
  __syncthreads();
  if (threadIdx.x == 0) {
    PN[blockIdx.x] = nn;
    PZ[blockIdx.x] = rz;
    if (first && blockIdx.x == gridDim.x - 1) S->rholast = rholast;
  }
}

// rho = sum(PZ);  beta = rho / rholast;  p = z + beta p;  rholast = rho
__global__ void __launch_bounds__(NT) cg_direction_kernel(long long n, double* __restrict__ p, const double* __restrict__ z,
                                                          const double* __restrict__ PZ, Scalars* S) {
  const double rho = sum_partials(PZ);
  const double beta = rho / S->rholast;
  GRID_LOOP(i, n) p[i] = fma(beta, p[i], z[i]);
  __syncthreads();
  // every block has read rholast before the last block can get here?  No: blocks run independently,
  // so the new value goes to a shadow slot and is committed by the next kernel of the stream
  if (blockIdx.x == 0 && threadIdx.x == 0) S->rho_new = rho;
}
__global__ void cg_commit_kernel(Scalars* S) { S->rholast = S->rho_new; }

// host_out[0] = sqrt(sum(P))  (host_out is pinned, mapped memory)
__global__ void __launch_bounds__(NT) norm_to_host_kernel(const double* __restrict__ P, double* host_out) {
  const double s = sum_partials(P);
  if (threadIdx.x == 0) host_out[0] = sqrt(s);
}

// r = b - r   (r holds A x)
__global__ void __launch_bounds__(NT) defect_kernel(long long n, const double* __restrict__ b, double* __restrict__ r) {
  GRID_LOOP(i, n) r[i] = b[i] - r[i];
}
__global__ void __launch_bounds__(NT) axpy_kernel(long long n, double a, const double* __restrict__ x, double* __restrict__ y) {
  GRID_LOOP(i, n) y[i] = fma(a, x[i], y[i]);
}
__global__ void __launch_bounds__(NT) norm2_kernel(long long n, const double* __restrict__ a, double* __restrict__ PA) {
  double acc = 0.0;
  GRID_LOOP(i, n) acc = fma(a[i], a[i], acc);
  acc = block_reduce(acc);
  if (threadIdx.x == 0) PA[blockIdx.x] = acc;
}

// dinv[row] = 1 / A(row,row) of a scalar CSR matrix (Dune::SeqJac with w = 1)
__global__ void diag_inv_kernel(long long nrows, const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
                                const double* __restrict__ values, double* __restrict__ dinv) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  double d = 1.0;
  for (uint64_t k = rowptr[row]; k < rowptr[row + 1]; k++)
    if (colidx[k] == (uint32_t)row) {
      d = values[k];
      break;
    }
  dinv[row] = 1.0 / d;
}

}  // namespace

struct KrylovWork {
  long long n = 0;
  double* vec[6] = {};
  double* partials = nullptr;  // 5 x NB
  Scalars* S = nullptr;
  double* host_norm = nullptr;  // pinned + mapped
  double* dev_norm = nullptr;   // device alias of host_norm
};

KrylovWork* krylov_create() { return new KrylovWork; }
void krylov_destroy(KrylovWork* w) {
  if (!w) return;
  for (double* v : w->vec)
    if (v) cudaFree(v);
  if (w->partials) cudaFree(w->partials);
  if (w->S) cudaFree(w->S);
  if (w->host_norm) cudaFreeHost(w->host_norm);
  delete w;
}

static void ensure(KrylovWork* w, long long n, int nvec) {
  if (w->n != n) {
    for (double*& v : w->vec)
      if (v) {
        cudaFree(v);
        v = nullptr;
      }
    w->n = n;
  }
  for (int i = 0; i < nvec; i++)
    if (!w->vec[i]) PDB_CUDA(cudaMalloc(&w->vec[i], (size_t)n * sizeof(double)));
  if (!w->partials) PDB_CUDA(cudaMalloc(&w->partials, 5 * NB * sizeof(double)));
  if (!w->S) PDB_CUDA(cudaMalloc(&w->S, sizeof(Scalars)));
  if (!w->host_norm) {
    PDB_CUDA(cudaHostAlloc(&w->host_norm, sizeof(double), cudaHostAllocMapped));
    PDB_CUDA(cudaHostGetDevicePointer(&w->dev_norm, w->host_norm, 0));
  }
}

static double read_norm(KrylovWork* w, const double* P, cudaStream_t s) {
  norm_to_host_kernel<<<1, NT, 0, s>>>(P, w->dev_norm);
  PDB_CUDA(cudaGetLastError());
  PDB_CUDA(cudaStreamSynchronize(s));
  return w->host_norm[0];
}

double krylov_two_norm(KrylovWork* w, long long n, const double* a, cudaStream_t s,
                       const std::function<void(double*, double*)>* allreduce) {
  ensure(w, w->n ? w->n : n, 0);
  norm2_kernel<<<NB, NT, 0, s>>>(n, a, w->partials);
  if (allreduce && *allreduce) (*allreduce)(w->partials, nullptr);
  return read_norm(w, w->partials, s);
}
int krylov_partial_count() { return NB; }

void krylov_axpy(long long n, double a, const double* x, double* y, cudaStream_t s) {
  axpy_kernel<<<NB, NT, 0, s>>>(n, a, x, y);
  PDB_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(NT) invert_kernel(long long n, double* __restrict__ d) {
  GRID_LOOP(i, n) d[i] = 1.0 / d[i];
}
void krylov_invert(long long n, double* d, cudaStream_t s) {
  invert_kernel<<<NB, NT, 0, s>>>(n, d);
  PDB_CUDA(cudaGetLastError());
}

void krylov_diag_inverse(long long nrows, const uint64_t* rowptr, const uint32_t* colidx, const double* values,
                         double* dinv, cudaStream_t s) {
  diag_inv_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, s>>>(nrows, rowptr, colidx, values, dinv);
  PDB_CUDA(cudaGetLastError());
}

// dune-istl Iteration::step: converged when def < def0 * reduction or def < 1e-30
static bool converged(double def, double def0, double reduction) { return def < def0 * reduction || def < 1e-30; }

int krylov_solve(KrylovWork* w, int solver, long long n, const KrylovOps& ops, double* x, double* b, double reduction,
                 unsigned maxit, cudaStream_t s, pdb200_solve_result* res) {
  const auto t0 = std::chrono::steady_clock::now();
  const double* dinv = ops.dinv;
  int launches = 0;
  res->converged = 0;
  res->iterations = 0;
  res->reduction = 0.0;
  res->conv_rate = 0.0;
  double def0 = 0.0, def = 0.0;
  double it = 0.0, it_done = 0.0;
  double *PA, *PB, *PC, *PD;
  // global sums of the overlapping scalar product: every producer of block partials is followed by the reduction
  auto AR = [&](double* P1, double* P2) {
    if (ops.allreduce) {
      ops.allreduce(P1, P2);
      launches++;
    }
  };
  if (solver == PDB200_SOLVER_BICGSTAB) {
    const bool gp = (bool)ops.prec;  // general preconditioner: y = W p / y = W r by separate launches
    ensure(w, n, dinv || gp ? 6 : 5);
    PA = w->partials, PB = PA + NB, PC = PB + NB, PD = PC + NB;
    double *r = w->vec[0], *rt = w->vec[1], *p = w->vec[2], *v = w->vec[3], *t = w->vec[4];
    double* y = dinv || gp ? w->vec[5] : nullptr;
    // r = b - A x  (BiCGSTABSolver::apply: op.applyscaleadd(-1, x, r))
    ops.apply(x, r);
    defect_kernel<<<NB, NT, 0, s>>>(n, b, r);
    bcgs_init_kernel<<<NB, NT, 0, s>>>(n, r, rt, p, v, PD, w->S);  // PD = rt.r = |r|^2
    launches += 2;
    AR(PD, nullptr);
    def0 = def = read_norm(w, PD, s);
    if (!converged(def0, def0, reduction) && def0 > 0.0) {
      for (it = 0.5; it < maxit; it += 0.5) {
        // next search direction and  v = A W p
        bcgs_direction_kernel<<<NB, NT, 0, s>>>(n, r, p, v, dinv, y, PD, w->S);
        if (gp) ops.prec(p, y);
        ops.apply(y ? y : p, v);
        dot_kernel<<<NB, NT, 0, s>>>(n, rt, v, PA, nullptr);
        AR(PA, nullptr);
        bcgs_half1_kernel<<<NB, NT, 0, s>>>(n, x, y ? y : p, r, v, dinv, y, PA, PC, w->S);
        AR(PC, nullptr);
        launches += 3;
        def = read_norm(w, PC, s);
        it_done = it;
        if (converged(def, def0, reduction)) break;
        it += 0.5;
        // second half: t = A W r
        if (gp) ops.prec(r, y);
        ops.apply(y ? y : r, t);
        dot_kernel<<<NB, NT, 0, s>>>(n, t, r, PA, PB);
        AR(PA, PB);
        bcgs_half2_kernel<<<NB, NT, 0, s>>>(n, x, y ? y : r, r, t, rt, PA, PB, PC, PD, w->S);
        AR(PC, PD);
        launches += 2;
        def = read_norm(w, PC, s);
        it_done = it;
        if (converged(def, def0, reduction)) break;
      }
    }
    it = it_done;  // Iteration::step records the (half) iteration of the last defect it saw
    // dune-istl hands the defect back in the right-hand side
    PDB_CUDA(cudaMemcpyAsync(b, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  } else if (solver == PDB200_SOLVER_CG) {
    const bool gp = (bool)ops.prec;
    ensure(w, n, dinv || gp ? 4 : 3);
    PA = w->partials, PB = PA + NB, PC = PB + NB, PD = PC + NB;
    double *r = w->vec[0], *p = w->vec[1], *q = w->vec[2];
    double* z = dinv || gp ? w->vec[3] : nullptr;
    ops.apply(x, r);
    defect_kernel<<<NB, NT, 0, s>>>(n, b, r);
    cg_init_kernel<<<NB, NT, 0, s>>>(n, r, dinv, p, PC, PD);  // PD = p.r (rholast)
    launches += 2;
    if (gp) {  // p = W r, rholast = p.r
      ops.prec(r, p);
      dot_kernel<<<NB, NT, 0, s>>>(n, p, r, PD, nullptr);
      launches++;
    }
    AR(PC, PD);
    def0 = def = read_norm(w, PC, s);
    if (!converged(def0, def0, reduction) && def0 > 0.0) {
      unsigned i = 1;
      for (; i <= maxit; i++) {
        ops.apply(p, q);
        dot_kernel<<<NB, NT, 0, s>>>(n, p, q, PA, nullptr);
        AR(PA, nullptr);
        cg_update_kernel<<<NB, NT, 0, s>>>(n, x, p, r, q, dinv, z, PA, PD, PC, PB, w->S, i == 1 ? 1 : 0);
        AR(PC, PB);
        launches += 2;
        def = read_norm(w, PC, s);
        it = i;
        if (converged(def, def0, reduction)) break;
        if (gp) {  // z = W r, rho = z.r (the fused kernel computed r.r into PB: replace it)
          ops.prec(r, z);
          dot_kernel<<<NB, NT, 0, s>>>(n, z, r, PB, nullptr);
          AR(PB, nullptr);
          launches++;
        }
        cg_direction_kernel<<<NB, NT, 0, s>>>(n, p, z ? z : r, PB, w->S);
        cg_commit_kernel<<<1, 1, 0, s>>>(w->S);
        launches += 2;
      }
      if (i > maxit) it = maxit;
    }
    PDB_CUDA(cudaMemcpyAsync(b, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  } else {
    throw Error("pdb200_solve: unknown solver");
  }
  PDB_CUDA(cudaGetLastError());
  PDB_CUDA(cudaStreamSynchronize(s));
  res->first_defect = def0;
  res->defect = def;
  res->iterations = (uint32_t)std::ceil(it);
  res->converged = def0 == 0.0 || converged(def, def0, reduction) ? 1 : 0;
  res->reduction = def0 > 0.0 ? def / def0 : 0.0;
  res->conv_rate = res->iterations ? std::pow(res->reduction, 1.0 / res->iterations) : 0.0;
  res->elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return launches;
}

}  // namespace pdb

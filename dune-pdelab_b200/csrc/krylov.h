// krylov.h — device-resident Krylov solvers (krylov.cu), bound to an operator by operator.cu
#pragma once

#include <cuda_runtime.h>

#include <functional>

#include "../../include/pdelab_b200.h"

namespace pdb {

struct KrylovOps {
  std::function<void(const double* in, double* out)> apply;  // out = A in (device pointers, overwrite)
  const double* dinv = nullptr;                               // point-Jacobi 1 / A_ii (fused into the vector kernels)
  std::function<void(const double* in, double* out)> prec;   // general preconditioner out = W in (e.g. block Jacobi);
                                                              // neither set = Richardson(1.0)
  // overlapping solvers (OverlappingScalarProduct): replaces the block partials of one or two inner products by
  // their sum over all ranks (P[0] = global sum, P[1..] = 0), in place, on the stream; unset = sequential
  std::function<void(double* P1, double* P2)> allreduce;
};

struct KrylovWork;
KrylovWork* krylov_create();
void krylov_destroy(KrylovWork*);
// solves A x = b (x: initial guess in, solution out; b: defect out); returns the number of launches of
// this module's kernels (the operator's own launches are counted by the caller's apply)
int krylov_solve(KrylovWork*, int solver, long long n, const KrylovOps& ops, double* x, double* b, double reduction,
                 unsigned maxit, cudaStream_t s, pdb200_solve_result* res);
double krylov_two_norm(KrylovWork*, long long n, const double* a, cudaStream_t s,
                       const std::function<void(double*, double*)>* allreduce = nullptr);
int krylov_partial_count();  // block partials per inner product
void krylov_axpy(long long n, double a, const double* x, double* y, cudaStream_t s);  // y += a x
void krylov_invert(long long n, double* d, cudaStream_t s);  // d <- 1 / d
void krylov_diag_inverse(long long nrows, const uint64_t* rowptr, const uint32_t* colidx, const double* values,
                         double* dinv, cudaStream_t s);

}  // namespace pdb

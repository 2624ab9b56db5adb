// operator.h — the operator handle behind include/pdelab_b200.h (shared by operator.cu and onestep.cu)
#pragma once

#include <cstdint>
#include <vector>

#include "common.cuh"
#include "host_tables.h"
#include "krylov.h"

using namespace pdb;

struct pdb200_operator {
  DevParams P;
  Kron1D K;
  int device = 0;
  double lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1};  // corners of the local box as given to pdb200_create
  int kernel_choice = PDB200_KERNEL_AUTO;
  cudaStream_t stream = nullptr;
  // device copies of the coefficient arrays
  std::vector<void*> owned;
  // staging for host-pointer calls
  double *dx = nullptr, *dy = nullptr;
  int* errflag = nullptr;
  // host-pointer calls of the fast kernel: transfers pipelined with the computation
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t pipe_ev[2 * 16 + 1] = {};
  FastPlan* fast = nullptr;
  KronPlan* kron = nullptr;
  FemPlan* fem = nullptr;
  MatrixPlan* matrix = nullptr;
  P2PHalo* p2p = nullptr;
  PeerComm* comm = nullptr;  // all-ranks reduction mailbox of the overlapping solvers
  KrylovWork* krylov = nullptr;
  BlockJacPlan* blockjac = nullptr;
  double* r0 = nullptr;  // R(0) of the affine DG residual, cached per coefficient set (fast path)
  bool r0_valid = false;
  double relaxation = 1.0;  // omega of the block SOR / SSOR preconditioners
  uint64_t launches = 0;
  uint64_t coeff_version = 0;  // bumped by pdb200_update_coefficients (one-step stage operators re-combine on change)
  const char* last_kernel = "";
  std::vector<double> xq, wq;

  ~pdb200_operator() {
    cudaSetDevice(device);
    for (void* p : owned) cudaFree(p);
    if (dx) cudaFree(dx);
    if (dy) cudaFree(dy);
    if (r0) cudaFree(r0);
    if (errflag) cudaFree(errflag);
    if (h2d_stream) cudaStreamDestroy(h2d_stream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
    for (auto& e : pipe_ev)
      if (e) cudaEventDestroy(e);
    dg_fast_plan_destroy(fast);
    dg_kron_plan_destroy(kron);
    fem_plan_destroy(fem);
    matrix_plan_destroy(matrix);
    p2p_destroy(p2p);
    comm_destroy(comm);
    krylov_destroy(krylov);
    dg_blockjac_destroy(blockjac);
  }
};

void pdb_set_last_error(const std::string& s);
void pdb_ensure_r0(pdb200_operator* op);             // operator.cu: compute op->r0 = R(0) if it is not cached
bool pdb_uses_cached_r0(const pdb200_operator* op);  // the residual runs as J x + cached R(0)

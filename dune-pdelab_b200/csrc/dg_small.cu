// dg_small.cu — Kronecker-factorised jacobian_apply / residual for the SMALL QkDG cells: dim = 2 with
// k = 1, 2 (n = 4, 9 DOFs per cell: the configurations of the reference's own DG tests,
// test/testconvectiondiffusiondg.cc, test/testfastdgassembler.cc, test/matrixfree/matrix_free_linear.cc)
// and dim = 3 with k = 1 (n = 8).
//
// Same operator identity as dg_fast.cu / dg_kron.cu, valid for cell-wise constant
// DIAGONAL diffusion tensors and cell-wise constant velocities (the convective terms fold into T, PL2 and PR2:
// dg_face.cuh Conv1); without convection:
//     y_e = |K| (M (x) ... (x) M) [ sum_d M^-1 L_d(z_{e-d}, z_e, z_{e+d}) / h_d^2 + c_e z_e ]
// i.e. exactly GridOperator::jacobian_apply for ConvectionDiffusionDG
// (localoperator/convectiondiffusiondg.hh:106-188, 271-471, 684-879); per line of n1 = k+1 nodes
//     t_i += sum_j T_ij o_j + PL1_i (d1.l) + PL2_i l_k + PR1_i (d0.r) + PR2_i r_0
// with the per-cell, per-direction matrix T and vectors PL*, PR* of dg_kron.cu.
//
// Mapping to the machine.  A cell is only 32..72 bytes, so a shared-memory tile with halo would be
// mostly halo.  One thread per cell, cells consecutive along x: a warp reads 32 neighbouring cells
// (1..2.3 KB contiguous) with vector loads; the 2*dim face neighbours are read the same way and hit
// L1/L2 (every cell is fetched from DRAM once).  Everything stays in registers; each thread writes
// its own n results (accumulate forms read-modify-write the same addresses).

#include "common.cuh"
#include "dg_face.cuh"

namespace pdb {

namespace {

using namespace dgface;

template <int DIM, int K>
__global__ void __launch_bounds__(128) dg_small_kernel(const DevParams P, const SmallConst<K> C, const double* __restrict__ z,
                                                       double* __restrict__ y, const double* __restrict__ r0,
                                                       int accumulate, int* __restrict__ errflag) {
  constexpr int N = SL<DIM, K>::N;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncells) return;
  const int Nx = P.N[0], Ny = P.N[1];
  int g[3];
  g[0] = (int)(cell % Nx);
  g[1] = (int)((cell / Nx) % Ny);
  g[2] = DIM == 3 ? (int)(cell / ((long long)Nx * Ny)) : 0;
  const long long stride[3] = {1, Nx, (long long)Nx * Ny};

  double o[N], t[N];
  load_cell<N>(z + cell * N, o);
  const double creact = P.c ? __ldg(P.c + cell) : 0.0;
  bool constrained = false;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double A0, cs[2], co[2], cg[2];
    bool onb[2];
    Conv1 V;
    constrained |= direction_coefs<K>(P, C, cell, g, d, stride, A0, cs, co, cg, onb, &V, errflag);
    double nb[2][N];
#pragma unroll
    for (int side = 0; side < 2; side++) {
      if (!onb[side]) {
        load_cell<N>(z + (cell + (side ? stride[d] : -stride[d])) * N, nb[side]);
      } else {
#pragma unroll
        for (int i = 0; i < N; i++) nb[side][i] = 0.0;
      }
    }
    if (d == 0)
      small_sweep<DIM, K, 0, true>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], creact, t, V);
    else if (d == 1)
      small_sweep<DIM, K, 1, false>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], 0.0, t, V);
    else
      small_sweep<DIM, K, 2, false>(C, o, nb[0], nb[1], A0, cs[0], co[0], cg[0], cs[1], co[1], cg[1], 0.0, t, V);
  }
  small_mass<DIM, K, 0>(C, C.vol, t);
  small_mass<DIM, K, 1>(C, 1.0, t);
  if (DIM == 3) small_mass<DIM, K, 2>(C, 1.0, t);
  double* __restrict__ out = y + cell * N;
  if (constrained) {  // constraints/p0.hh:31-41 + constrain_residual: the rows are SET to zero
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = 0.0;
    return;
  }
  if (r0) {
    double rr[N];
    load_cell<N>(r0 + cell * N, rr);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] += rr[i];
  }
  if (accumulate) {
#pragma unroll
    for (int i = 0; i < N; i++) t[i] += out[i];
  }
  store_cell<N>(out, t);
}

template <int DIM, int K>
void launch_variant(const DevParams& P, const Kron1D& K1, const double* z, double* y, const double* r0, bool overwrite,
                    cudaStream_t s, int* errflag) {
  SmallConst<K> C;
  fill_small_const<K>(C, P, K1);
  const unsigned blocks = (unsigned)((P.ncells + 127) / 128);
  dg_small_kernel<DIM, K><<<blocks, 128, 0, s>>>(P, C, z, y, r0, overwrite ? 0 : 1, errflag);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace

bool dg_small_supported(const DevParams& P) {
  // cell-wise constant diagonal A; a cell-wise constant velocity is a Kronecker term too
  if (!(P.dg && P.basis == PDB200_BASIS_LAGRANGE && P.pw == 0 && P.a_mode != PDB200_A_FULL && P.m >= P.k + 1)) return false;
  return (P.dim == 2 && (P.k == 1 || P.k == 2)) || (P.dim == 3 && P.k == 1);
}

int launch_dg_small(const DevParams& P, const Kron1D& K1, const double* z, double* y, const double* r0, bool overwrite,
                    cudaStream_t s, int* errflag) {
  if (r0 && overwrite) throw Error("the residual form accumulates (r += J x + R(0))");
  if (P.dim == 2 && P.k == 1) launch_variant<2, 1>(P, K1, z, y, r0, overwrite, s, errflag);
  else if (P.dim == 2 && P.k == 2) launch_variant<2, 2>(P, K1, z, y, r0, overwrite, s, errflag);
  else if (P.dim == 3 && P.k == 1) launch_variant<3, 1>(P, K1, z, y, r0, overwrite, s, errflag);
  else throw Error("dg_small: unsupported (dim, degree)");
  return 1;
}

}  // namespace pdb

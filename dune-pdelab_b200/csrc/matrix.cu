// matrix.cu — sparsity pattern, assembled Jacobian and SpMV.
//
// What it replaces (paths relative to /root/reference/dune/pdelab/):
//   GridOperator::fill_pattern -> DefaultLocalPatternAssemblerEngine -> BCRSPattern::add_link ->
//   allocate_bcrs_matrix        gridoperator/gridoperator.hh:168-173, default/patternengine.hh:146-204,
//                               localoperator/pattern.hh:13-47, backend/istl/bcrspattern.hh:96-119,
//                               backend/istl/bcrsmatrixbackend.hh:90-121 (setIndices: ascending columns)
//   GridOperator::jacobian -> jacobian_volume/_skeleton/_boundary -> scatter_jacobian ->
//   BCRSMatrix::operator()      gridoperator.hh:184-189, default/jacobianengine.hh,
//                               convectiondiffusionfem.hh:140-203,279-325, convectiondiffusiondg.hh:199-266,
//                               484-669,902-1044, gridoperator/common/assemblerutilities.hh:376-460
//   handle_dirichlet_constraints -> set_trivial_rows   assemblerutilities.hh:666-684, bcrsmatrix.hh:254-258
//
// The reference discovers the pattern by inserting links one by one (linear search per link) and
// scatters local matrices with a binary search per entry.  On a structured grid both are closed
// forms: the columns of a row are the lattice points of the box of cells around the row's lattice
// point (conforming Qk) or the DOFs of the cell and its face neighbours (QkDG), and their order
// inside the row follows from the container-index formula.  So
//   * row lengths are evaluated arithmetically and turned into rowptr by a device scan;
//   * colidx is written by one warp per row, each lane decoding its slot -> column;
//   * the Jacobian is assembled by ROW GATHER: the thread that owns a stored entry sums the
//     contributions of the cells that contain both DOFs, in ascending cell order (the order in
//     which the reference's scatter adds them), and writes the value once.  No atomics, no
//     colouring, no search, and colidx is not even read: 8 B per non-zero of HBM traffic.
// Local matrix entries of the conforming operator use the exactly integrated 1-D matrices (the
// reference's (k+1)-point Gauss rule integrates the same polynomials exactly, so values agree to
// rounding); QkDG blocks are integrated with the reference's quadrature loops.

#include <vector>

#include "common.cuh"
#include "host_tables.h"

namespace pdb {

namespace {

typedef unsigned long long u64;

// exactly integrated 1-D matrices of the Lagrange basis on [0,1] (k <= 2)
struct Mat1D {
  double M[9];  // int p_i p_j
  double K[9];  // int p_i' p_j'
  double C[9];  // int p_i' p_j   (row = test derivative)
};

// ---- exclusive scan of a row-length functor into rowptr[0..n] -------------------------------

constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u64 block_reduce(u64 v, u64* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 t = 0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < SCAN_THREADS / 32 ? sh[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

template <class F>
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(F f, u64 n, u64* __restrict__ bsum) {
  __shared__ u64 sh[32];
  const u64 base = (u64)blockIdx.x * SCAN_CHUNK + (u64)threadIdx.x * SCAN_ITEMS;
  u64 v = 0;
  for (int i = 0; i < SCAN_ITEMS; i++)
    if (base + i < n) v += f(base + i);
  v = block_reduce(v, sh);
  if (threadIdx.x == 0) bsum[blockIdx.x] = v;
}

// in-place exclusive scan of the block sums by ONE block; total goes to bsum[nb]
__global__ void __launch_bounds__(1024) scan_bsums_kernel(u64* __restrict__ bsum, u64 nb) {
  __shared__ u64 sh[1024];
  __shared__ u64 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (u64 start = 0; start < nb; start += 1024) {
    const u64 i = start + threadIdx.x;
    const u64 v = i < nb ? bsum[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      u64 t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) bsum[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

template <class F>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_write_kernel(F f, u64 n, const u64* __restrict__ bsum, u64* __restrict__ out) {
  __shared__ u64 sh[SCAN_THREADS];
  const u64 base = (u64)blockIdx.x * SCAN_CHUNK + (u64)threadIdx.x * SCAN_ITEMS;
  u64 loc[SCAN_ITEMS], v = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    loc[i] = base + i < n ? f(base + i) : 0;
    v += loc[i];
  }
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {
    u64 t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  u64 run = bsum[blockIdx.x] + sh[threadIdx.x] - v;
  for (int i = 0; i < SCAN_ITEMS; i++)
    if (base + i < n) {
      out[base + i] = run;
      run += loc[i];
    }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = bsum[gridDim.x];
}

template <class F>
int device_row_scan(F f, u64 n, u64* out /* n+1, device */, cudaStream_t s) {
  const u64 nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  u64* bsum = nullptr;
  PDB_CUDA(cudaMalloc(&bsum, (nb + 1) * sizeof(u64)));
  scan_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(f, n, bsum);
  scan_bsums_kernel<<<1, 1024, 0, s>>>(bsum, nb);
  scan_write_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(f, n, bsum, out);
  PDB_CUDA(cudaGetLastError());
  PDB_CUDA(cudaStreamSynchronize(s));
  PDB_CUDA(cudaFree(bsum));
  return 3;
}

// ---- conforming Qk: rows are lattice points -------------------------------------------------

// container index -> lattice point (inverse of qk_lattice_index)
__device__ __forceinline__ void qk_index_to_lattice(const QkLayout& L, long long idx, int p[3]) {
  p[0] = p[1] = p[2] = 0;
  if (L.k == 1) {
    for (int d = 0; d < L.dim; d++) {
      p[d] = (int)(idx % (L.N[d] + 1));
      idx /= L.N[d] + 1;
    }
    return;
  }
  int edim = 0;
  for (int e = 1; e <= L.dim; e++)
    if (idx >= L.block_off[e]) edim = e;
  idx -= L.block_off[edim];
  int s = 0;
  for (int g = 0; g < (1 << L.dim); g++)
    if (__popc(g) == edim && idx >= L.group_off[g]) s = g;  // group offsets ascend with g inside a block
  idx -= L.group_off[s];
  for (int d = 0; d < L.dim; d++) {
    const int ext = (s >> d) & 1;
    const int sz = ext ? L.N[d] : L.N[d] + 1;
    p[d] = 2 * (int)(idx % sz) + ext;
    idx /= sz;
  }
}

struct QkRow {
  int p[3], lo[3], hi[3];
  int len;
};

// the columns of row p: all lattice points of the cells that contain p (FullVolumePattern)
__device__ __forceinline__ QkRow qk_row(const QkLayout& L, long long row) {
  QkRow R;
  qk_index_to_lattice(L, row, R.p);
  R.len = 1;
  for (int d = 0; d < 3; d++) {
    R.lo[d] = R.hi[d] = 0;
    if (d >= L.dim) continue;
    const int k = L.k, pd = R.p[d];
    if (pd % k != 0) {
      R.lo[d] = (pd / k) * k;
      R.hi[d] = R.lo[d] + k;
    } else {
      R.lo[d] = max(pd - k, 0);
      R.hi[d] = min(pd + k, k * L.N[d]);
    }
    R.len *= R.hi[d] - R.lo[d] + 1;
  }
  return R;
}

// slot (position inside the row, ascending container index) -> column lattice point
__device__ __forceinline__ void qk_slot_to_lattice(const QkLayout& L, const QkRow& R, int slot, int q[3]) {
  q[0] = q[1] = q[2] = 0;
  if (L.k == 1) {
    for (int d = 0; d < L.dim; d++) {
      const int cnt = R.hi[d] - R.lo[d] + 1;
      q[d] = R.lo[d] + slot % cnt;
      slot /= cnt;
    }
    return;
  }
  // groups in container order: by entity dimension, then by bitset value
  for (int edim = 0; edim <= L.dim; edim++)
    for (int s = 0; s < (1 << L.dim); s++) {
      if (__popc(s) != edim) continue;
      int cnt[3] = {1, 1, 1}, first[3] = {0, 0, 0}, total = 1;
      for (int d = 0; d < L.dim; d++) {
        const int par = (s >> d) & 1;
        first[d] = R.lo[d] + (((R.lo[d] & 1) != par) ? 1 : 0);
        cnt[d] = first[d] <= R.hi[d] ? (R.hi[d] - first[d]) / 2 + 1 : 0;
        total *= cnt[d];
      }
      if (slot < total) {
        for (int d = 0; d < L.dim; d++) {
          q[d] = first[d] + 2 * (slot % cnt[d]);
          slot /= cnt[d];
        }
        return;
      }
      slot -= total;
    }
}

struct QkRowLen {
  QkLayout L;
  __device__ u64 operator()(u64 row) const { return (u64)qk_row(L, (long long)row).len; }
};

template <typename IDX>
__global__ void __launch_bounds__(256)
    qk_colidx_kernel(const QkLayout L, const u64* __restrict__ rowptr, u64 nrows, IDX* __restrict__ colidx) {
  const u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const QkRow R = qk_row(L, (long long)row);
  const u64 start = rowptr[row];
  for (int slot = lane; slot < R.len; slot += 32) {
    int q[3];
    qk_slot_to_lattice(L, R, slot, q);
    colidx[start + slot] = (IDX)qk_lattice_index(L, q);
  }
}

// local matrix entry (i = test, j = trial) of cell `cell`:  jacobian_volume,
// convectiondiffusionfem.hh:140-203, with exactly integrated 1-D factors
__device__ __forceinline__ double qk_volume_entry(const DevParams& P, const Mat1D& T, long long cell, const int li[3],
                                                  const int lj[3]) {
  const int n1 = P.n1;
  double m[3] = {1, 1, 1}, kk[3] = {0, 0, 0}, cij[3] = {0, 0, 0}, cji[3] = {0, 0, 0};
  for (int d = 0; d < P.dim; d++) {
    m[d] = T.M[li[d] * n1 + lj[d]];
    kk[d] = T.K[li[d] * n1 + lj[d]];
    cij[d] = T.C[li[d] * n1 + lj[d]];
    cji[d] = T.C[lj[d] * n1 + li[d]];
  }
  double A[3][3];
  load_A(P, cell, A);
  double v = 0.0;
  for (int a = 0; a < P.dim; a++) {
    double t = kk[a];
    for (int d = 0; d < P.dim; d++)
      if (d != a) t *= m[d];
    v += A[a][a] * P.ih[a] * P.ih[a] * t;
  }
  if (P.a_mode == PDB200_A_FULL)
    for (int a = 0; a < P.dim; a++)
      for (int b = 0; b < P.dim; b++) {
        if (a == b) continue;
        // int (d_b phi_j)(d_a phi_i): direction a carries p_i' p_j, direction b carries p_i p_j'
        double t = cij[a] * cji[b];
        for (int d = 0; d < P.dim; d++)
          if (d != a && d != b) t *= m[d];
        v += A[a][b] * P.ih[a] * P.ih[b] * t;
      }
  if (P.b)
    for (int a = 0; a < P.dim; a++) {
      double t = cij[a];
      for (int d = 0; d < P.dim; d++)
        if (d != a) t *= m[d];
      v -= __ldg(P.b + cell * P.dim + a) * P.ih[a] * t;
    }
  if (P.c) v += __ldg(P.c + cell) * m[0] * m[1] * m[2];
  return v * P.vol;
}

// jacobian_boundary (outflow faces only), convectiondiffusionfem.hh:279-325
__device__ __forceinline__ double qk_boundary_entry(const DevParams& P, const Mat1D& T, long long cell, const int c[3],
                                                    const int li[3], const int lj[3]) {
  if (!P.bctype || !P.b) return 0.0;
  double v = 0.0;
  for (int dir = 0; dir < P.dim; dir++)
    for (int side = 0; side < 2; side++) {
      const bool on = side ? c[dir] == P.N[dir] - 1 : c[dir] == 0;
      if (!on || P.side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;
      const int node = side ? P.k : 0;
      if (li[dir] != node || lj[dir] != node) continue;  // p_i(xi) = delta at the face
      if (P.bctype[bface_index(P, c, dir, side)] != PDB200_BC_OUTFLOW) continue;
      double t = __ldg(P.b + cell * P.dim + dir) * (side ? 1.0 : -1.0) * P.area[dir];
      for (int d = 0; d < P.dim; d++)
        if (d != dir) t *= T.M[li[d] * P.n1 + lj[d]];
      v += t;
    }
  return v;
}

__global__ void __launch_bounds__(256)
    qk_assemble_kernel(const DevParams P, const QkLayout L, const Mat1D T, const u64* __restrict__ rowptr, u64 nrows,
                       const unsigned char* __restrict__ constrained, double* __restrict__ values, int fresh) {
  const u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const QkRow R = qk_row(L, (long long)row);
  const u64 start = rowptr[row];
  const bool con = constrained && constrained[row];
  const int k = P.k;
  for (int slot = lane; slot < R.len; slot += 32) {
    int q[3];
    qk_slot_to_lattice(L, R, slot, q);
    if (con) {  // set_trivial_rows: clear the row, unit diagonal
      values[start + slot] = (q[0] == R.p[0] && q[1] == R.p[1] && q[2] == R.p[2]) ? 1.0 : 0.0;
      continue;
    }
    // cells containing both lattice points, per direction
    int c0[3] = {0, 0, 0}, nc[3] = {1, 1, 1};
    for (int d = 0; d < P.dim; d++) {
      const int lo = max(R.p[d], q[d]), hi = min(R.p[d], q[d]);  // cell c contains both iff k c <= hi, lo <= k c + k
      int first = (lo - k + k - 1) / k;                          // ceil((lo - k) / k), lo - k >= -k
      if (lo - k < 0) first = 0;
      int last = hi / k;
      if (last > P.N[d] - 1) last = P.N[d] - 1;
      c0[d] = first;
      nc[d] = last - first + 1;
    }
    double v = 0.0;
    for (int a2 = 0; a2 < nc[2]; a2++)
      for (int a1 = 0; a1 < nc[1]; a1++)
        for (int a0 = 0; a0 < nc[0]; a0++) {
          const int c[3] = {c0[0] + a0, c0[1] + a1, c0[2] + a2};
          const int li[3] = {R.p[0] - k * c[0], R.p[1] - k * c[1], R.p[2] - k * c[2]};
          const int lj[3] = {q[0] - k * c[0], q[1] - k * c[1], q[2] - k * c[2]};
          const long long cell = cell_index(P.N, c[0], c[1], c[2]);
          v += qk_volume_entry(P, T, cell, li, lj) + qk_boundary_entry(P, T, cell, c, li, lj);
        }
    values[start + slot] = fresh ? v : values[start + slot] + v;
  }
}

__global__ void qk_mv_kernel(const QkLayout L, const u64* __restrict__ rowptr, u64 nrows,
                             const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y) {
  const u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const QkRow R = qk_row(L, (long long)row);
  const u64 start = rowptr[row];
  double acc = 0.0;
  for (int slot = lane; slot < R.len; slot += 32) {
    int q[3];
    qk_slot_to_lattice(L, R, slot, q);
    acc = fma(values[start + slot], __ldg(x + qk_lattice_index(L, q)), acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

__global__ void set_flags_kernel(unsigned char* __restrict__ flags, const uint64_t* __restrict__ idx, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[idx[i]] = 1;
}

// ---- QkDG: block rows are cells ---------------------------------------------------------------

struct DgRow {
  int c[3];
  int nb;             // blocks in the row
  long long nbr[7];   // column cells ascending
  signed char face[7];  // -1: the cell itself, else 2*dir + side of the face shared with that neighbour
};

__device__ __forceinline__ DgRow dg_row(const DevParams& P, long long e) {
  DgRow R;
  long long t = e;
  R.c[0] = (int)(t % P.N[0]);
  t /= P.N[0];
  R.c[1] = (int)(t % P.N[1]);
  R.c[2] = (int)(t / P.N[1]);
  const long long stride[3] = {1, (long long)P.N[0], (long long)P.N[0] * P.N[1]};
  R.nb = 0;
  for (int d = P.dim - 1; d >= 0; d--)
    if (R.c[d] > 0) {
      R.nbr[R.nb] = e - stride[d];
      R.face[R.nb++] = (signed char)(2 * d);
    }
  R.nbr[R.nb] = e;
  R.face[R.nb++] = -1;
  for (int d = 0; d < P.dim; d++)
    if (R.c[d] < P.N[d] - 1) {
      R.nbr[R.nb] = e + stride[d];
      R.face[R.nb++] = (signed char)(2 * d + 1);
    }
  return R;
}

struct DgBlockLen {
  DevParams P;
  __device__ u64 operator()(u64 e) const {
    long long t = (long long)e;
    int c[3];
    c[0] = (int)(t % P.N[0]);
    t /= P.N[0];
    c[1] = (int)(t % P.N[1]);
    c[2] = (int)(t / P.N[1]);
    int nb = 1;
    for (int d = 0; d < P.dim; d++) nb += (c[d] > 0) + (c[d] < P.N[d] - 1);
    return (u64)nb;
  }
};

template <typename IDX>
__global__ void dg_colidx_kernel(const DevParams P, const u64* __restrict__ browptr, int block_layout,
                                 u64* __restrict__ rowptr_out, IDX* __restrict__ colidx) {
  // one block of threads per cell
  const long long e = blockIdx.x;
  const DgRow R = dg_row(P, e);
  const u64 b0 = browptr[e];
  const int n = P.n;
  if (block_layout) {
    if (threadIdx.x < R.nb) colidx[b0 + threadIdx.x] = (IDX)R.nbr[threadIdx.x];
    if (threadIdx.x == 0) {
      rowptr_out[e] = b0;
      if (e == P.ncells - 1) rowptr_out[P.ncells] = b0 + R.nb;
    }
    return;
  }
  const int rowlen = n * R.nb;
  for (int i = 0; i < n; i++) {
    const u64 start = (u64)n * n * b0 + (u64)i * rowlen;
    if (threadIdx.x == 0) rowptr_out[e * n + i] = start;
    for (int s = threadIdx.x; s < rowlen; s += blockDim.x) colidx[start + s] = (IDX)(R.nbr[s / n] * n + s % n);
  }
  if (threadIdx.x == 0 && e == P.ncells - 1) rowptr_out[P.ncells * n] = (u64)n * n * (b0 + R.nb);
}

// phi_i and its physical gradient at the point with 1-D table indices pt[]
__device__ __forceinline__ void dg_basis(const DevParams& P, const int pt[3], int i, double& phi, double grad[3]) {
  const int n1 = P.n1;
  int a[3] = {0, 0, 0};
  for (int d = 0; d < P.dim; d++) {
    a[d] = i % n1;
    i /= n1;
  }
  double pv[3] = {1, 1, 1}, dv[3] = {0, 0, 0};
  for (int d = 0; d < P.dim; d++) {
    pv[d] = P.P[pt[d] * n1 + a[d]];
    dv[d] = P.DP[pt[d] * n1 + a[d]] * P.ih[d];
  }
  phi = pv[0] * pv[1] * pv[2];
  grad[0] = dv[0] * pv[1] * pv[2];
  grad[1] = pv[0] * dv[1] * pv[2];
  grad[2] = P.dim == 3 ? pv[0] * pv[1] * dv[2] : 0.0;
}

__device__ __forceinline__ double dot3d(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Entry (i = test DOF of cell e, j = trial DOF of the column cell) of the block with face code `face`.
__device__ double dg_entry(const DevParams& P, long long e, const int c[3], int face, int i, int j, int* errflag) {
  const int m = P.m, dim = P.dim;
  double A_s[3][3], b_s[3] = {0, 0, 0};
  load_A(P, e, A_s);
  if (P.b)
    for (int d = 0; d < dim; d++) b_s[d] = P.b[e * dim + d];
  const long long stride[3] = {1, (long long)P.N[0], (long long)P.N[0] * P.N[1]};
  double v = 0.0;
  if (face < 0) {
    // jacobian_volume, convectiondiffusiondg.hh:199-266
    const double c_s = P.c ? P.c[e] : 0.0;
    for (int q = 0; q < P.nq; q++) {
      int pt[3] = {0, 0, 0}, qq = q;
      double w = 1.0;
      for (int d = 0; d < dim; d++) {
        pt[d] = qq % m;
        qq /= m;
        w *= P.wq[pt[d]];
      }
      double pi, gi[3], pj, gj[3], Agj[3];
      dg_basis(P, pt, i, pi, gi);
      dg_basis(P, pt, j, pj, gj);
      for (int a = 0; a < 3; a++) Agj[a] = A_s[a][0] * gj[0] + A_s[a][1] * gj[1] + A_s[a][2] * gj[2];
      v += (dot3d(Agj, gi) - pj * dot3d(b_s, gi) + c_s * pj * pi) * w * P.vol;
    }
  }
  for (int dir = 0; dir < dim; dir++)
    for (int side = 0; side < 2; side++) {
      if (face >= 0 && face != 2 * dir + side) continue;
      const bool onb = side ? c[dir] == P.N[dir] - 1 : c[dir] == 0;
      const double nsign = side ? 1.0 : -1.0;
      double An_s[3];
      for (int d = 0; d < 3; d++) An_s[d] = A_s[d][dir] * nsign;
      const double area = P.area[dir];
      if (!onb) {
        // jacobian_skeleton, :484-669, this cell's rows (ss / sn when it is the inside cell,
        // nn / ns when it is the outside cell)
        const long long other = e + (side ? stride[dir] : -stride[dir]);
        double A_o[3][3], An_o[3], b_F[3] = {0, 0, 0};
        load_A(P, other, A_o);
        for (int d = 0; d < 3; d++) An_o[d] = A_o[d][dir] * nsign;
        if (P.b) {
          const long long bc = side ? other : e;
          for (int d = 0; d < dim; d++) b_F[d] = P.b[bc * dim + d];
        }
        const double h_F = fmin(P.vol, P.vol) / area;
        double omega_s, omega_o, harm;
        if (P.weights_on) {
          const double ds = An_s[dir] * nsign, dn = An_o[dir] * nsign;
          omega_s = dn / (ds + dn + 1e-20);
          omega_o = ds / (ds + dn + 1e-20);
          harm = 2.0 * ds * dn / (ds + dn + 1e-20);
        } else {
          omega_s = omega_o = 0.5;
          harm = 1.0;
        }
        const double penalty = (P.alpha / h_F) * harm * P.k * (P.k + dim - 1);
        const double betan = b_F[dir] * nsign;
        const bool take_self = side == 0 ? (betan >= 0.0) : !((-betan) >= 0.0);
        for (int q = 0; q < P.nfq; q++) {
          int pt_s[3] = {0, 0, 0}, pt_o[3], qq = q;
          double w = 1.0;
          for (int d = 0; d < dim; d++)
            if (d != dir) {
              pt_s[d] = qq % m;
              qq /= m;
              w *= P.wq[pt_s[d]];
            }
          for (int d = 0; d < 3; d++) pt_o[d] = pt_s[d];
          pt_s[dir] = side ? m + 1 : m;
          pt_o[dir] = side ? m : m + 1;
          const double factor = w * area;
          double pi, gi[3], pj, gj[3];
          dg_basis(P, pt_s, i, pi, gi);
          if (face < 0) {
            dg_basis(P, pt_s, j, pj, gj);
            v += ((take_self ? pj * betan : 0.0) * pi - omega_s * dot3d(An_s, gj) * pi +
                  P.theta * omega_s * pj * dot3d(An_s, gi) + penalty * pj * pi) * factor;
          } else {
            dg_basis(P, pt_o, j, pj, gj);
            v += ((take_self ? 0.0 : pj * betan) * pi - omega_o * dot3d(An_o, gj) * pi -
                  P.theta * omega_s * pj * dot3d(An_s, gi) - penalty * pj * pi) * factor;
          }
        }
      } else if (face < 0 && P.side_kind[dir][side] != PDB200_SIDE_PROCESSOR) {
        // jacobian_boundary, :902-1044
        const long long bf = bface_index(P, c, dir, side);
        const int bctype = P.bctype ? (int)P.bctype[bf] : (int)PDB200_BC_DIRICHLET;
        if (bctype == PDB200_BC_NONE || bctype == PDB200_BC_NEUMANN) continue;
        const double h_F = P.vol / area;
        const double harm = P.weights_on ? An_s[dir] * nsign : 1.0;
        const double penalty = (P.alpha / h_F) * harm * P.k * (P.k + dim - 1);
        const double betan = b_s[dir] * nsign;
        if (bctype == PDB200_BC_OUTFLOW && betan < -1e-30) {
          *errflag = 1;
          continue;
        }
        for (int q = 0; q < P.nfq; q++) {
          int pt[3] = {0, 0, 0}, qq = q;
          double w = 1.0;
          for (int d = 0; d < dim; d++)
            if (d != dir) {
              pt[d] = qq % m;
              qq /= m;
              w *= P.wq[pt[d]];
            }
          pt[dir] = side ? m + 1 : m;
          const double factor = w * area;
          double pi, gi[3], pj, gj[3];
          dg_basis(P, pt, i, pi, gi);
          dg_basis(P, pt, j, pj, gj);
          if (bctype == PDB200_BC_OUTFLOW)
            v += pj * betan * factor * pi;
          else
            v += ((betan >= 0.0 ? pj * betan : 0.0) * pi - dot3d(An_s, gj) * pi + P.theta * pj * dot3d(An_s, gi) +
                  penalty * pj * pi) * factor;
        }
      }
    }
  return v;
}

__global__ void __launch_bounds__(128)
    dg_assemble_kernel(const DevParams P, const u64* __restrict__ browptr, int block_layout, double* __restrict__ values,
                       int fresh, int* __restrict__ errflag) {
  const long long e = blockIdx.x;
  const DgRow R = dg_row(P, e);
  const u64 b0 = browptr[e];
  const int n = P.n;
  bool constrained = false;  // P0ParallelConstraints: cells with a processor face (constraints/p0.hh:31-41)
  for (int d = 0; d < P.dim; d++) {
    if (R.c[d] == 0 && P.side_kind[d][0] == PDB200_SIDE_PROCESSOR) constrained = true;
    if (R.c[d] == P.N[d] - 1 && P.side_kind[d][1] == PDB200_SIDE_PROCESSOR) constrained = true;
  }
  const int total = R.nb * n * n;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int slot = t / (n * n), i = (t / n) % n, j = t % n;
    const u64 pos = block_layout ? (b0 + slot) * (u64)(n * n) + (u64)i * n + j
                                 : (u64)n * n * b0 + (u64)i * (n * R.nb) + (u64)slot * n + j;
    if (constrained) {
      values[pos] = (R.face[slot] < 0 && i == j) ? 1.0 : 0.0;
      continue;
    }
    const double v = dg_entry(P, e, R.c, R.face[slot], i, j, errflag);
    values[pos] = fresh ? v : values[pos] + v;
  }
}

__global__ void dg_mv_kernel(const DevParams P, const u64* __restrict__ browptr, int block_layout,
                             const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y) {
  const long long e = blockIdx.x;
  const DgRow R = dg_row(P, e);
  const u64 b0 = browptr[e];
  const int n = P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = warp; i < n; i += nwarps) {
    double acc = 0.0;
    for (int s = lane; s < n * R.nb; s += 32) {
      const int slot = s / n, j = s % n;
      const u64 pos = block_layout ? (b0 + slot) * (u64)(n * n) + (u64)i * n + j
                                   : (u64)n * n * b0 + (u64)i * (n * R.nb) + (u64)s;
      acc = fma(values[pos], __ldg(x + R.nbr[slot] * n + j), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) y[e * n + i] = acc;
  }
}

}  // namespace

struct MatrixPlan {
  DevParams P;
  QkLayout L;
  Mat1D T;
  u64 nrows = 0, nnz = 0, nbrows = 0, nblocks = 0;
  u64* rowptr = nullptr;           // Qk: scalar CSR row pointers; DG: block row pointers (cells)
  unsigned char* flags = nullptr;  // Qk: constrained rows
};

MatrixPlan* matrix_plan_create(const DevParams& P, FemPlan* fem, cudaStream_t s) {
  MatrixPlan* plan = new MatrixPlan;
  plan->P = P;
  try {
    if (P.dg) {
      plan->nbrows = (u64)P.ncells;
      PDB_CUDA(cudaMalloc(&plan->rowptr, (plan->nbrows + 1) * sizeof(u64)));
      DgBlockLen f{P};
      device_row_scan(f, plan->nbrows, plan->rowptr, s);
      PDB_CUDA(cudaMemcpy(&plan->nblocks, plan->rowptr + plan->nbrows, sizeof(u64), cudaMemcpyDeviceToHost));
      plan->nrows = (u64)P.ndofs;
      plan->nnz = plan->nblocks * (u64)P.n * (u64)P.n;
    } else {
      plan->L = fem_plan_layout(fem);
      plan->nrows = (u64)P.ndofs;
      PDB_CUDA(cudaMalloc(&plan->rowptr, (plan->nrows + 1) * sizeof(u64)));
      QkRowLen f{plan->L};
      device_row_scan(f, plan->nrows, plan->rowptr, s);
      PDB_CUDA(cudaMemcpy(&plan->nnz, plan->rowptr + plan->nrows, sizeof(u64), cudaMemcpyDeviceToHost));
      long long ncon = 0;
      const uint64_t* con = fem_plan_constrained(fem, &ncon);
      if (ncon) {
        PDB_CUDA(cudaMalloc(&plan->flags, plan->nrows));
        PDB_CUDA(cudaMemsetAsync(plan->flags, 0, plan->nrows, s));
        set_flags_kernel<<<(unsigned)((ncon + 255) / 256), 256, 0, s>>>(plan->flags, con, ncon);
        PDB_CUDA(cudaGetLastError());
      }
      // exact 1-D matrices in long double (Gauss rule with k+2 points)
      const int k = P.k, n1 = P.n1;
      std::vector<long double> ex, ew;
      host_gauss(k + 2, ex, ew);
      for (int i = 0; i < n1; i++)
        for (int j = 0; j < n1; j++) {
          long double M = 0, K = 0, C = 0;
          for (int q = 0; q < k + 2; q++) {
            const long double pi = host_lagrange_p_ld(k, i, ex[q]), pj = host_lagrange_p_ld(k, j, ex[q]);
            const long double di = host_lagrange_dp_ld(k, i, ex[q]), dj = host_lagrange_dp_ld(k, j, ex[q]);
            M += ew[q] * pi * pj;
            K += ew[q] * di * dj;
            C += ew[q] * di * pj;
          }
          plan->T.M[i * n1 + j] = (double)M;
          plan->T.K[i * n1 + j] = (double)K;
          plan->T.C[i * n1 + j] = (double)C;
        }
    }
    PDB_CUDA(cudaStreamSynchronize(s));
  } catch (...) {
    matrix_plan_destroy(plan);
    throw;
  }
  return plan;
}

void matrix_plan_destroy(MatrixPlan* p) {
  if (!p) return;
  if (p->rowptr) cudaFree(p->rowptr);
  if (p->flags) cudaFree(p->flags);
  delete p;
}

void matrix_pattern_size(MatrixPlan* p, int layout, uint64_t* nrows, uint64_t* nnz) {
  if (layout == PDB200_LAYOUT_BCSR) {
    *nrows = p->nbrows;
    *nnz = p->nblocks;
  } else {
    *nrows = p->nrows;
    *nnz = p->nnz;
  }
}

int matrix_pattern_write(MatrixPlan* p, int layout, void* rowptr, bool rowptr_dev, void* colidx, bool colidx_dev,
                         bool col32, cudaStream_t s) {
  const DevParams& P = p->P;
  const u64 nr = layout == PDB200_LAYOUT_BCSR ? p->nbrows : p->nrows;
  const u64 nz = layout == PDB200_LAYOUT_BCSR ? p->nblocks : p->nnz;
  if (col32 && (u64)P.ndofs > 0xffffffffull) throw Error("32-bit column indices need fewer than 2^32 DOFs");
  const size_t isz = col32 ? 4 : 8;
  u64* rp = (u64*)rowptr;
  void* ci = colidx;
  if (!rowptr_dev) PDB_CUDA(cudaMalloc(&rp, (nr + 1) * sizeof(u64)));
  if (!colidx_dev) PDB_CUDA(cudaMalloc(&ci, std::max<u64>(nz, 1) * isz));
  int launches = 1;
  if (P.dg) {
    if (col32)
      dg_colidx_kernel<uint32_t><<<(unsigned)P.ncells, 128, 0, s>>>(P, p->rowptr, layout == PDB200_LAYOUT_BCSR, rp, (uint32_t*)ci);
    else
      dg_colidx_kernel<u64><<<(unsigned)P.ncells, 128, 0, s>>>(P, p->rowptr, layout == PDB200_LAYOUT_BCSR, rp, (u64*)ci);
  } else {
    PDB_CUDA(cudaMemcpyAsync(rp, p->rowptr, (nr + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    const u64 blocks = (nr * 32 + 255) / 256;
    if (col32)
      qk_colidx_kernel<uint32_t><<<(unsigned)blocks, 256, 0, s>>>(p->L, p->rowptr, nr, (uint32_t*)ci);
    else
      qk_colidx_kernel<u64><<<(unsigned)blocks, 256, 0, s>>>(p->L, p->rowptr, nr, (u64*)ci);
  }
  PDB_CUDA(cudaGetLastError());
  if (!rowptr_dev) {
    PDB_CUDA(cudaMemcpyAsync(rowptr, rp, (nr + 1) * sizeof(u64), cudaMemcpyDeviceToHost, s));
  }
  if (!colidx_dev) {
    PDB_CUDA(cudaMemcpyAsync(colidx, ci, nz * isz, cudaMemcpyDeviceToHost, s));
  }
  if (!rowptr_dev || !colidx_dev) {
    PDB_CUDA(cudaStreamSynchronize(s));
    if (!rowptr_dev) cudaFree(rp);
    if (!colidx_dev) cudaFree(ci);
  }
  return launches;
}

int matrix_assemble(MatrixPlan* p, int layout, double* values, bool values_dev, bool fresh, int* errflag,
                    cudaStream_t s) {
  const DevParams& P = p->P;
  double* v = values;
  if (!values_dev) {
    PDB_CUDA(cudaMalloc(&v, std::max<u64>(p->nnz, 1) * sizeof(double)));
    if (!fresh) PDB_CUDA(cudaMemcpyAsync(v, values, p->nnz * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  if (P.dg) {
    dg_assemble_kernel<<<(unsigned)P.ncells, 128, 0, s>>>(P, p->rowptr, layout == PDB200_LAYOUT_BCSR, v, fresh ? 1 : 0, errflag);
  } else {
    const u64 blocks = (p->nrows * 32 + 255) / 256;
    qk_assemble_kernel<<<(unsigned)blocks, 256, 0, s>>>(P, p->L, p->T, p->rowptr, p->nrows, p->flags, v, fresh ? 1 : 0);
  }
  PDB_CUDA(cudaGetLastError());
  if (!values_dev) {
    PDB_CUDA(cudaMemcpyAsync(values, v, p->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    PDB_CUDA(cudaStreamSynchronize(s));
    cudaFree(v);
  }
  return 1;
}

int matrix_mv(MatrixPlan* p, int layout, const double* values, const double* x, double* y, cudaStream_t s) {
  const DevParams& P = p->P;
  if (P.dg) {
    dg_mv_kernel<<<(unsigned)P.ncells, 128, 0, s>>>(P, p->rowptr, layout == PDB200_LAYOUT_BCSR, values, x, y);
  } else {
    const u64 blocks = (p->nrows * 32 + 255) / 256;
    qk_mv_kernel<<<(unsigned)blocks, 256, 0, s>>>(p->L, p->rowptr, p->nrows, values, x, y);
  }
  PDB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pdb

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

#include <algorithm>
#include <cstdlib>
#include <cstring>
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

// n / d for 32-bit n with a precomputed multiplier (libdivide's branch-free u32 scheme)
struct FastDiv {
  uint32_t d, magic;
  int more;  // d == 1: more = -1
};
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
  if (f.more < 0) return n;
  const uint32_t q = __umulhi(f.magic, n);
  return (((n - q) >> 1) + q) >> f.more;
}

// Row decode of a conforming space: rows (container indices) are grouped by entity type; inside a
// group they are lexicographic in the anchor coordinates (ordering/leafgridviewordering.hh:166-184
// + YaspGrid index sets).  Built on the host from QkLayout; all indices fit 32 bits here.
struct QkDecode {
  int dim, k, ng;
  int N[3];
  uint32_t start[9];       // first row of group g (container order), start[ng] = number of rows
  uint8_t sbits[8];        // extension bitset of group g
  uint8_t group_of_s[8];   // inverse map
  uint32_t sz0[8], sz1[8];
  FastDiv d0[8], d1[8];
};

struct QkRow {
  int p[3], lo[3], hi[3];
  int len, shape;
};

// the columns of a row: all lattice points of the cells that contain the row's lattice point
// (FullVolumePattern, localoperator/pattern.hh:13-25)
__device__ __forceinline__ QkRow qk_row(const QkDecode& D, uint32_t row) {
  QkRow R;
  int g = 0;
#pragma unroll
  for (int i = 1; i < 8; i++)
    if (i < D.ng && row >= D.start[i]) g = i;
  uint32_t r = row - D.start[g];
  const uint32_t q0 = fast_div(r, D.d0[g]);
  const uint32_t a0 = r - q0 * D.sz0[g];
  const uint32_t q1 = fast_div(q0, D.d1[g]);
  const uint32_t a1 = q0 - q1 * D.sz1[g];
  const int a[3] = {(int)a0, (int)a1, (int)q1};
  const int s = D.sbits[g], k = D.k;
  R.len = 1;
  R.shape = 0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    R.p[d] = R.lo[d] = R.hi[d] = 0;
    if (d >= D.dim) continue;
    const int pd = k * a[d] + ((s >> d) & 1);
    R.p[d] = pd;
    if ((s >> d) & 1) {  // interior node of a cell (k = 2 only)
      R.lo[d] = pd - 1;
      R.hi[d] = pd + 1;
    } else {
      R.lo[d] = max(pd - k, 0);
      R.hi[d] = min(pd + k, k * D.N[d]);
    }
    const int len = R.hi[d] - R.lo[d] + 1;
    R.len *= len;
    R.shape |= (len == 2 * k + 1 ? 1 : 0) << d;
  }
  return R;
}

// lattice point -> container index (LFSIndexCache::containerIndex, lfsindexcache.hh:603-633)
__device__ __forceinline__ uint32_t qk_col_index(const QkDecode& D, const int q[3]) {
  if (D.k == 1) return (uint32_t)q[0] + D.sz0[0] * ((uint32_t)q[1] + D.sz1[0] * (uint32_t)q[2]);
  const int s = (q[0] & 1) | ((q[1] & 1) << 1) | ((q[2] & 1) << 2);
  const int g = D.group_of_s[s];
  return D.start[g] + (uint32_t)(q[0] >> 1) + D.sz0[g] * ((uint32_t)(q[1] >> 1) + D.sz1[g] * (uint32_t)(q[2] >> 1));
}

struct QkRowLen {
  QkDecode D;
  __device__ u64 operator()(u64 row) const { return (u64)qk_row(D, (uint32_t)row).len; }
};

// ---- slot decode tables --------------------------------------------------------------------
// The box of a row starts at an even lattice index in every direction (k = 2), so the order of
// its columns depends only on the box SHAPE (len_d in {k+1, 2k+1}): 8 shapes.  Two small tables
// per shape, built on the host by sorting the box offsets with the container-index key:
//   slot2off[shape][slot] = d0 | d1 << 3 | d2 << 6     (offset of the column inside the box)
//   off2slot[shape][d0 + 5 (d1 + 5 d2)] = slot
struct QkLut {
  uint16_t slot2off[8][125];
  uint8_t off2slot[8][125];
};

// unit local matrices: the local Jacobian of a cell with cell-wise constant coefficients is
//   sum_t w_t(cell) T_t,   T_t[i * n + j] (i = test, j = trial), built on the host from the exactly
// integrated 1-D matrices (jacobian_volume, convectiondiffusionfem.hh:140-203)
constexpr int QK_MAX_TABLES = 13;
struct QkTables {
  int nt;                    // number of tables
  int kind[QK_MAX_TABLES];   // 0: kappa (scalar A), 1: A[a][b], 2: -b[a], 3: c
  int ia[QK_MAX_TABLES], ib[QK_MAX_TABLES];
};

// weight w_t of the cell (uniform over the lanes that work on one row)
__device__ __forceinline__ double qk_cell_weight(const DevParams& P, const QkTables& Q, long long cell, int t) {
  switch (Q.kind[t]) {
    case 0: return P.a_mode == PDB200_A_IDENTITY ? 1.0 : __ldg(P.A + cell);
    case 1:
      return P.a_mode == PDB200_A_DIAGONAL ? __ldg(P.A + cell * P.dim + Q.ia[t])
                                           : __ldg(P.A + cell * P.dim * P.dim + Q.ia[t] * P.dim + Q.ib[t]);
    case 2: return -__ldg(P.b + cell * P.dim + Q.ia[t]);
    default: return __ldg(P.c + cell);
  }
}

// jacobian_boundary (outflow faces only), convectiondiffusionfem.hh:279-325
__device__ __forceinline__ double qk_boundary_entry(const DevParams& P, const Mat1D& T, long long cell, const int c[3],
                                                    const int li[3], const int lj[3]) {
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

constexpr int QK_THREADS = 256;

// G lanes work on one row (G = 32 for n = 27, 16 for n = 9, 8 for n <= 8)
template <int G, typename IDX>
__global__ void __launch_bounds__(QK_THREADS)
    qk_colidx_kernel(const QkDecode D, const QkLut* __restrict__ lut_g, const u64* __restrict__ rowptr, u64 nrows,
                     IDX* __restrict__ colidx, const uint32_t* __restrict__ rowlist) {
  __shared__ QkLut lut;
  for (int i = threadIdx.x; i < (int)(sizeof(QkLut) / 4); i += QK_THREADS) ((uint32_t*)&lut)[i] = ((const uint32_t*)lut_g)[i];
  __syncthreads();
  const int lane = threadIdx.x % G;
  u64 row = ((u64)blockIdx.x * QK_THREADS + threadIdx.x) / G;
  if (row >= nrows) return;
  if (rowlist) row = rowlist[row];  // nrows = length of the list
  const QkRow R = qk_row(D, (uint32_t)row);
  const u64 start = rowptr[row];
  for (int slot = lane; slot < R.len; slot += G) {
    const int off = lut.slot2off[R.shape][slot];
    const int q[3] = {R.lo[0] + (off & 7), R.lo[1] + ((off >> 3) & 7), R.lo[2] + (off >> 6)};
    colidx[start + slot] = (IDX)qk_col_index(D, q);
  }
}

// Row-gather assembly.  For every row: loop over the <= 2^dim cells that contain the row's lattice
// point in ascending cell order (uniform trip count, invalid candidates are predicated off); lane j
// adds the cell's local entry (i, j) into the row buffer in shared memory at the slot of column
// j; then the buffer is written out contiguously.  scatter_jacobian + UncachedMatrixView::add
// (assemblerutilities.hh:449-460, uncachedmatrixview.hh:259-262) without search or atomics.
// outflow boundary terms of one (row, cell) pass — rare, kept out of line
template <int DIM, int K>
__device__ __forceinline__ double qk_outflow_entry(const DevParams& P, const Mat1D& T1, int c0, int c1, int c2, int i, int j) {
  constexpr int N1 = K + 1;
  const int c[3] = {c0, c1, c2};
  const int li[3] = {i % N1, (i / N1) % N1, i / (N1 * N1)};
  const int lj[3] = {j % N1, (j / N1) % N1, j / (N1 * N1)};
  return qk_boundary_entry(P, T1, cell_index(P.N, c0, c1, c2), c, li, lj);
}

// Local Jacobian entry (i, j) of one cell with coefficients in the point-wise layout (P.pw != 0): the quadrature
// loops of jacobian_volume (convectiondiffusionfem.hh:140-203, A / b / c at every point) and, on outflow faces, of
// jacobian_boundary (:279-325, b at the face points), in the reference's order.  Parity path.
template <int DIM, int K>
__device__ double qk_pw_entry(const DevParams& P, int c0, int c1, int c2, int i, int j, bool outflow) {
  constexpr int N1 = K + 1;
  const int c[3] = {c0, c1, c2};
  const long long cell = cell_index(P.N, c0, c1, c2);
  const int li[3] = {i % N1, (i / N1) % N1, i / (N1 * N1)};
  const int lj[3] = {j % N1, (j / N1) % N1, j / (N1 * N1)};
  const int m = P.m;
  double A[3][3], b[3];
  load_A_cell(P, cell, A);
  const bool pwA = pw_A(P);
  double v = 0.0;
  for (int q = 0; q < P.nq; q++) {
    int pt[3] = {0, 0, 0};
    double w = 1.0;
    {
      int qq = q;
      for (int d = 0; d < DIM; d++) {
        pt[d] = qq % m;
        qq /= m;
        w *= P.wq[pt[d]];
      }
    }
    if (pwA) load_A_at(P, cell, q, A);
    load_b(P, cell, q, b);
    const double cc = load_c(P, cell, q);
    double phi_i = 1.0, phi_j = 1.0, gi[3] = {0, 0, 0}, gj[3] = {0, 0, 0};
    for (int d = 0; d < DIM; d++) {
      phi_i *= P.P[pt[d] * N1 + li[d]];
      phi_j *= P.P[pt[d] * N1 + lj[d]];
    }
    for (int d = 0; d < DIM; d++) {
      double a = P.ih[d] * P.DP[pt[d] * N1 + li[d]], bb = P.ih[d] * P.DP[pt[d] * N1 + lj[d]];
      for (int l = 0; l < DIM; l++)
        if (l != d) {
          a *= P.P[pt[l] * N1 + li[l]];
          bb *= P.P[pt[l] * N1 + lj[l]];
        }
      gi[d] = a;
      gj[d] = bb;
    }
    double Agj[3];
    for (int a = 0; a < 3; a++) Agj[a] = A[a][0] * gj[0] + A[a][1] * gj[1] + A[a][2] * gj[2];
    const double adv = b[0] * gi[0] + b[1] * gi[1] + b[2] * gi[2];
    v += (Agj[0] * gi[0] + Agj[1] * gi[1] + Agj[2] * gi[2] - phi_j * adv + cc * phi_j * phi_i) * (w * P.vol);
  }
  if (outflow)
    for (int dir = 0; dir < DIM; dir++)
      for (int side = 0; side < 2; side++) {
        const bool on = side ? c[dir] == P.N[dir] - 1 : c[dir] == 0;
        if (!on || P.side_kind[dir][side] == PDB200_SIDE_PROCESSOR) continue;
        const int node = side ? K : 0;
        if (li[dir] != node || lj[dir] != node) continue;  // p_i(xi) = delta at the face
        if (P.bctype[bface_index(P, c, dir, side)] != PDB200_BC_OUTFLOW) continue;  // type at the face centre
        for (int q = 0; q < P.nfq; q++) {
          double w = 1.0, pij = 1.0;
          int qq = q;
          for (int d = 0; d < DIM; d++)
            if (d != dir) {
              const int ptd = qq % m;
              qq /= m;
              w *= P.wq[ptd];
              pij *= P.P[ptd * N1 + li[d]] * P.P[ptd * N1 + lj[d]];
            }
          load_b(P, cell, face_pt(P, dir, side, q), b);
          v += b[dir] * (side ? 1.0 : -1.0) * pij * (w * P.area[dir]);
        }
      }
  return v;
}

// NT1: a single unit table (scalar or identity diffusion, no b, no c) — the cell weights then
// travel by warp shuffle; otherwise they are staged in shared memory.  PW: coefficients in the point-wise layout,
// entries by quadrature (qk_pw_entry) instead of unit tables.
template <int G, int DIM, int K, bool NT1, bool OUTFLOW, bool PW = false>
__global__ void __launch_bounds__(QK_THREADS)
    qk_assemble_kernel(const DevParams P, const QkDecode D, const Mat1D T1, const QkTables Q,
                       const double* __restrict__ tables_g, const QkLut* __restrict__ lut_g,
                       const u64* __restrict__ rowptr, u64 nrows, const unsigned char* __restrict__ constrained,
                       double* __restrict__ values, int fresh, const uint32_t* __restrict__ rowlist) {
  constexpr int N1 = K + 1, N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  constexpr int MAXLEN = DIM == 3 ? (2 * K + 1) * (2 * K + 1) * (2 * K + 1) : (2 * K + 1) * (2 * K + 1);
  constexpr int ROWS_PER_CTA = QK_THREADS / G;
  constexpr int NCAND = DIM == 3 ? 8 : 4;
  extern __shared__ double sm[];
  const int nt = PW ? 0 : (NT1 ? 1 : Q.nt);
  double* tab = sm;                                    // [nt][N*N]
  double* buf = tab + nt * N * N;                      // [ROWS_PER_CTA][MAXLEN]
  double* wbuf = buf + ROWS_PER_CTA * MAXLEN;          // [ROWS_PER_CTA][NCAND][nt]   (unused if NT1)
  const uint8_t* off2slot = (const uint8_t*)(wbuf + (NT1 ? 0 : ROWS_PER_CTA * NCAND * nt));  // [8][125]
  for (int i = threadIdx.x; i < nt * N * N; i += QK_THREADS) tab[i] = tables_g[i];
  for (int i = threadIdx.x; i < 250; i += QK_THREADS) ((uint32_t*)off2slot)[i] = ((const uint32_t*)lut_g->off2slot)[i];
  __syncthreads();
  const int lane = threadIdx.x % G, grp = threadIdx.x / G;
  // lanes that work on the same row: trip counts and branches below are uniform inside a group
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  double* rb = buf + grp * MAXLEN;
  double* wb = wbuf + grp * NCAND * nt;
  const u64 row0 = (u64)blockIdx.x * ROWS_PER_CTA + grp;
  const u64 rstride = (u64)gridDim.x * ROWS_PER_CTA;
  const int joff = (lane % N1) + 5 * (((lane / N1) % N1) + 5 * (lane / (N1 * N1)));
  const int N0c = P.N[0], N1c = P.N[1], N2c = P.N[2];
  for (u64 rix = row0; rix < nrows; rix += rstride) {
    const u64 row = rowlist ? (u64)rowlist[rix] : rix;  // with a list, nrows is its length
    const QkRow R = qk_row(D, (uint32_t)row);
    const u64 start = rowptr[row];
    const uint8_t* o2s = off2slot + R.shape * 125;
    if (constrained && constrained[row]) {  // set_trivial_rows: clear the row, unit diagonal
      const int diag = o2s[(R.p[0] - R.lo[0]) + 5 * ((R.p[1] - R.lo[1]) + 5 * (R.p[2] - R.lo[2]))];
      for (int s = lane; s < R.len; s += G) values[start + s] = s == diag ? 1.0 : 0.0;
      continue;
    }
    // the two candidate cells per direction, floor((p-1)/K) and floor(p/K), and whether they exist
    int cl[3] = {0, 0, 0};
    bool v0[3] = {true, true, true}, v1[3] = {false, false, false};
    const int Nc[3] = {N0c, N1c, N2c};
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const int pd = R.p[d], c = pd / K;
      if (pd % K != 0) {
        cl[d] = c;
      } else {
        cl[d] = c - 1;
        v0[d] = c - 1 >= 0;
        v1[d] = c < Nc[d];
      }
    }
    // weights of the candidate cells: one round of loads
    double wreg = 0.0;
    if (NT1) {
      if (lane < NCAND) {
        const int a0 = lane & 1, a1 = (lane >> 1) & 1, a2 = lane >> 2;
        const bool ok = (a0 ? v1[0] : v0[0]) && (a1 ? v1[1] : v0[1]) && (DIM == 3 ? (a2 ? v1[2] : v0[2]) : a2 == 0);
        if (ok) wreg = P.a_mode == PDB200_A_IDENTITY ? 1.0 : __ldg(P.A + cell_index(P.N, cl[0] + a0, cl[1] + a1, cl[2] + a2));
      }
    } else {
#pragma unroll
      for (int ci = 0; ci < NCAND; ci++) {
        const int a0 = ci & 1, a1 = (ci >> 1) & 1, a2 = ci >> 2;
        const bool ok = (a0 ? v1[0] : v0[0]) && (a1 ? v1[1] : v0[1]) && (DIM == 3 ? (a2 ? v1[2] : v0[2]) : a2 == 0);
        if (ok) {
          const long long cell = cell_index(P.N, cl[0] + a0, cl[1] + a1, cl[2] + a2);
          for (int t = lane; t < nt; t += G) wb[ci * nt + t] = qk_cell_weight(P, Q, cell, t);
        }
      }
    }
    for (int s = lane; s < R.len; s += G) rb[s] = 0.0;
    __syncwarp(gmask);
    // passes over the candidate cells in ascending cell order (the reference's scatter order)
#pragma unroll
    for (int ci = 0; ci < NCAND; ci++) {
      const int a0 = ci & 1, a1 = (ci >> 1) & 1, a2 = ci >> 2;
      const bool ok = (a0 ? v1[0] : v0[0]) && (a1 ? v1[1] : v0[1]) && (DIM == 3 ? (a2 ? v1[2] : v0[2]) : a2 == 0);
      const double w1 = NT1 ? __shfl_sync(gmask, wreg, ci, G) : 0.0;
      if (!ok) continue;
      const int c0 = cl[0] + a0, c1 = cl[1] + a1, c2 = DIM == 3 ? cl[2] + a2 : 0;
      const int i = (R.p[0] - K * c0) + N1 * ((R.p[1] - K * c1) + N1 * (DIM == 3 ? R.p[2] - K * c2 : 0));
      const int pbase = (K * c0 - R.lo[0]) + 5 * ((K * c1 - R.lo[1]) + 5 * (DIM == 3 ? K * c2 - R.lo[2] : 0));
      if (lane < N) {
        double v;
        if (PW) {
          v = qk_pw_entry<DIM, K>(P, c0, c1, c2, i, lane, OUTFLOW);
        } else if (NT1) {
          v = w1 * tab[i * N + lane];
        } else {
          v = 0.0;
          for (int t = 0; t < nt; t++) v = fma(wb[ci * nt + t], tab[(t * N + i) * N + lane], v);
        }
        if (OUTFLOW && !PW) v += qk_outflow_entry<DIM, K>(P, T1, c0, c1, c2, i, lane);
        rb[o2s[pbase + joff]] += v;
      }
      __syncwarp(gmask);
    }
    for (int s = lane; s < R.len; s += G) values[start + s] = fresh ? rb[s] : values[start + s] + rb[s];
    __syncwarp(gmask);
  }
}

// ---- interior rows, one x-line of rows per CTA ------------------------------------------------------
// All rows of an entity group whose lattice point is not on the boundary have the same box shape, so
// the column offset of slot s is the same for every row: a thread keeps ONE slot and walks along the
// x-line of rows.  Everything that depends on (group, slot) — the cells that contain both the row's
// and the column's lattice point and the local indices (i, j) in each — is decoded once per thread;
// per row what is left is  v = sum_k T[i_k][j_k] * kappa[cell_k]  in ascending cell order (the order of
// the reference's scatter_jacobian, assemblerutilities.hh:449-460) and one coalesced store, or, for
// the pattern, colidx = base + row (the column's group is fixed by the offset parity).
// Rows on the boundary (clipped boxes, constrained rows) stay with the generic kernels above, which
// are then driven by an explicit row list.
// row stride of the staged coefficient rows: = 4 (mod 16) doubles, so that the four rows and the two
// x-neighbours of a candidate set fall into distinct shared-memory banks
__host__ __device__ inline int qk_kap_stride(int N0) { return (N0 + 11) / 16 * 16 + 4; }

template <int DIM, int K, typename IDX, bool VALUES>
__global__ void __launch_bounds__(256, 4) qk_interior_kernel(const DevParams P, const QkDecode D, int g, int L, const QkLut* __restrict__ lut_g,
                                   const double* __restrict__ tab_g, const u64* __restrict__ rowptr,
                                   double* __restrict__ values, IDX* __restrict__ colidx, int fresh) {
  constexpr int N1 = K + 1, N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  extern __shared__ double kap[];  // [2][2][N0] coefficients of the candidate cells of this line
  const int s = D.sbits[g];
  const bool vt[3] = {!(s & 1), !((s >> 1) & 1), !((s >> 2) & 1)};  // vertex-type direction: two candidate cells
  const int N0 = P.N[0], KS = qk_kap_stride(N0);
  const int sz0 = (int)D.sz0[g], sz1 = (int)D.sz1[g];
  const int a1 = (int)blockIdx.x + (vt[1] ? 1 : 0), a2 = DIM == 3 ? (int)blockIdx.y + (vt[2] ? 1 : 0) : 0;
  const int lo0 = vt[0] ? 1 : 0, hi0 = N0 - 1;
  // rows of the line are packed densely over the CTA: R = blockDim / L rows in flight, slot fixed per thread
  const int R = (int)blockDim.x / L;
  const int rsub = (int)threadIdx.x / L, slot = (int)threadIdx.x - rsub * L;
  const bool active = rsub < R;
  int shape = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++) shape |= (vt[d] ? 1 : 0) << d;
  const int off = active ? lut_g->slot2off[shape][slot] : 0;
  const int o[3] = {off & 7, (off >> 3) & 7, off >> 6};
  int e[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < DIM; d++) e[d] = o[d] - (vt[d] ? K : 1);
  const u64 base = rowptr[D.start[g] + (u64)lo0 + (u64)sz0 * ((u64)a1 + (u64)sz1 * (u64)a2)];
  if (VALUES) {
    const int cb1 = a1 - (vt[1] ? 1 : 0), cb2 = DIM == 3 ? a2 - (vt[2] ? 1 : 0) : 0;
    const int cn1 = vt[1] ? 2 : 1, cn2 = DIM == 3 ? (vt[2] ? 2 : 1) : 1;
    for (int i = threadIdx.x; i < cn2 * cn1 * N0; i += blockDim.x) {
      const int x = i % N0, j1 = (i / N0) % cn1, j2 = i / (N0 * cn1);
      kap[(j2 * 2 + j1) * KS + x] =
          P.a_mode == PDB200_A_IDENTITY ? 1.0 : __ldg(P.A + cell_index(P.N, x, cb1 + j1, cb2 + j2));
    }
    __syncthreads();
    // per direction: candidate cell (0: lower / only, 1: upper), local row index, local column index
    int nd[3] = {1, 1, 1}, dl[3][2] = {}, li[3][2] = {}, lj[3][2] = {};
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      int n = 0;
      if (vt[d]) {
        if (e[d] <= 0) { dl[d][n] = 0; li[d][n] = K; lj[d][n] = e[d] + K; n++; }
        if (e[d] >= 0) { dl[d][n] = 1; li[d][n] = 0; lj[d][n] = e[d]; n++; }
      } else {
        dl[d][n] = 0; li[d][n] = 1; lj[d][n] = e[d] + 1; n++;
      }
      nd[d] = n;
    }
    // the <= 2^DIM (cell, i, j) combinations in ascending cell order; a combination is evaluated only
    // if some lane of the warp needs it (most entries couple through a single cell)
    double coef[8];
    int idx[8];
    bool any[8];
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++)
#pragma unroll
      for (int c1 = 0; c1 < 2; c1++)
#pragma unroll
        for (int c0 = 0; c0 < 2; c0++) {
          const int k = c0 + 2 * c1 + 4 * c2;
          const bool on = active && c0 < nd[0] && c1 < nd[1] && c2 < nd[2];
          const int i = li[0][c0] + N1 * (li[1][c1] + N1 * (DIM == 3 ? li[2][c2] : 0));
          const int j = lj[0][c0] + N1 * (lj[1][c1] + N1 * (DIM == 3 ? lj[2][c2] : 0));
          coef[k] = on ? tab_g[i * N + j] : 0.0;
          idx[k] = on ? (dl[2][c2] * 2 + dl[1][c1]) * KS + dl[0][c0] : 0;
          any[k] = __any_sync(0xffffffffu, on);
        }
    if (!active) return;
    const int xoff = vt[0] ? -1 : 0;
    double* dst = values + base + (size_t)rsub * L + slot;
    for (int a0 = lo0 + rsub; a0 <= hi0; a0 += R, dst += (size_t)R * L) {
      const double* kp = kap + a0 + xoff;
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (any[k]) v = fma(coef[k], kp[idx[k]], v);  // lanes that do not need it carry coef = 0, idx = 0
      *dst = fresh ? v : *dst + v;
    }
  } else {
    if (!active) return;
    long long col0;
    if (K == 1) {
      col0 = (lo0 + e[0]) + (long long)(N0 + 1) * ((a1 + e[1]) + (long long)(P.N[1] + 1) * (DIM == 3 ? a2 + e[2] : 0));
    } else {
      int par = 0, sh[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const int q = ((s >> d) & 1) + e[d];  // lattice offset of the column relative to 2 a_d
        const int pb = q & 1;
        par |= pb << d;
        sh[d] = (q - pb) / 2;
      }
      const int g2 = D.group_of_s[par];
      col0 = (long long)D.start[g2] + (lo0 + sh[0]) +
             (long long)D.sz0[g2] * ((a1 + sh[1]) + (long long)D.sz1[g2] * (DIM == 3 ? a2 + sh[2] : 0));
    }
    IDX* dst = colidx + base + (size_t)rsub * L + slot;
    for (int a0 = lo0 + rsub; a0 <= hi0; a0 += R, dst += (size_t)R * L) *dst = (IDX)(col0 + (a0 - lo0));
  }
}

// ---- values of the interior rows, several x-lines per CTA --------------------------------------------
// The decode of a thread's slot (which cells couple the row's and the column's lattice point, and with which local
// indices) is the same for every line of an entity group, and with one line per CTA it dominated short lines: two
// dependent table look-ups and the coefficient rows cost about as long as writing the line.  Here a CTA decodes once
// and walks QKV_LINES consecutive lines; the coefficient rows of the next line arrive by cp.async while the current
// one is written.  Along a line a thread takes CONSECUTIVE rows, so the coefficient of the upper x-cell of one row is
// the lower x-cell's of the next: per entry at most four shared-memory loads (one per (y, z) candidate pair) instead
// of eight, with the addresses held in four running pointers.  Sum order = ascending cell index, as in
// qk_interior_kernel and the reference's scatter (assemblerutilities.hh:449-460).
constexpr int QKV_LINES = 8;

__device__ __forceinline__ void qk_cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int DIM, int K, bool VT0>
__global__ void __launch_bounds__(256, 4)
    qk_interior_values_kernel(const DevParams P, const QkDecode D, int g, int L, int n1, int lpb, const QkLut* __restrict__ lut_g,
                              const double* __restrict__ tab_g, const u64* __restrict__ rowptr, double* __restrict__ values,
                              int fresh) {
  constexpr int N1 = K + 1, N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  extern __shared__ double kap[];  // [2 buffers][rows (j2 * 2 + j1)][KS] coefficients of the candidate cells of a line
  const int s = D.sbits[g];
  const bool vt[3] = {VT0, !((s >> 1) & 1), !((s >> 2) & 1)};  // vertex-type direction: two candidate cells
  const int N0 = P.N[0], KS = qk_kap_stride(N0);
  const int sz0 = (int)D.sz0[g], sz1 = (int)D.sz1[g];
  const int sh1 = vt[1] ? 1 : 0, sh2 = DIM == 3 && vt[2] ? 1 : 0;
  const int a2 = DIM == 3 ? (int)blockIdx.y + sh2 : 0, cb2 = DIM == 3 ? (int)blockIdx.y : 0;
  const int l0 = (int)blockIdx.x * lpb, l1 = min(n1, l0 + lpb);  // lines of this CTA: a1 = l + sh1
  const int lo0 = VT0 ? 1 : 0, hi0 = N0 - 1;
  const int R = (int)blockDim.x / L;  // row chunks of a line worked on side by side
  const int rsub = (int)threadIdx.x / L, slot = (int)threadIdx.x - rsub * L;
  const bool active = rsub < R;
  int shape = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++) shape |= (vt[d] ? 1 : 0) << d;
  const int off = active ? lut_g->slot2off[shape][slot] : 0;
  const int o[3] = {off & 7, (off >> 3) & 7, off >> 6};
  int e[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < DIM; d++) e[d] = o[d] - (vt[d] ? K : 1);
  // per direction: candidate cell (0: lower / only, 1: upper), local row index, local column index
  int nd[3] = {1, 1, 1}, dl[3][2] = {}, li[3][2] = {}, lj[3][2] = {};
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    int n = 0;
    if (vt[d]) {
      if (e[d] <= 0) { dl[d][n] = 0; li[d][n] = K; lj[d][n] = e[d] + K; n++; }
      if (e[d] >= 0) { dl[d][n] = 1; li[d][n] = 0; lj[d][n] = e[d]; n++; }
    } else {
      dl[d][n] = 0; li[d][n] = 1; lj[d][n] = e[d] + 1; n++;
    }
    nd[d] = n;
  }
  // per (y, z) candidate pair m = c1 + 2 c2: coefficient row and the table entries of the lower / upper (or only) x-cell
  double cL[4], cU[4];
  int ro[4];
  bool any[4];
#pragma unroll
  for (int c2 = 0; c2 < 2; c2++)
#pragma unroll
    for (int c1 = 0; c1 < 2; c1++) {
      const int m = c1 + 2 * c2;
      const bool on = active && c1 < nd[1] && c2 < nd[2];
      const int i12 = N1 * (li[1][c1] + N1 * (DIM == 3 ? li[2][c2] : 0));
      const int j12 = N1 * (lj[1][c1] + N1 * (DIM == 3 ? lj[2][c2] : 0));
      cL[m] = cU[m] = 0.0;
#pragma unroll
      for (int c0 = 0; c0 < 2; c0++)
        if (on && c0 < nd[0]) {
          const double tv = tab_g[(li[0][c0] + i12) * N + lj[0][c0] + j12];
          if (VT0 && dl[0][c0] == 0) cL[m] = tv; else cU[m] = tv;
        }
      ro[m] = on ? (dl[2][c2] * 2 + dl[1][c1]) * KS : 0;
      any[m] = __any_sync(0xffffffffu, on);
    }
  const int cn1 = vt[1] ? 2 : 1, cn2 = DIM == 3 ? (vt[2] ? 2 : 1) : 1;
  const int BS = ((cn2 - 1) * 2 + cn1) * KS;  // doubles per buffer: rows (j2 * 2 + j1)
  const bool ident = P.a_mode == PDB200_A_IDENTITY;
  if (ident)
    for (int i = threadIdx.x; i < 2 * BS; i += blockDim.x) kap[i] = 1.0;
  auto stage = [&](int l, int buf) {  // coefficient rows of line l -> buffer buf
    if (ident) return;
    double* kb = kap + buf * BS;
    for (int i = threadIdx.x; i < cn2 * cn1 * N0; i += blockDim.x) {
      const int x = i % N0, j1 = (i / N0) % cn1, j2 = i / (N0 * cn1);
      qk_cp_async8(kb + (j2 * 2 + j1) * KS + x, P.A + cell_index(P.N, x, l + j1, cb2 + j2));
    }
  };
  auto line_base = [&](int l) { return rowptr[D.start[g] + (u64)lo0 + (u64)sz0 * ((u64)(l + sh1) + (u64)sz1 * (u64)a2)]; };
  const int nrows = hi0 - lo0 + 1, chunk = (nrows + R - 1) / R;
  const int a_beg = lo0 + rsub * chunk, a_end = min(hi0 + 1, a_beg + chunk);
  stage(l0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  u64 base = line_base(l0);
  for (int l = l0; l < l1; l++) {
    const int buf = (l - l0) & 1;
    u64 next_base = 0;
    if (l + 1 < l1) {
      stage(l + 1, buf ^ 1);
      next_base = line_base(l + 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the rows of the next line has arrived
    __syncthreads();
    if (active && a_beg < a_end) {
      const double* kb = kap + buf * BS + a_beg;
      const double* kp0 = kb + ro[0];
      const double* kp1 = kb + ro[1];
      const double* kp2 = kb + ro[2];
      const double* kp3 = kb + ro[3];
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
      if (VT0) {
        if (any[0]) p0 = kp0[-1];
        if (any[1]) p1 = kp1[-1];
        if (any[2]) p2 = kp2[-1];
        if (any[3]) p3 = kp3[-1];
      }
      double* dst = values + base + (size_t)(a_beg - lo0) * L + slot;
#pragma unroll 4
      for (int a0 = a_beg; a0 < a_end; a0++, dst += L) {
        double v = 0.0;
        if (any[0]) { const double u = *kp0++; if (VT0) v = fma(cL[0], p0, v); v = fma(cU[0], u, v); p0 = u; }
        if (any[1]) { const double u = *kp1++; if (VT0) v = fma(cL[1], p1, v); v = fma(cU[1], u, v); p1 = u; }
        if (any[2]) { const double u = *kp2++; if (VT0) v = fma(cL[2], p2, v); v = fma(cU[2], u, v); p2 = u; }
        if (any[3]) { const double u = *kp3++; if (VT0) v = fma(cL[3], p3, v); v = fma(cU[3], u, v); p3 = u; }
        *dst = fresh ? v : *dst + v;
      }
    }
    __syncthreads();  // the buffer is restaged two lines on
    base = next_base;
  }
}

template <int G>
__global__ void __launch_bounds__(QK_THREADS)
    qk_mv_kernel(const QkDecode D, const QkLut* __restrict__ lut_g, const u64* __restrict__ rowptr, u64 nrows,
                 const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y,
                 const uint32_t* __restrict__ rowlist) {
  __shared__ QkLut lut;
  for (int i = threadIdx.x; i < (int)(sizeof(QkLut) / 4); i += QK_THREADS) ((uint32_t*)&lut)[i] = ((const uint32_t*)lut_g)[i];
  __syncthreads();
  const int lane = threadIdx.x % G;
  const u64 rix = ((u64)blockIdx.x * QK_THREADS + threadIdx.x) / G;
  const bool live = rix < nrows;
  const u64 row = live && rowlist ? (u64)rowlist[rix] : rix;  // with a list, nrows is its length
  double acc = 0.0;
  if (live) {
    const QkRow R = qk_row(D, (uint32_t)row);
    const u64 start = rowptr[row];
    for (int slot = lane; slot < R.len; slot += G) {
      const int off = lut.slot2off[R.shape][slot];
      const int q[3] = {R.lo[0] + (off & 7), R.lo[1] + ((off >> 3) & 7), R.lo[2] + (off >> 6)};
      acc = fma(values[start + slot], __ldg(x + qk_col_index(D, q)), acc);
    }
  }
  for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, G);
  if (live && lane == 0) y[row] = acc;
}

// y = A x for the interior rows, one x-line of rows per CTA (the SpMV twin of qk_interior_kernel):
// for interior rows the column of slot s is  col0(s) + (a0 - lo0),  so a lane keeps the col0 of its
// <= 4 slots (slot = lane + 32 j) in registers and the warp walks along the line: values are read
// fully coalesced, x is read as 125 (Q2, 3-D) unit-stride streams that live in L1, colidx is not read
// at all and nothing is decoded per entry.  One warp-shuffle reduction per row.
template <int DIM, int K>
__global__ void __launch_bounds__(256) qk_mv_interior_kernel(const DevParams P, const QkDecode D, int g, int L,
                                                             const QkLut* __restrict__ lut_g,
                                                             const u64* __restrict__ rowptr,
                                                             const double* __restrict__ values,
                                                             const double* __restrict__ x, double* __restrict__ y) {
  constexpr int MAXJ = 4;  // (2K+1)^3 = 125 <= 128 slots
  const int s = D.sbits[g];
  const bool vt[3] = {!(s & 1), !((s >> 1) & 1), !((s >> 2) & 1)};
  const int N0 = P.N[0];
  const int sz0 = (int)D.sz0[g], sz1 = (int)D.sz1[g];
  const int a1 = (int)blockIdx.x + (vt[1] ? 1 : 0), a2 = DIM == 3 ? (int)blockIdx.y + (vt[2] ? 1 : 0) : 0;
  const int lo0 = vt[0] ? 1 : 0, hi0 = N0 - 1;
  int shape = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++) shape |= (vt[d] ? 1 : 0) << d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  uint32_t col0[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; j++) {
    const int slot = lane + 32 * j;
    col0[j] = 0;
    if (slot < L) {
      const int off = lut_g->slot2off[shape][slot];
      const int o[3] = {off & 7, (off >> 3) & 7, off >> 6};
      int e[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < DIM; d++) e[d] = o[d] - (vt[d] ? K : 1);
      long long c;
      if (K == 1) {
        c = (lo0 + e[0]) + (long long)(N0 + 1) * ((a1 + e[1]) + (long long)(P.N[1] + 1) * (DIM == 3 ? a2 + e[2] : 0));
      } else {
        int par = 0, sh[3] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          const int q = ((s >> d) & 1) + e[d];  // lattice offset of the column relative to 2 a_d
          const int pb = q & 1;
          par |= pb << d;
          sh[d] = (q - pb) / 2;
        }
        const int g2 = D.group_of_s[par];
        c = (long long)D.start[g2] + (lo0 + sh[0]) +
            (long long)D.sz0[g2] * ((a1 + sh[1]) + (long long)D.sz1[g2] * (DIM == 3 ? a2 + sh[2] : 0));
      }
      col0[j] = (uint32_t)c;
    }
  }
  const u64 row0 = (u64)D.start[g] + (u64)lo0 + (u64)sz0 * ((u64)a1 + (u64)sz1 * (u64)a2);
  const u64 base = rowptr[row0];
  // four rows per pass: 8 KB of values per warp in flight
  const int nrows = hi0 - lo0 + 1;
  constexpr int U = 4;
  for (int r = warp; r < nrows; r += U * W) {
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      acc[u] = 0.0;
      const int ru = r + u * W;
      if (ru < nrows) {
        const double* __restrict__ vrow = values + base + (size_t)ru * L;
#pragma unroll
        for (int j = 0; j < MAXJ; j++)
          if (lane + 32 * j < L) acc[u] = fma(vrow[lane + 32 * j], __ldg(x + col0[j] + ru), acc[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; u++) acc[u] += __shfl_down_sync(0xffffffffu, acc[u], o);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < U; u++)
        if (r + u * W < nrows) y[row0 + r + u * W] = acc[u];
    }
  }
}

// Short rows (L <= 32 entries: Q1, and the cell-interior group of Q2): a warp per row would idle most
// lanes and pay a shuffle reduction per 9..27 products.  Here the values of a chunk of rows are staged
// in shared memory with coalesced loads, then ONE THREAD PER ROW runs over its L entries: the shared
// reads are conflict-free (row stride L is odd) and for a fixed slot consecutive threads read
// consecutive x (col0(slot) + r), so the x gather is coalesced too.
// PARTS > 1 (long rows, up to 125 entries): PARTS threads share a row (slots part, part + PARTS, ...), the
// chunk holds 256 / PARTS rows so that it still fits shared memory, partial sums meet in shared memory.
template <int DIM, int K, int PARTS>
__global__ void __launch_bounds__(256) qk_mv_interior_short_kernel(const DevParams P, const QkDecode D, int g, int L,
                                                                   const QkLut* __restrict__ lut_g,
                                                                   const u64* __restrict__ rowptr,
                                                                   const double* __restrict__ values,
                                                                   const double* __restrict__ x, double* __restrict__ y) {
  constexpr int ROWS = 256 / PARTS;
  extern __shared__ double vs[];  // [ROWS][L]
  __shared__ uint32_t col0s[128];
  __shared__ double part_sum[PARTS > 1 ? 256 : 1];
  const int s = D.sbits[g];
  const bool vt[3] = {!(s & 1), !((s >> 1) & 1), !((s >> 2) & 1)};
  const int N0 = P.N[0];
  const int sz0 = (int)D.sz0[g], sz1 = (int)D.sz1[g];
  const int a1 = (int)blockIdx.x + (vt[1] ? 1 : 0), a2 = DIM == 3 ? (int)blockIdx.y + (vt[2] ? 1 : 0) : 0;
  const int lo0 = vt[0] ? 1 : 0, hi0 = N0 - 1;
  int shape = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++) shape |= (vt[d] ? 1 : 0) << d;
  if ((int)threadIdx.x < L) {
    const int off = lut_g->slot2off[shape][threadIdx.x];
    const int o[3] = {off & 7, (off >> 3) & 7, off >> 6};
    int e[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; d++) e[d] = o[d] - (vt[d] ? K : 1);
    long long c;
    if (K == 1) {
      c = (lo0 + e[0]) + (long long)(N0 + 1) * ((a1 + e[1]) + (long long)(P.N[1] + 1) * (DIM == 3 ? a2 + e[2] : 0));
    } else {
      int par = 0, sh[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const int q = ((s >> d) & 1) + e[d];
        const int pb = q & 1;
        par |= pb << d;
        sh[d] = (q - pb) / 2;
      }
      const int g2 = D.group_of_s[par];
      c = (long long)D.start[g2] + (lo0 + sh[0]) +
          (long long)D.sz0[g2] * ((a1 + sh[1]) + (long long)D.sz1[g2] * (DIM == 3 ? a2 + sh[2] : 0));
    }
    col0s[threadIdx.x] = (uint32_t)c;
  }
  const u64 row0 = (u64)D.start[g] + (u64)lo0 + (u64)sz0 * ((u64)a1 + (u64)sz1 * (u64)a2);
  const u64 base = rowptr[row0];
  const int nrows = hi0 - lo0 + 1;
  // the line is cut into gridDim.z pieces of whole ROWS-row chunks
  const int chunks = (nrows + ROWS - 1) / ROWS, per = (chunks + (int)gridDim.z - 1) / (int)gridDim.z;
  const int rl = threadIdx.x % ROWS, part = threadIdx.x / ROWS;
  for (int ch = (int)blockIdx.z * per; ch < min(chunks, ((int)blockIdx.z + 1) * per); ch++) {
    const int r0 = ch * ROWS, nr = min(ROWS, nrows - r0);
    __syncthreads();  // col0s ready / previous chunk consumed
    const double* __restrict__ src = values + base + (size_t)r0 * L;
    for (int i = threadIdx.x; i < nr * L; i += 256) vs[i] = src[i];
    __syncthreads();
    double acc = 0.0;
    if (rl < nr) {
      const int r = r0 + rl;
      const double* v = vs + rl * L;
      for (int slot = part; slot < L; slot += PARTS) acc = fma(v[slot], __ldg(x + col0s[slot] + r), acc);
    }
    if (PARTS == 1) {
      if (rl < nr) y[row0 + r0 + rl] = acc;
    } else {
      part_sum[threadIdx.x] = acc;
      __syncthreads();
      if (part == 0 && rl < nr) {
#pragma unroll
        for (int q = 1; q < PARTS; q++) acc += part_sum[q * ROWS + rl];
        y[row0 + r0 + rl] = acc;
      }
    }
  }
}

// Long rows (45 / 75 / 125 entries: the face, edge and vertex groups of Q2 in 3-D, 95 % of the non-zeros): the values of
// 32 consecutive rows of the line are ONE contiguous piece of the array, so a stage of the ring is ONE bulk copy
// (cp.async.bulk -> mbarrier) issued by one thread; three stages are in flight per CTA while the previous ones are
// consumed, which is what the single-buffered version above lacks (it waits for its chunk behind a barrier:
// long-scoreboard bound at 27 % of the HBM peak, profiles/r01_spmv_staged_ncu_summary.json; the warp-per-row kernel
// gathers x with 32 sectors per load instead).  Consumers: thread = (row of the chunk, eighth of the slots), all its
// <= 16 x loads in flight at once; shared reads are conflict-free (odd row stride), for a fixed slot the 32 lanes read
// 32 consecutive x, the eight partial sums of a row meet in shared memory.  Bulk copies need 16-byte aligned sources and row starts are odd as often as even
// (all row lengths are odd): the copy starts one entry early where needed and the consumers skip it.
constexpr int MVP_ROWS = 32, MVP_PARTS = 8, MVP_STAGES = 3, MVP_THREADS = MVP_ROWS * MVP_PARTS;
__device__ __forceinline__ uint32_t mv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int DIM, int K, int L>
__global__ void __launch_bounds__(MVP_THREADS) qk_mv_interior_pipe_kernel(const DevParams P, const QkDecode D, int g,
                                                                          const QkLut* __restrict__ lut_g,
                                                                          const u64* __restrict__ rowptr, u64 nnz_total,
                                                                          const double* __restrict__ values,
                                                                          const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(16) double vs[];  // [STAGES][ROWS * L + 2]
  __shared__ uint32_t col0s[128];
  __shared__ double part_sum[2][MVP_THREADS];
  __shared__ __align__(8) uint64_t full[MVP_STAGES];
  const int s = D.sbits[g];
  const bool vt[3] = {!(s & 1), !((s >> 1) & 1), !((s >> 2) & 1)};
  const int N0 = P.N[0];
  const int sz0 = (int)D.sz0[g], sz1 = (int)D.sz1[g];
  const int a1 = (int)blockIdx.x + (vt[1] ? 1 : 0), a2 = DIM == 3 ? (int)blockIdx.y + (vt[2] ? 1 : 0) : 0;
  const int lo0 = vt[0] ? 1 : 0, hi0 = N0 - 1;
  int shape = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++) shape |= (vt[d] ? 1 : 0) << d;
  if ((int)threadIdx.x < L) {
    const int off = lut_g->slot2off[shape][threadIdx.x];
    const int o[3] = {off & 7, (off >> 3) & 7, off >> 6};
    int e[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; d++) e[d] = o[d] - (vt[d] ? K : 1);
    long long c;
    if (K == 1) {
      c = (lo0 + e[0]) + (long long)(N0 + 1) * ((a1 + e[1]) + (long long)(P.N[1] + 1) * (DIM == 3 ? a2 + e[2] : 0));
    } else {
      int par = 0, sh[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const int q = ((s >> d) & 1) + e[d];
        const int pb = q & 1;
        par |= pb << d;
        sh[d] = (q - pb) / 2;
      }
      const int g2 = D.group_of_s[par];
      c = (long long)D.start[g2] + (lo0 + sh[0]) +
          (long long)D.sz0[g2] * ((a1 + sh[1]) + (long long)D.sz1[g2] * (DIM == 3 ? a2 + sh[2] : 0));
    }
    col0s[threadIdx.x] = (uint32_t)c;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < MVP_STAGES; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mv_smem_u32(&full[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const u64 row0 = (u64)D.start[g] + (u64)lo0 + (u64)sz0 * ((u64)a1 + (u64)sz1 * (u64)a2);
  const u64 base = rowptr[row0];
  const int nrows = hi0 - lo0 + 1;
  // the line is cut into gridDim.z pieces of whole chunks
  const int chunks_all = (nrows + MVP_ROWS - 1) / MVP_ROWS, per = (chunks_all + (int)gridDim.z - 1) / (int)gridDim.z;
  const int ch0 = (int)blockIdx.z * per, ch1 = min(chunks_all, ch0 + per);
  const int stage_doubles = MVP_ROWS * L + 2;
  // chunk ch: entries [c0, c0 + nr * L) of the array; the copy covers [c0 - par, ...) rounded up to 16 bytes.  A chunk
  // that a bulk copy cannot move (the rounded-up window would end behind the array, or the array is not 16-byte
  // aligned) is loaded by the threads themselves; every thread evaluates the same predicate.
  const bool unaligned = ((uintptr_t)values & 15) != 0;
  auto window = [&](int ch, u64& a, u64& len) {
    const int r0 = ch * MVP_ROWS, nr = min(MVP_ROWS, nrows - r0);
    const u64 c0 = base + (u64)r0 * L;
    a = c0 & ~(u64)1;
    len = (((c0 - a) + (u64)nr * L) + 1) & ~(u64)1;
    return !(unaligned || a + len > nnz_total);
  };
  auto issue = [&](int ch) {  // thread 0 only
    u64 a, len;
    if (!window(ch, a, len)) return;
    const uint32_t bytes = (uint32_t)(len * 8);
    const int st = (ch - ch0) % MVP_STAGES;
    const uint32_t bar = mv_smem_u32(&full[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     mv_smem_u32(vs + (size_t)st * stage_doubles)),
                 "l"(values + a), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (threadIdx.x == 0)
    for (int ch = ch0; ch < min(ch1, ch0 + MVP_STAGES); ch++) issue(ch);
  const int rl = threadIdx.x % MVP_ROWS, part = threadIdx.x / MVP_ROWS;
  constexpr int NJ = (L + MVP_PARTS - 1) / MVP_PARTS;
  uint32_t cr[NJ];  // the columns of this thread's slots (slot = part + 8 j) relative to the row
#pragma unroll
  for (int j = 0; j < NJ; j++) cr[j] = col0s[min(part + j * MVP_PARTS, L - 1)];
  unsigned phase = 0;  // bit st: parity of the next completion of full[st]
  for (int ch = ch0; ch < ch1; ch++) {
    const int it = ch - ch0, st = it % MVP_STAGES;
    const int r0 = ch * MVP_ROWS, nr = min(MVP_ROWS, nrows - r0);
    const u64 c0 = base + (u64)r0 * L;
    const int par = (int)(c0 & 1);
    double* __restrict__ stage = vs + (size_t)st * stage_doubles;
    // the x values first: all loads of the thread in flight before anything waits (the "+d" list below keeps the
    // compiler from sinking them between the multiply-adds)
    double xv[NJ], vv[NJ];
    {
      const double* __restrict__ xr = x + (r0 + min(rl, nr - 1));
#pragma unroll
      for (int j = 0; j < NJ; j++) xv[j] = __ldg(xr + cr[j]);
    }
    u64 wa, wl;
    if (!window(ch, wa, wl)) {
      for (int i = threadIdx.x; i < nr * L; i += MVP_THREADS) stage[par + i] = values[c0 + i];
      __syncthreads();
    } else {
      const uint32_t bar = mv_smem_u32(&full[st]), parity = (phase >> st) & 1u;
      phase ^= 1u << st;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "MVWAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra MVDONE;\n\t"
          "bra MVWAIT;\n\t"
          "MVDONE:\n\t"
          "}" ::"r"(bar),
          "r"(parity)
          : "memory");
    }
    {
      const double* __restrict__ v = stage + par + min(rl, nr - 1) * L;
#pragma unroll
      for (int j = 0; j < NJ; j++) vv[j] = v[min(part + j * MVP_PARTS, L - 1)];
    }
    if constexpr (NJ == 16)
      asm volatile("" : "+d"(xv[0]), "+d"(xv[1]), "+d"(xv[2]), "+d"(xv[3]), "+d"(xv[4]), "+d"(xv[5]), "+d"(xv[6]), "+d"(xv[7]),
                        "+d"(xv[8]), "+d"(xv[9]), "+d"(xv[10]), "+d"(xv[11]), "+d"(xv[12]), "+d"(xv[13]), "+d"(xv[14]), "+d"(xv[15]));
    else if constexpr (NJ == 10)
      asm volatile("" : "+d"(xv[0]), "+d"(xv[1]), "+d"(xv[2]), "+d"(xv[3]), "+d"(xv[4]), "+d"(xv[5]), "+d"(xv[6]), "+d"(xv[7]),
                        "+d"(xv[8]), "+d"(xv[9]));
    else
      asm volatile("" : "+d"(xv[0]), "+d"(xv[1]), "+d"(xv[2]), "+d"(xv[3]), "+d"(xv[4]), "+d"(xv[5]));
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; j += 2) {
      if (part + j * MVP_PARTS < L) acc0 = fma(vv[j], xv[j], acc0);
      if (j + 1 < NJ && part + (j + 1) * MVP_PARTS < L) acc1 = fma(vv[j + 1], xv[j + 1], acc1);
    }
    part_sum[it & 1][threadIdx.x] = acc0 + acc1;
    __syncthreads();  // the stage is consumed, the partial sums are there (and those of chunk ch - 1 have been read)
    if (threadIdx.x == 0 && ch + MVP_STAGES < ch1) issue(ch + MVP_STAGES);
    if (part == 0 && rl < nr) {
      double a = part_sum[it & 1][rl];
#pragma unroll
      for (int q = 1; q < MVP_PARTS; q++) a += part_sum[it & 1][q * MVP_ROWS + rl];
      y[row0 + r0 + rl] = a;
    }
  }
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
  double A_s[3][3], b_s[3];
  load_A_cell(P, e, A_s);
  const bool pwA = pw_A(P);  // permeabilityIsConstantPerCell() == false
  const long long stride[3] = {1, (long long)P.N[0], (long long)P.N[0] * P.N[1]};
  double v = 0.0;
  if (face < 0) {
    // jacobian_volume, convectiondiffusiondg.hh:199-266
    for (int q = 0; q < P.nq; q++) {
      if (pwA) load_A_at(P, e, q, A_s);       // :234-237
      load_b(P, e, q, b_s);                   // :255
      const double c_s = load_c(P, e, q);     // :258
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
      if (pwA) load_A_cell(P, e, A_s);
      for (int d = 0; d < 3; d++) An_s[d] = A_s[d][dir] * nsign;
      const double area = P.area[dir];
      if (!onb) {
        // jacobian_skeleton, :484-669, this cell's rows (ss / sn when it is the inside cell,
        // nn / ns when it is the outside cell)
        const long long other = e + (side ? stride[dir] : -stride[dir]);
        double A_o[3][3], An_o[3], b_F[3];
        load_A_cell(P, other, A_o);
        for (int d = 0; d < 3; d++) An_o[d] = A_o[d][dir] * nsign;
        const long long bc = side ? other : e;  // velocity of the larger-index cell on its lower face (:613)
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
        double penalty = (P.alpha / h_F) * harm * P.k * (P.k + dim - 1);
        for (int q = 0; q < P.nfq; q++) {
          if (pwA) {  // :575-590
            load_A_at(P, e, face_pt(P, dir, side, q), A_s);
            load_A_at(P, other, face_pt(P, dir, 1 - side, q), A_o);
            for (int d = 0; d < 3; d++) {
              An_s[d] = A_s[d][dir] * nsign;
              An_o[d] = A_o[d][dir] * nsign;
            }
            if (P.weights_on) {
              const double ds = An_s[dir] * nsign, dn = An_o[dir] * nsign;
              omega_s = dn / (ds + dn + 1e-20);
              omega_o = ds / (ds + dn + 1e-20);
              harm = 2.0 * ds * dn / (ds + dn + 1e-20);
              penalty = (P.alpha / h_F) * harm * P.k * (P.k + dim - 1);
            }
          }
          load_b(P, bc, face_pt(P, dir, 0, q), b_F);
          const double betan = b_F[dir] * nsign;
          const bool take_self = side == 0 ? (betan >= 0.0) : !((-betan) >= 0.0);
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
        const double h_F = P.vol / area;
        double harm = P.weights_on ? An_s[dir] * nsign : 1.0;
        double penalty = (P.alpha / h_F) * harm * P.k * (P.k + dim - 1);
        for (int q = 0; q < P.nfq; q++) {
          if (pwA) {  // :967-977
            load_A_at(P, e, face_pt(P, dir, side, q), A_s);
            for (int d = 0; d < 3; d++) An_s[d] = A_s[d][dir] * nsign;
            if (P.weights_on) {
              harm = An_s[dir] * nsign;
              penalty = (P.alpha / h_F) * harm * P.k * (P.k + dim - 1);
            }
          }
          const int bctype = load_bctype(P, bf, q);  // :979
          if (bctype == PDB200_BC_NONE || bctype == PDB200_BC_NEUMANN) continue;
          load_b(P, e, face_pt(P, dir, side, q), b_s);  // :995
          const double betan = b_s[dir] * nsign;
          if (bctype == PDB200_BC_OUTFLOW && betan < -1e-30) {
            *errflag = 1;
            continue;
          }
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
  QkDecode D;
  QkTables Q;
  double* tables = nullptr;  // Qk: unit local matrices (device)
  QkLut* lut = nullptr;      // Qk: slot decode tables (device)
  u64 nrows = 0, nnz = 0, nbrows = 0, nblocks = 0;
  u64* rowptr = nullptr;           // Qk: scalar CSR row pointers; DG: block row pointers (cells)
  unsigned char* flags = nullptr;  // Qk: constrained rows
  uint32_t* brows = nullptr;       // Qk: rows whose lattice point lies on the boundary (ascending)
  u64 nbr = 0;
  bool interior_ok = false;        // Qk: the interior-line kernels apply (every group has interior rows)
};

static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.magic = 0;
  f.more = -1;
  if (d <= 1) return f;
  int L = 31;
  while (!((d >> L) & 1)) L--;  // floor(log2 d)
  if ((d & (d - 1)) == 0) {
    f.more = L - 1;
    return f;
  }
  const unsigned __int128 num = (unsigned __int128)1 << (32 + L);
  uint64_t m = (uint64_t)(num / d);
  const uint64_t rem = (uint64_t)(num - (unsigned __int128)m * d);
  m += m;
  const uint64_t twice_rem = rem + rem;
  if (twice_rem >= d) m += 1;
  f.magic = (uint32_t)(m + 1);
  f.more = L;
  return f;
}

static QkDecode make_qk_decode(const DevParams& P, const QkLayout& L) {
  if ((unsigned long long)L.ndofs >= 0xffffffffull) throw Error("conforming matrix path: needs fewer than 2^32 DOFs");
  QkDecode D;
  std::memset(&D, 0, sizeof(D));
  D.dim = P.dim;
  D.k = P.k;
  for (int d = 0; d < 3; d++) D.N[d] = P.N[d];
  auto set_group = [&](int g, int s, uint32_t start) {
    D.start[g] = start;
    D.sbits[g] = (uint8_t)s;
    D.group_of_s[s] = (uint8_t)g;
    const uint32_t z0 = ((s >> 0) & 1) ? P.N[0] : P.N[0] + 1;
    const uint32_t z1 = P.dim > 1 ? (((s >> 1) & 1) ? P.N[1] : P.N[1] + 1) : 1;
    D.sz0[g] = z0;
    D.sz1[g] = z1;
    D.d0[g] = make_fastdiv(z0);
    D.d1[g] = make_fastdiv(z1);
  };
  if (P.k == 1) {
    D.ng = 1;
    set_group(0, 0, 0);
  } else {
    int g = 0;
    for (int edim = 0; edim <= P.dim; edim++)
      for (int s = 0; s < (1 << P.dim); s++) {
        int pc = 0;
        for (int d = 0; d < P.dim; d++) pc += (s >> d) & 1;
        if (pc == edim) set_group(g++, s, (uint32_t)(L.block_off[edim] + L.group_off[s]));
      }
    D.ng = g;
  }
  D.start[D.ng] = (uint32_t)L.ndofs;
  // self-check of the multipliers on the values that occur
  for (int g = 0; g < D.ng; g++)
    for (uint32_t n : {0u, 1u, D.sz0[g] - 1, D.sz0[g], D.sz0[g] + 1, 0x7fffffffu, 0xfffffffeu, (uint32_t)L.ndofs}) {
      auto hostdiv = [](uint32_t n, const FastDiv& f) -> uint32_t {
        if (f.more < 0) return n;
        const uint32_t q = (uint32_t)(((uint64_t)f.magic * n) >> 32);
        return (((n - q) >> 1) + q) >> f.more;
      };
      if (hostdiv(n, D.d0[g]) != n / D.sz0[g] || hostdiv(n, D.d1[g]) != n / D.sz1[g]) throw Error("fast division self-check failed");
    }
  return D;
}

// slot decode tables and unit local matrices of the conforming space (host side, one-off)
static void build_qk_tables(MatrixPlan* plan) {
  const DevParams& P = plan->P;
  const int dim = P.dim, k = P.k, n1 = P.n1, n = P.n;
  // --- LUT: columns of a box sorted by container index.  k = 1: lexicographic.  k = 2: by entity
  // dimension (number of odd offsets), then extension bitset, then lexicographic anchors.
  std::vector<QkLut> lut(1);
  std::memset(&lut[0], 0, sizeof(QkLut));
  for (int sh = 0; sh < 8; sh++) {
    int len[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) len[d] = ((sh >> d) & 1) ? 2 * k + 1 : k + 1;
    if (dim == 2 && (sh & 4)) continue;
    struct Item { int key; int d[3]; };
    std::vector<Item> items;
    for (int d2 = 0; d2 < len[2]; d2++)
      for (int d1 = 0; d1 < len[1]; d1++)
        for (int d0 = 0; d0 < len[0]; d0++) {
          Item it;
          it.d[0] = d0; it.d[1] = d1; it.d[2] = d2;
          int sbits = 0, edim = 0;
          if (k == 2) {
            sbits = (d0 & 1) | ((d1 & 1) << 1) | ((d2 & 1) << 2);
            edim = (d0 & 1) + (d1 & 1) + (d2 & 1);
          }
          // anchors are offsets / 2 inside the group: lexicographic order of the offsets is the same
          it.key = ((edim * 8 + sbits) * 8 + d2) * 64 + d1 * 8 + d0;
          items.push_back(it);
        }
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
    for (size_t slot = 0; slot < items.size(); slot++) {
      const int* d = items[slot].d;
      lut[0].slot2off[sh][slot] = (uint16_t)(d[0] | (d[1] << 3) | (d[2] << 6));
      lut[0].off2slot[sh][d[0] + 5 * (d[1] + 5 * d[2])] = (uint8_t)slot;
    }
  }
  PDB_CUDA(cudaMalloc(&plan->lut, sizeof(QkLut)));
  PDB_CUDA(cudaMemcpy(plan->lut, lut.data(), sizeof(QkLut), cudaMemcpyHostToDevice));
  // --- unit local matrices
  QkTables& Q = plan->Q;
  Q.nt = 0;
  auto add = [&](int kind, int a, int b) {
    Q.kind[Q.nt] = kind;
    Q.ia[Q.nt] = a;
    Q.ib[Q.nt] = b;
    Q.nt++;
  };
  if (P.a_mode == PDB200_A_IDENTITY || P.a_mode == PDB200_A_SCALAR) add(0, 0, 0);
  else if (P.a_mode == PDB200_A_DIAGONAL)
    for (int a = 0; a < dim; a++) add(1, a, a);
  else
    for (int a = 0; a < dim; a++)
      for (int b = 0; b < dim; b++) add(1, a, b);
  if (P.b)
    for (int a = 0; a < dim; a++) add(2, a, 0);
  if (P.c) add(3, 0, 0);
  const Mat1D& T = plan->T;
  std::vector<double> tab((size_t)Q.nt * n * n, 0.0);
  for (int t = 0; t < Q.nt; t++)
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        int li[3] = {0, 0, 0}, lj[3] = {0, 0, 0}, ii = i, jj = j;
        for (int d = 0; d < dim; d++) {
          li[d] = ii % n1; ii /= n1;
          lj[d] = jj % n1; jj /= n1;
        }
        auto stiff = [&](int a, int b) {  // int (d_b phi_j)(d_a phi_i) * |K|
          long double v = P.vol * P.ih[a] * P.ih[b];
          for (int d = 0; d < dim; d++) {
            if (a == b) v *= d == a ? T.K[li[d] * n1 + lj[d]] : T.M[li[d] * n1 + lj[d]];
            else if (d == a) v *= T.C[li[d] * n1 + lj[d]];
            else if (d == b) v *= T.C[lj[d] * n1 + li[d]];
            else v *= T.M[li[d] * n1 + lj[d]];
          }
          return v;
        };
        long double v = 0;
        if (Q.kind[t] == 0) {
          for (int a = 0; a < dim; a++) v += stiff(a, a);
        } else if (Q.kind[t] == 1) {
          v = stiff(Q.ia[t], Q.ib[t]);
        } else if (Q.kind[t] == 2) {  // int phi_j d_a phi_i * |K|   (enters with weight -b_a)
          v = P.vol * P.ih[Q.ia[t]];
          for (int d = 0; d < dim; d++) v *= d == Q.ia[t] ? T.C[li[d] * n1 + lj[d]] : T.M[li[d] * n1 + lj[d]];
        } else {
          v = P.vol;
          for (int d = 0; d < dim; d++) v *= T.M[li[d] * n1 + lj[d]];
        }
        tab[((size_t)t * n + i) * n + j] = (double)v;
      }
  PDB_CUDA(cudaMalloc(&plan->tables, tab.size() * sizeof(double)));
  PDB_CUDA(cudaMemcpy(plan->tables, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
}

// rows handled by the generic kernels when the interior-line kernels take the rest
static void build_qk_boundary_rows(MatrixPlan* plan) {
  const DevParams& P = plan->P;
  const QkDecode& D = plan->D;
  plan->interior_ok = true;
  for (int d = 0; d < P.dim; d++)
    if (P.N[d] < 2) plan->interior_ok = false;  // no interior vertex in that direction
  if ((size_t)4 * qk_kap_stride(P.N[0]) * sizeof(double) > 200 * 1024) plan->interior_ok = false;  // coefficient rows of a line in smem
  if (!plan->interior_ok) return;
  std::vector<uint32_t> rows;
  for (int g = 0; g < D.ng; g++) {
    const int s = D.sbits[g];
    const bool vt[3] = {!(s & 1), !((s >> 1) & 1), !((s >> 2) & 1)};
    int sz[3] = {1, 1, 1};
    for (int d = 0; d < P.dim; d++) sz[d] = vt[d] ? P.N[d] + 1 : P.N[d];
    for (int a2 = 0; a2 < sz[2]; a2++)
      for (int a1 = 0; a1 < sz[1]; a1++) {
        const bool line_b = (P.dim > 1 && vt[1] && (a1 == 0 || a1 == sz[1] - 1)) ||
                            (P.dim > 2 && vt[2] && (a2 == 0 || a2 == sz[2] - 1));
        const uint32_t r0 = D.start[g] + (uint32_t)sz[0] * ((uint32_t)a1 + (uint32_t)sz[1] * (uint32_t)a2);
        if (line_b) {
          for (int a0 = 0; a0 < sz[0]; a0++) rows.push_back(r0 + a0);
        } else if (vt[0]) {
          rows.push_back(r0);
          rows.push_back(r0 + sz[0] - 1);
        }
      }
  }
  std::sort(rows.begin(), rows.end());
  plan->nbr = rows.size();
  if (plan->nbr) {
    PDB_CUDA(cudaMalloc(&plan->brows, rows.size() * sizeof(uint32_t)));
    PDB_CUDA(cudaMemcpy(plan->brows, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
}

// launches the interior-line kernel for every entity group; returns the number of launches
template <int DIM, int K, typename IDX, bool VALUES>
static int launch_qk_interior(MatrixPlan* p, double* values, IDX* colidx, bool fresh, cudaStream_t s) {
  const DevParams& P = p->P;
  const QkDecode& D = p->D;
  int launches = 0;
  for (int g = 0; g < D.ng; g++) {
    const int sb = D.sbits[g];
    const bool vt[3] = {!(sb & 1), !((sb >> 1) & 1), !((sb >> 2) & 1)};
    int L = 1;
    for (int d = 0; d < DIM; d++) L *= vt[d] ? 2 * K + 1 : K + 1;
    const int n1 = vt[1] ? P.N[1] - 1 : P.N[1], n2 = DIM == 3 ? (vt[2] ? P.N[2] - 1 : P.N[2]) : 1;
    const int nrow0 = vt[0] ? P.N[0] - 1 : P.N[0];
    if (n1 <= 0 || n2 <= 0 || nrow0 <= 0) continue;
    const int threads = 256;
    const size_t smem2 = (size_t)2 * ((DIM == 3 && vt[2] ? 2 : 0) + (vt[1] ? 2 : 1)) * qk_kap_stride(P.N[0]) * sizeof(double);
    if (VALUES && smem2 <= 160 * 1024) {
      // lines per CTA: 8 on a full machine, fewer on small grids so that every SM still gets a few CTAs
      int lpb = (int)std::max<long long>(1, std::min<long long>(QKV_LINES, (long long)n1 * n2 / (4 * 148)));
      if (const char* e = getenv("PDB200_QKV_LINES")) lpb = std::max(1, atoi(e));  // tests: the multi-line walk on small grids
      const dim3 grid((n1 + lpb - 1) / lpb, n2);
      if (vt[0]) {
        if (smem2 > 48 * 1024)
          PDB_CUDA(cudaFuncSetAttribute(qk_interior_values_kernel<DIM, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        qk_interior_values_kernel<DIM, K, true><<<grid, threads, smem2, s>>>(P, D, g, L, n1, lpb, p->lut, p->tables, p->rowptr, values,
                                                                           fresh ? 1 : 0);
      } else {
        if (smem2 > 48 * 1024)
          PDB_CUDA(cudaFuncSetAttribute(qk_interior_values_kernel<DIM, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        qk_interior_values_kernel<DIM, K, false><<<grid, threads, smem2, s>>>(P, D, g, L, n1, lpb, p->lut, p->tables, p->rowptr, values,
                                                                            fresh ? 1 : 0);
      }
    } else if (VALUES) {  // lines too long for two coefficient buffers: one line per CTA
      const size_t smem = (size_t)4 * qk_kap_stride(P.N[0]) * sizeof(double);
      if (smem > 48 * 1024)
        PDB_CUDA(cudaFuncSetAttribute(qk_interior_kernel<DIM, K, IDX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      qk_interior_kernel<DIM, K, IDX, true><<<dim3(n1, n2), threads, smem, s>>>(P, D, g, L, p->lut, p->tables, p->rowptr, values,
                                                                              colidx, fresh ? 1 : 0);
    } else {
      qk_interior_kernel<DIM, K, IDX, false><<<dim3(n1, n2), threads, 0, s>>>(P, D, g, L, p->lut, p->tables, p->rowptr, values,
                                                                             colidx, 0);
    }
    PDB_CUDA(cudaGetLastError());
    launches++;
  }
  return launches;
}

template <int DIM, int K>
static int launch_qk_mv_interior(MatrixPlan* p, const double* values, const double* x, double* y, cudaStream_t s) {
  const DevParams& P = p->P;
  const QkDecode& D = p->D;
  int launches = 0;
  for (int g = 0; g < D.ng; g++) {
    const int sb = D.sbits[g];
    const bool vt[3] = {!(sb & 1), !((sb >> 1) & 1), !((sb >> 2) & 1)};
    int L = 1;
    for (int d = 0; d < DIM; d++) L *= vt[d] ? 2 * K + 1 : K + 1;
    const int n1 = vt[1] ? P.N[1] - 1 : P.N[1], n2 = DIM == 3 ? (vt[2] ? P.N[2] - 1 : P.N[2]) : 1;
    const int nrow0 = vt[0] ? P.N[0] - 1 : P.N[0];
    if (n1 <= 0 || n2 <= 0 || nrow0 <= 0) continue;
    // long lines (2-D grids) are cut along x so that the launch fills the machine
    const long long lines = (long long)n1 * n2;
    static const bool mv_pipe = [] { const char* e = getenv("PDB200_MV_PIPE"); return !(e && e[0] == '0'); }();
    if (L <= 32) {
      const size_t smem = (size_t)256 * L * sizeof(double);
      if (smem > 48 * 1024)
        PDB_CUDA(cudaFuncSetAttribute(qk_mv_interior_short_kernel<DIM, K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
      const int chunks = (nrow0 + 255) / 256;
      const int nz = (int)std::max<long long>(1, std::min<long long>(chunks, (148 * 8 + lines - 1) / lines));
      qk_mv_interior_short_kernel<DIM, K, 1><<<dim3(n1, n2, nz), 256, smem, s>>>(P, D, g, L, p->lut, p->rowptr, values, x, y);
    } else if (mv_pipe && DIM == 3 && K == 2 && (L == 45 || L == 75 || L == 125)) {  // ring of bulk copies
      const size_t smem = (size_t)MVP_STAGES * (MVP_ROWS * L + 2) * sizeof(double);
      const int chunks = (nrow0 + MVP_ROWS - 1) / MVP_ROWS;
      const int nz = (int)std::max<long long>(1, std::min<long long>(chunks, (148 * 8 + lines - 1) / lines));
#define PDB_MV_PIPE(LL)                                                                                                      \
  do {                                                                                                                       \
    PDB_CUDA(cudaFuncSetAttribute(qk_mv_interior_pipe_kernel<3, 2, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    qk_mv_interior_pipe_kernel<3, 2, LL><<<dim3(n1, n2, nz), MVP_THREADS, smem, s>>>(P, D, g, p->lut, p->rowptr, p->nnz, values, x, y); \
  } while (0)
      if (L == 45) PDB_MV_PIPE(45);
      else if (L == 75) PDB_MV_PIPE(75);
      else PDB_MV_PIPE(125);
#undef PDB_MV_PIPE
    } else if (L > 64) {  // 75 / 125 entries: the chunk would cap the occupancy; one warp per row instead
      qk_mv_interior_kernel<DIM, K><<<dim3(n1, n2), 256, 0, s>>>(P, D, g, L, p->lut, p->rowptr, values, x, y);
    } else {
      const size_t smem = (size_t)64 * L * sizeof(double);
      if (smem > 48 * 1024)
        PDB_CUDA(cudaFuncSetAttribute(qk_mv_interior_short_kernel<DIM, K, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
      const int chunks = (nrow0 + 63) / 64;
      const int nz = (int)std::max<long long>(1, std::min<long long>(chunks, (148 * 8 + lines - 1) / lines));
      qk_mv_interior_short_kernel<DIM, K, 4><<<dim3(n1, n2, nz), 256, smem, s>>>(P, D, g, L, p->lut, p->rowptr, values, x, y);
    }
    PDB_CUDA(cudaGetLastError());
    launches++;
  }
  return launches;
}

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
      plan->D = make_qk_decode(P, plan->L);
      QkRowLen f{plan->D};
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
      build_qk_tables(plan);
      build_qk_boundary_rows(plan);
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
  if (p->brows) cudaFree(p->brows);
  if (p->tables) cudaFree(p->tables);
  if (p->lut) cudaFree(p->lut);
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
    const int G = P.n > 16 ? 32 : (P.n > 8 ? 16 : 8);
    // interior rows: one x-line per CTA; boundary rows (or everything on degenerate grids): generic kernel
    const uint32_t* list = p->interior_ok ? p->brows : nullptr;
    const u64 ngen = p->interior_ok ? p->nbr : nr;
    if (p->interior_ok) {
#define PDB_INT(DD, KK)                                                                                   \
  if (P.dim == DD && P.k == KK)                                                                           \
    launches += col32 ? launch_qk_interior<DD, KK, uint32_t, false>(p, nullptr, (uint32_t*)ci, false, s)  \
                      : launch_qk_interior<DD, KK, u64, false>(p, nullptr, (u64*)ci, false, s);
      PDB_INT(2, 1) PDB_INT(2, 2) PDB_INT(3, 1) PDB_INT(3, 2)
#undef PDB_INT
    }
    const u64 blocks = (ngen * G + QK_THREADS - 1) / QK_THREADS;
#define PDB_COLIDX(GG)                                                                                        \
  if (col32) qk_colidx_kernel<GG, uint32_t><<<(unsigned)blocks, QK_THREADS, 0, s>>>(p->D, p->lut, p->rowptr, ngen, (uint32_t*)ci, list); \
  else qk_colidx_kernel<GG, u64><<<(unsigned)blocks, QK_THREADS, 0, s>>>(p->D, p->lut, p->rowptr, ngen, (u64*)ci, list)
    if (blocks > 0) {
      if (G == 32) { PDB_COLIDX(32); } else if (G == 16) { PDB_COLIDX(16); } else { PDB_COLIDX(8); }
    }
#undef PDB_COLIDX
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

template <int G, int DIM, int K, bool NT1, bool OUTFLOW, bool PW = false>
static int launch_qk_assemble_t(MatrixPlan* p, double* v, bool fresh, cudaStream_t s) {
  constexpr int N1 = K + 1, N = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  constexpr int MAXLEN = DIM == 3 ? (2 * K + 1) * (2 * K + 1) * (2 * K + 1) : (2 * K + 1) * (2 * K + 1);
  constexpr int ROWS_PER_CTA = QK_THREADS / G;
  constexpr int NCAND = DIM == 3 ? 8 : 4;
  const int nt = PW ? 0 : (NT1 ? 1 : p->Q.nt);
  const size_t smem = ((size_t)nt * N * N + (size_t)ROWS_PER_CTA * MAXLEN + (NT1 ? 0 : (size_t)ROWS_PER_CTA * NCAND * nt)) * sizeof(double) + 1000;
  auto kern = qk_assemble_kernel<G, DIM, K, NT1, OUTFLOW, PW>;
  PDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148, per_sm = 1;
  PDB_CUDA(cudaGetDevice(&dev));
  PDB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QK_THREADS, smem));
  // a single unit table and no boundary terms: interior rows go through the x-line kernel
  int launches = 0;
  const bool split = NT1 && !OUTFLOW && p->interior_ok;
  if (split) launches += launch_qk_interior<DIM, K, uint32_t, true>(p, v, nullptr, fresh, s);
  const u64 nrows = split ? p->nbr : p->nrows;
  const uint32_t* list = split ? p->brows : nullptr;
  const u64 want = (nrows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  const unsigned blocks = (unsigned)std::min<u64>(want, (u64)sms * std::max(per_sm, 1));
  if (blocks > 0) {
    kern<<<blocks, QK_THREADS, smem, s>>>(p->P, p->D, p->T, p->Q, p->tables, p->lut, p->rowptr, nrows, p->flags, v,
                                          fresh ? 1 : 0, list);
    launches++;
  }
  return launches;
}

template <int G, int DIM, int K>
static int launch_qk_assemble_nt(MatrixPlan* p, double* v, bool fresh, cudaStream_t s) {
  const bool outflow = p->P.bctype != nullptr && p->P.b != nullptr;  // jacobian_boundary has outflow terms only
  if (p->P.pw)  // coefficients sampled per quadrature point
    return outflow ? launch_qk_assemble_t<G, DIM, K, false, true, true>(p, v, fresh, s)
                   : launch_qk_assemble_t<G, DIM, K, false, false, true>(p, v, fresh, s);
  if (outflow) return launch_qk_assemble_t<G, DIM, K, false, true>(p, v, fresh, s);
  if (p->Q.nt == 1) return launch_qk_assemble_t<G, DIM, K, true, false>(p, v, fresh, s);
  return launch_qk_assemble_t<G, DIM, K, false, false>(p, v, fresh, s);
}

static int launch_qk_assemble(MatrixPlan* p, double* v, bool fresh, cudaStream_t s) {
  const DevParams& P = p->P;
  if (P.dim == 2 && P.k == 1) return launch_qk_assemble_nt<8, 2, 1>(p, v, fresh, s);
  if (P.dim == 2 && P.k == 2) return launch_qk_assemble_nt<16, 2, 2>(p, v, fresh, s);
  if (P.dim == 3 && P.k == 1) return launch_qk_assemble_nt<8, 3, 1>(p, v, fresh, s);
  if (P.dim == 3 && P.k == 2) return launch_qk_assemble_nt<32, 3, 2>(p, v, fresh, s);
  throw Error("conforming Jacobian: unsupported (dim, degree)");
}

int matrix_assemble(MatrixPlan* p, int layout, double* values, bool values_dev, bool fresh, int* errflag,
                    cudaStream_t s) {
  const DevParams& P = p->P;
  double* v = values;
  if (!values_dev) {
    PDB_CUDA(cudaMalloc(&v, std::max<u64>(p->nnz, 1) * sizeof(double)));
    if (!fresh) PDB_CUDA(cudaMemcpyAsync(v, values, p->nnz * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  int launches = 1;
  if (P.dg) {
    dg_assemble_kernel<<<(unsigned)P.ncells, 128, 0, s>>>(P, p->rowptr, layout == PDB200_LAYOUT_BCSR, v, fresh ? 1 : 0, errflag);
  } else {
    launches = launch_qk_assemble(p, v, fresh, s);
  }
  PDB_CUDA(cudaGetLastError());
  if (!values_dev) {
    PDB_CUDA(cudaMemcpyAsync(values, v, p->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    PDB_CUDA(cudaStreamSynchronize(s));
    cudaFree(v);
  }
  return launches;
}

int matrix_mv(MatrixPlan* p, int layout, const double* values, const double* x, double* y, cudaStream_t s) {
  const DevParams& P = p->P;
  int launches = 0;
  if (P.dg) {
    launches = 1;
    dg_mv_kernel<<<(unsigned)P.ncells, 128, 0, s>>>(P, p->rowptr, layout == PDB200_LAYOUT_BCSR, values, x, y);
  } else {
    const int G = P.n > 16 ? 32 : (P.n > 8 ? 16 : 8);
    // interior rows: one x-line per CTA; boundary rows (or everything on degenerate grids): generic kernel
    const uint32_t* list = p->interior_ok ? p->brows : nullptr;
    const u64 ngen = p->interior_ok ? p->nbr : p->nrows;
    if (p->interior_ok) {
      if (P.dim == 2 && P.k == 1) launches += launch_qk_mv_interior<2, 1>(p, values, x, y, s);
      else if (P.dim == 2 && P.k == 2) launches += launch_qk_mv_interior<2, 2>(p, values, x, y, s);
      else if (P.dim == 3 && P.k == 1) launches += launch_qk_mv_interior<3, 1>(p, values, x, y, s);
      else launches += launch_qk_mv_interior<3, 2>(p, values, x, y, s);
    }
    const u64 blocks = (ngen * G + QK_THREADS - 1) / QK_THREADS;
    if (blocks > 0) {
      if (G == 32) qk_mv_kernel<32><<<(unsigned)blocks, QK_THREADS, 0, s>>>(p->D, p->lut, p->rowptr, ngen, values, x, y, list);
      else if (G == 16) qk_mv_kernel<16><<<(unsigned)blocks, QK_THREADS, 0, s>>>(p->D, p->lut, p->rowptr, ngen, values, x, y, list);
      else qk_mv_kernel<8><<<(unsigned)blocks, QK_THREADS, 0, s>>>(p->D, p->lut, p->rowptr, ngen, values, x, y, list);
      launches++;
    }
  }
  PDB_CUDA(cudaGetLastError());
  return launches;
}

}  // namespace pdb

// dg_fast.cu — bandwidth-oriented jacobian_apply for the headline configuration:
//   QkDG k = 2, dim = 3, cell-wise constant DIAGONAL diffusion tensor, optional cell-wise constant velocity b and
//   reaction c, SIPG/NIPG/IIPG with or without harmonic weights, Dirichlet / Neumann / Outflow / None boundary faces.
//
// What it computes is exactly GridOperator::jacobian_apply for ConvectionDiffusionDG
// (gridoperator/gridoperator.hh:192-197 -> localoperator/convectiondiffusiondg.hh:106-188,
// 271-471, 684-879), but not with the reference's quadrature loops.  On an axis-aligned grid
// with cell-wise constant diagonal A every integrand is a product of 1-D polynomials of degree
// <= 2k, which the reference's (k+1)-point Gauss rule integrates exactly, so the operator equals
//     y_e = |K| (M (x) M (x) M) [ sum_d (1/h_d) M^-1 L_d(z_{e-d}, z_e, z_{e+d}) + c_e z_e ]
// with M the exact 1-D mass matrix and L_d the 1-D SIPG operator along direction d (volume
// stiffness + both face terms, coefficients from A_dd of the cell and its two d-neighbours).
// With a cell-wise constant velocity the convective terms are 1-D operators as well (HAS_B variants): the volume term
// -u b.grad psi (:178-187) becomes (b_d / h_d) [ u'(x_i) + (M^-1 e_0)_i u(0) - (M^-1 e_k)_i u(1) ] along direction d
// (integration by parts; u' is a polynomial of the space), and the upwind flux (:426-448) a selection between the two
// traces at either face, with the velocity of the larger-index cell.
// Derivation and the numpy statement of the same formula: DESIGN.md §5, tests/kron_reference.py.
// Results agree with the quadrature form to rounding (tests: <= 1e-12 relative to the oracle).
//
// Mapping to the machine (B200, sm_100a):
//   * CTA = 8x4x4 cells, one thread per cell (128 threads, 3 CTAs/SM).
//   * The cell tile plus its x/y halo (12 x 6 x 4 cells) is brought into shared memory by ONE TMA tensor copy
//     (cp.async.bulk.tensor.4d; out-of-domain cells are zero-filled by the TMA unit; the global tensor is viewed as
//     [54 doubles = 2 cells][Nx/2][Ny][Nz] so that all strides are multiples of 16 B), the two z-halo layers by
//     eight 1-D bulk copies (cp.async.bulk, one x-row of 8 cells each), all completing on one mbarrier.
//   * Shared memory is the busiest unit of this kernel (189 LDS.64 + 27 STS.64 per cell against 1083 fp64
//     instructions: the 128 B/clk crossbar and the fp64 pipe are equally loaded), so every access is laid out
//     bank-conflict free: a half-warp holds the rows cy and cy + 2 of one z-layer; rows of the main region are
//     12 cells = 648 words apart (two rows = 16 banks), and the z-halo rows and the output rows — moved by 1-D bulk
//     copies, which only need 16-byte alignment — put the rows cy >= 2 another 64 B (16 banks) further on.
//     (The first version used five TMA boxes and one dense output box: 31 % of its shared-memory wavefronts were
//     conflict replays — profiles/r01_v5_dg_fast_ncu_full_summary.json.)
//   * A thread keeps its 27 DOFs and 27 accumulators in registers; neighbour traces and normal
//     derivatives are read from the shared tile.
//   * A warp is one z-layer of the tile.  The z-sweep runs FIRST; after ONE block-wide barrier behind it a warp
//     only ever touches its own layer, so each warp stages its 4 result rows in its own (dead) layer and sends
//     them off with 4 bulk row stores (or reduce-adds) as soon as IT is done — no barrier, staging pass or store
//     loop of the whole block at the end (27 % of the warp samples in the first layout,
//     profiles/r02_v6_dg_fast_ncu_full_summary.json).  (Tried on top, A/B on one box against this version: the
//     barrier split into an mbarrier arrive behind the z-sweep and a wait behind the x/y sweeps, and the coefficient
//     loads issued before the barrier set-up — 1-2 % slower, not kept.)  Global traffic is fully coalesced: 8 B/DOF read (+ halo
//     re-reads served by L2) + 8 B/DOF written.

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace pdb {

namespace {

constexpr int TX = 8, TY = 4, TZ = 4;
constexpr int NLOC = 27;
constexpr int ROWX = TX + 4;                      // cells per x-row in smem: x0-2 .. x0+TX+1
constexpr int ROWY = TY + 2;                      // rows per z-layer: y0-1 .. y0+TY
constexpr int XROW = TX * NLOC;                   // doubles of one x-row of the tile (8 cells)
constexpr int BANKSHIFT = 8;                      // 8 doubles = 64 B = 16 banks
constexpr int R0 = 0;                             // rows region [TZ][ROWY][ROWX] cells (one TMA box)
constexpr int ROWS_BYTES = TZ * ROWY * ROWX * NLOC * 8;
constexpr int ZHALF = 2 * XROW + BANKSHIFT;       // z-halo: rows {0,1} | 64 B | rows {2,3}
constexpr int ZSIZE = 2 * ZHALF;                  // (the second pad keeps the next region 16-byte aligned and apart)
constexpr int R3 = R0 + TZ * ROWY * ROWX * NLOC;  // z-halo lower
constexpr int R4 = R3 + ZSIZE;                    // z-halo upper
constexpr int SMEM_DOUBLES = R4 + ZSIZE;
constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;      // 76,288 B: three CTAs per SM
constexpr int LAYER = ROWY * ROWX * NLOC;         // doubles of one z-layer of the rows region
// output stage of warp cz: the first ZSIZE doubles of its own layer (rows laid out like a z-halo layer), and behind
// it the rows of R(0) in the residual form
static_assert(2 * ZSIZE <= LAYER, "stage and R(0) rows must fit into one layer");
static_assert((R3 * 8) % 16 == 0 && (R4 * 8) % 16 == 0 && (XROW * 8) % 16 == 0 && (LAYER * 8) % 16 == 0 && (ZSIZE * 8) % 16 == 0,
              "bulk copies need 16-byte alignment");
__host__ __device__ constexpr int zhalo_row(int cy) { return (cy >> 1) * ZHALF + (cy & 1) * XROW; }

struct FastConst {
  // see Kron1D; all six vectors are pre-multiplied by |K| / 30^3 so that the three mass sweeps can
  // use the integer matrix 30 M = [[4,2,-1],[2,16,2],[-1,2,4]] of the quadratic Lagrange basis
  double E0[3], E1[3], m0[3], mk[3], q0[3], q1[3];
  double ih2[3];     // 1/h_d^2
  double ih[3];      // 1/h_d
  double alpha_pen;  // alpha * k (k + dim - 1)
  double theta, scale;  // scale = |K| / 27000
};

// Where the tiles sit in the local box.  Ghost layers in y and z (processor sides of an overlapping
// partition) are kept OUT of the tiling: the tile origin is shifted by the lower ghost layer and
// the active range ends before the upper one, so every rank tiles exactly its owned cells (same
// work as a single-GPU run of that size); ghost rows are zeroed by the tiles next to them.
// x is never shifted: the TMA view pairs cells along x (ghost cells in x are computed and zeroed
// through the `constrained` path instead).
struct TileFrame {
  int off[3];  // first tile of this launch
  int org[3];  // cell coordinate of tile (0,0,0)
  int lim[3];  // exclusive upper bound of the tiled cell range
  double* out; // the output vector (ghost rows are written with plain stores)
  int* err;    // device error flag: "Outflow boundary condition on inflow" (convectiondiffusiondg.hh:802-806)
  int pf;      // L2 prefetch distance in tiles of the launch's linear block order (0 = off)
  int accumulate;    // y += J x (TMA reduce-add store) instead of y = J x
  const double* r0;  // residual form: R(0) is added to the staged tile before it leaves
  // ---- one-launch step of the overlapping partition (fz != nullptr): 1-D grid =
  //      [push blocks][interior tiles][slabs of tiles that read a neighbour's layer from the mailbox]
  const FusedTable* fz;
  unsigned long long epoch;
  int npush, push_blocks;  // push blocks in front of the tiles, blocks per side
  int push_side[6];        // push block b works for side push_side[b / push_blocks]
  int tail_start, tail_z;  // the upper part of the interior box runs LAST (box nboxes - 1; same x/y extents as box 0)
  int nboxes;              // tile boxes in launch order: box 0 = lower part of the interior
  int box_start[8];        // first linear tile number of every box (box_start[nboxes] = number of tiles)
  int box[7][6];           // {offset[3], extent[3]} in tiles
  unsigned div_m[7][2];    // division by extent[0] and extent[1] as multiply-high + shifts (fast_div)
  unsigned char div_s[7][2][2];
  int side_tiles[6];       // tiles that read the receive buffer of a side (the last one sends the ack)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// warms L2 with the core box of a tile that a later CTA will load (no shared-memory destination)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
// (accumulate forms: the bulk reduce-add lets the L2 perform the read-modify-write; every element is touched once per
// launch, so the result does not depend on the order in which the rows arrive)
// 1-D bulk copies (16-byte aligned, size a multiple of 16): one x-row of cells
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ unsigned long long f_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void f_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long f_globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= want; a neighbour that never shows up ends in the error flag (p2p_check), not in a hung GPU
__device__ __forceinline__ void f_spin(const unsigned long long* flag, unsigned long long want, int* err) {
  if (f_ld_acquire_sys(flag) >= want) return;
  const unsigned long long t0 = f_globaltimer_ns();
  while (f_ld_acquire_sys(flag) < want) {
    if (f_globaltimer_ns() - t0 > 10ull * 1000 * 1000 * 1000) {
      atomicExch(err, 1);
      break;
    }
    __nanosleep(64);
  }
}

// n / d for any 32-bit n by multiply-high and two shifts (Granlund & Montgomery): every CTA decodes its tile number,
// and two hardware-emulated divisions cost ~50 issue slots per warp — 3 % of this kernel
__device__ __forceinline__ unsigned fast_div(unsigned n, unsigned m, unsigned s1, unsigned s2) {
  const unsigned t = __umulhi(m, n);
  return (t + ((n - t) >> s1)) >> s2;
}
// linear tile number of the fused launch -> tile coordinates; false behind the last tile
__device__ __forceinline__ bool fused_tile(const TileFrame& TF, int t, int& bx, int& by, int& bz) {
  if (t >= TF.box_start[TF.nboxes]) return false;
  int b = 0;
  while (t >= TF.box_start[b + 1]) b++;
  const unsigned u = (unsigned)(t - TF.box_start[b]);
  const unsigned q = fast_div(u, TF.div_m[b][0], TF.div_s[b][0][0], TF.div_s[b][0][1]);  // u / ex
  const unsigned r = fast_div(q, TF.div_m[b][1], TF.div_s[b][1][0], TF.div_s[b][1][1]);  // q / ey
  bx = TF.box[b][0] + (int)(u - q * (unsigned)TF.box[b][3]);
  by = TF.box[b][1] + (int)(q - r * (unsigned)TF.box[b][4]);
  bz = TF.box[b][2] + (int)r;
  return true;
}

// A push block of the fused launch: the owned boundary layer of x goes straight into the neighbour's receive buffer
// (remote 16-byte stores over NVLink), the last block of a side publishes ready = epoch.  Same protocol as
// p2p_push_kernel (halo.cu), which replaces the CopyDataHandle communication of boilerplate/pdelab.hh:872-880.
__device__ __noinline__ void fused_push(const TileFrame& TF, const double* __restrict__ xin, int b) {
  const int side = TF.push_side[b / TF.push_blocks], blk = b % TF.push_blocks;
  const FusedSide& S = TF.fz->s[side];
  if (threadIdx.x == 0) f_spin(S.my_ack, TF.epoch - 1, TF.fz->err);  // the neighbour has consumed the previous layer
  __syncthreads();
  const double2* __restrict__ src = (const double2*)xin;
  double2* __restrict__ dst = S.peer_buf;
  const long long total = S.total2, chunk = S.chunk2, stride = S.stride2, off = S.src_off2;
  const long long step = (long long)TF.push_blocks * blockDim.x;
  long long i = (long long)blk * blockDim.x + threadIdx.x;
  for (; i + 7 * step < total; i += 8 * step) {  // eight independent loads in flight per thread
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long j = i + u * step, c = j / chunk;
      v[u] = src[off + c * stride + (j - c * chunk)];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) dst[i + u * step] = v[u];
  }
  for (; i < total; i += step) {
    const long long c = i / chunk;
    dst[i] = src[off + c * stride + (i - c * chunk)];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&TF.fz->counters[side], 1u);
    if (done == (unsigned)TF.push_blocks - 1) {
      TF.fz->counters[side] = 0;
      __threadfence_system();
      f_st_release_sys(S.peer_ready, TF.epoch);
    }
  }
}

// The three places where a tile next to a processor side differs from any other tile; kept out of line so that the
// code of all other tiles stays what it is without the fused step.  pmask: bit (sd - 2) set = the tile reads the
// receive buffer of side sd (2, 3: lower / upper y; 4, 5: lower / upper z).
__device__ __noinline__ void fused_wait_sides(const TileFrame& TF, int pmask) {
#pragma unroll 1
  for (int sd = 2; sd < 6; sd++)
    if (pmask >> (sd - 2) & 1) f_spin(TF.fz->s[sd].my_ready, TF.epoch, TF.fz->err);
  asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies that follow read what remote stores wrote
}
// the y-halo row of a warp's layer is a neighbour's boundary layer: the TMA box brought x's (stale) ghost row; the warp
// overwrites the eight cells its y-sweep reads from the receive buffer ([Nz][Nx] cells).  L1 may hold the previous
// epoch's lines: ld.global.cg.
__device__ __noinline__ void fused_patch_y(const TileFrame& TF, double* tile, int pmask, int gz, int Nx, int x0, int ncx,
                                           int cz, int lane) {
  const int n2 = ncx * NLOC / 2;
#pragma unroll 1
  for (int sd = 0; sd < 2; sd++) {
    if (!(pmask >> sd & 1)) continue;
    const double2* __restrict__ src = (const double2*)(TF.fz->s[2 + sd].my_buf + ((long long)gz * Nx + x0) * NLOC);
    double2* __restrict__ dst = (double2*)(tile + R0 + ((cz * ROWY + (sd ? ROWY - 1 : 0)) * ROWX + 2) * NLOC);
    for (int i = lane; i < n2; i += 32) dst[i] = __ldcg(src + i);
  }
}
// every warp has its halo data in shared memory: this tile is done with the receive buffers; the last tile of a side
// hands the buffer back to the neighbour (ack = epoch)
__device__ __noinline__ void fused_ack(const TileFrame& TF, int pmask) {
  __threadfence();
#pragma unroll 1
  for (int sd = 2; sd < 6; sd++) {
    if (!(pmask >> (sd - 2) & 1)) continue;
    const unsigned int done = atomicAdd(&TF.fz->counters[6 + sd], 1u);
    if (done == (unsigned)TF.side_tiles[sd] - 1) {
      TF.fz->counters[6 + sd] = 0;
      __threadfence();
      f_st_release_sys(TF.fz->s[sd].peer_ack, TF.epoch);
    }
  }
}

// cell indices fit 32 bits on one GPU (512^3 = 2^27 cells); only the byte offset is 64-bit
template <int AMODE>
__device__ __forceinline__ double load_adiag(const DevParams& P, int cell, int d) {
  if (AMODE == PDB200_A_IDENTITY) return 1.0;
  if (AMODE == PDB200_A_SCALAR) return __ldg(P.A + cell);
  if (AMODE == PDB200_A_DIAGONAL) return __ldg(P.A + (long long)cell * 3 + d);
  return __ldg(P.A + (long long)cell * 9 + d * 4);
}

// One face of the 1-D operator along a direction: the three scalars
//   cs = w_self a / h^2, co = w_other a_other / h^2, cg = alpha pen harm / h^2 (penalty),
// following convectiondiffusiondg.hh:326-346 (interior) and :717-734 (Dirichlet boundary).
// kind: 0 interior, 1 Dirichlet boundary, 2 no u-dependent term (None/Neumann/Outflow with b=0,
// processor boundary).
// 1/x for positive normal x: MUFU.RCP64H seed (>= 20 bits) + one Newton step with a quadratic
// correction (error 2^-20 -> 2^-60, i.e. below the rounding of the products it enters).  Branch-free on
// purpose: the IEEE division's slow-path call serialises the six face set-ups.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);  // y (1 + e + e^2): cubic convergence
}

// Branch-free (selects only) so that the six reciprocal chains of a cell interleave.  The penalty
// coefficient is not returned: cg = alpha_pen (cs + co) with weights on (interior: 2 alpha_pen hh,
// Dirichlet: alpha_pen a / h^2), alpha_pen / h^2 on every interior / Dirichlet face with weights off (penalty_coef).
template <bool WEIGHTS_ON>
__device__ __forceinline__ void face_coef(int kind, double a, double ao, double ih2, double& cs, double& co) {
  const double aih = a * ih2;
  double csi, coi;  // interior face
  if (WEIGHTS_ON) {
    // w_self a = w_other a_other = a a_other / (a + a_other + 1e-20) = harmonic average / 2
    csi = coi = aih * ao * fast_rcp(a + ao + 1e-20);
  } else {
    csi = 0.5 * aih;
    coi = 0.5 * ao * ih2;
  }
  // Dirichlet boundary: w_self = 1; faces without u-dependent terms: -0.0 marks "no penalty" for penalty_coef
  cs = kind == 0 ? csi : (kind == 1 ? aih : (WEIGHTS_ON ? 0.0 : -0.0));
  co = kind == 0 ? coi : 0.0;
}
template <bool WEIGHTS_ON>
__device__ __forceinline__ double penalty_coef(double cs, double co, double ih2, double alpha_pen) {
  if (WEIGHTS_ON) return alpha_pen * (cs + co);
  return face_has_penalty(cs) ? alpha_pen * ih2 : 0.0;
}

// Adds (1/h_d) M^-1 L_d(l, o, r) for the nine lines of a cell along direction S-stride.
//   S = 1 (x), 3 (y), 9 (z): stride of the local node index along the direction.
//   t_i += P1_i u'_s(0) + P2_i u'_s(1) + P3_i u'_l(1) + P4_i u'_r(0) + P5_i [u]_L + P6_i [u]_R
// The own values are re-read from shared memory in every sweep instead of being kept in 54
// registers: the kernel is fp64-issue bound, not LDS bound, and the registers buy occupancy.
// convection along one direction (HAS_B): cb = b_d / h_d of the cell; cuLs / cuLo = (beta_L / h_d) on the own /
// the neighbour's trace at the lower face, whichever is upwind (convectiondiffusiondg.hh:426-448; boundary faces
// :797-822, 860), cuRs / cuRo the same at the upper face
struct Conv1D {
  double cb, cuLs, cuLo, cuRs, cuRo;
};

template <int S, bool FIRST, bool HAS_C, bool WEIGHTS_ON, bool HAS_B>
__device__ __forceinline__ void sweep(const double* __restrict__ no, double (&t)[NLOC], const double* __restrict__ nl,
                                      const double* __restrict__ nr, const FastConst& F, double creact, double A0,
                                      double ih2, double csL, double coL, double csR, double coR, const Conv1D& V) {
  double P1[3], P2[3], P3[3], P4[3], P5[3], P6[3];
  double P5o[3], P6o[3];  // HAS_B: the jumps split into their two traces (the upwind flux weights them differently)
  const double ctL = -F.theta * csL, ctR = F.theta * csR;
  const double cgL = penalty_coef<WEIGHTS_ON>(csL, coL, ih2, F.alpha_pen);
  const double cgR = penalty_coef<WEIGHTS_ON>(csR, coR, ih2, F.alpha_pen);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    P1[i] = fma(F.E0[i], A0, F.m0[i] * csL);
    P2[i] = fma(F.E1[i], A0, -F.mk[i] * csR);
    P3[i] = F.m0[i] * coL;
    P4[i] = -F.mk[i] * coR;
    P5[i] = fma(F.m0[i], cgL, F.q0[i] * ctL);
    P6[i] = fma(F.mk[i], cgR, F.q1[i] * ctR);
    if (HAS_B) {
      P5o[i] = fma(F.m0[i], V.cuLo, -P5[i]);          // on the left neighbour's trace l2
      P6o[i] = fma(F.mk[i], V.cuRo, -P6[i]);          // on the right neighbour's trace r0
      P5[i] = fma(F.m0[i], V.cuLs + V.cb, P5[i]);     // on the own trace o0 (upwind flux + boundary part of the volume term)
      P6[i] = fma(F.mk[i], V.cuRs - V.cb, P6[i]);     // on the own trace o2
    }
  }
  const double sb = HAS_B ? V.cb * F.scale : 0.0;     // b_d / h_d u'(x_i): u'(0) = dls, u'(1/2) = o2 - o0, u'(1) = drs
  if (HAS_B) {
    P1[0] += sb;
    P2[2] += sb;
  }
  constexpr int SA = S == 1 ? 3 : 1;  // strides of the two tangential node indices
  constexpr int SB = S == 9 ? 3 : 9;
#pragma unroll
  for (int b = 0; b < 3; b++)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int base = a * SA + b * SB;
      const double o0 = no[base], o1 = no[base + S], o2 = no[base + 2 * S];
      const double l0 = nl[base], l1 = nl[base + S], l2 = nl[base + 2 * S];
      const double r0 = nr[base], r1 = nr[base + S], r2 = nr[base + 2 * S];
      // p'(0) = (-3, 4, -1), p'(1) = (1, -4, 3) for the quadratic Lagrange basis on {0, 1/2, 1}
      const double dls = fma(4.0, o1, -fma(3.0, o0, o2));
      const double drs = fma(-4.0, o1, fma(3.0, o2, o0));
      const double dlo = fma(-4.0, l1, fma(3.0, l2, l0));
      const double dro = fma(4.0, r1, -fma(3.0, r0, r2));
      const double jl = o0 - l2, jr = o2 - r0;
      const double ov[3] = {o0, o1, o2};
#pragma unroll
      for (int i = 0; i < 3; i++) {
        double acc;
        if (FIRST)
          acc = HAS_C ? fma(P1[i], dls, creact * ov[i]) : P1[i] * dls;
        else
          acc = fma(P1[i], dls, t[base + i * S]);
        acc = fma(P2[i], drs, acc);
        acc = fma(P3[i], dlo, acc);
        acc = fma(P4[i], dro, acc);
        if (HAS_B) {
          acc = fma(P5[i], o0, acc);
          acc = fma(P5o[i], l2, acc);
          acc = fma(P6[i], o2, acc);
          acc = fma(P6o[i], r0, acc);
          if (i == 1) acc = fma(sb, o2 - o0, acc);
        } else {
          acc = fma(P5[i], jl, acc);
          acc = fma(P6[i], jr, acc);
        }
        t[base + i * S] = acc;
      }
    }
}

// v <- (30 M along stride S) v with 30 M = [[4,2,-1],[2,16,2],[-1,2,4]]; the factor |K|/30^3 is
// already inside t (FastConst)
template <int S>
__device__ __forceinline__ void mass_sweep(double (&t)[NLOC]) {
  constexpr int SA = S == 1 ? 3 : 1;
  constexpr int SB = S == 9 ? 3 : 9;
#pragma unroll
  for (int b = 0; b < 3; b++)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int base = a * SA + b * SB;
      const double v0 = t[base], v1 = t[base + S], v2 = t[base + 2 * S];
      const double e = v0 + v2;
      const double w = fma(2.0, v1, -e);  // 4 v0 + 2 v1 - v2 = 5 v0 + (2 v1 - v0 - v2)
      t[base] = fma(5.0, v0, w);
      t[base + S] = fma(16.0, v1, e + e);
      t[base + 2 * S] = fma(5.0, v2, w);
    }
}

// One tile of 8 x 4 x 4 cells.  LIST: the tile comes from the tile list of the one-launch overlapping step (1-D grid),
// BND: it is next to a processor side and reads the neighbour's boundary layer from the mailbox — bits of pmask: lower y,
// upper y, lower z, upper z.  The code of every other tile is the same with and without the fused step.
template <int AMODE, bool HAS_C, bool WEIGHTS_ON, bool HAS_B, bool LIST, bool BND>
__device__ __forceinline__ void tile_body(const CUtensorMap& tm_rows, const CUtensorMap& tm_pf, const double* __restrict__ xin,
                                          const DevParams& P, const FastConst& F, const TileFrame& TF, const int bx,
                                          const int by, const int bz, const int pmask) {
  extern __shared__ __align__(128) double tile[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t wbar[TZ];  // residual form: the R(0) rows of every warp arrive on its own barrier
  const int tid = threadIdx.x;
  const int x0 = bx * TX, y0 = TF.org[1] + by * TY, z0 = TF.org[2] + bz * TZ;

  if (tid == 0) {
    mbar_init(&bar, 1);
    if (TF.r0)
      for (int w = 0; w < TZ; w++) mbar_init(&wbar[w], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int Nx = P.N[0], Ny = P.N[1], Nz = P.N[2];
  const int ncx = min(TX, Nx - x0);                   // cells of an x-row inside the vector (even: Nx is even)
  const int nrow = min(TY, Ny - y0);                  // rows of the tile inside the vector
  const bool zlo_in = z0 - 1 >= 0, zup_in = z0 + TZ < Nz;
  if (tid == 0) {
    const uint32_t rowbytes = (uint32_t)ncx * NLOC * 8;
    mbar_expect_tx(&bar, ROWS_BYTES + ((zlo_in ? nrow : 0) + (zup_in ? nrow : 0)) * rowbytes);
    tma_load_4d(tile + R0, &tm_rows, 0, x0 / 2 - 1, y0 - 1, z0, &bar);
    if (BND) fused_wait_sides(TF, pmask);  // the neighbours' layers must have arrived (their push blocks run first)
    // z-halo layers: one bulk copy per x-row; a layer outside the vector is zeroed by the threads that read it.  Next
    // to a processor side the layer is read from the receive buffer ([Ny][Nx] cells) instead of from x's ghost layer.
    const double* zlo_src = BND && (pmask & 4) ? TF.fz->s[4].my_buf : xin + (long long)(z0 - 1) * Ny * Nx * NLOC;
    const double* zup_src = BND && (pmask & 8) ? TF.fz->s[5].my_buf : xin + (long long)(z0 + TZ) * Ny * Nx * NLOC;
    for (int r = 0; r < nrow; r++) {
      if (zlo_in) bulk_load(tile + R3 + zhalo_row(r), zlo_src + ((long long)(y0 + r) * Nx + x0) * NLOC, rowbytes, &bar);
      if (zup_in) bulk_load(tile + R4 + zhalo_row(r), zup_src + ((long long)(y0 + r) * Nx + x0) * NLOC, rowbytes, &bar);
    }
  }
  if (TF.pf && tid < ROWY * (TZ + 2) + 1) {
    // L2 prefetch for the tile TF.pf places further on in launch order: thread 0 the core box of the vector (every
    // input byte is prefetched once), threads 1.. one x-row each of the coefficient box that tile's cells will load
    int pbx, pby, pbz;
    bool pvalid;
    if (LIST) {
      pvalid = fused_tile(TF, (int)blockIdx.x - TF.npush + TF.pf, pbx, pby, pbz);
    } else {
      unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) + TF.pf;
      pbx = lin % gridDim.x + TF.off[0];
      lin /= gridDim.x;
      pby = lin % gridDim.y + TF.off[1];
      pbz = lin / gridDim.y + TF.off[2];
      pvalid = lin / gridDim.y < gridDim.z;
    }
    if (pvalid) {
      const int qx = pbx * TX, qy = TF.org[1] + pby * TY, qz = TF.org[2] + pbz * TZ;
      if (tid == 0) {
        tma_prefetch_4d(&tm_pf, 0, qx / 2, qy, qz);
      } else if (AMODE != PDB200_A_IDENTITY) {
        const int r = tid - 1, ky = qy - 1 + r % ROWY, kz = qz - 1 + r / ROWY;
        if (ky >= 0 && ky < Ny && kz >= 0 && kz < Nz && qx < Nx) {
          const long long c0 = (long long)max(qx - 1, 0) + (long long)Nx * (ky + (long long)Ny * kz);
          const int per = AMODE == PDB200_A_SCALAR ? 1 : (AMODE == PDB200_A_DIAGONAL ? 3 : 9);
          const double* a = P.A + c0 * per;
          prefetch_l2(a);                                   // 10 cells: at most two (scalar) .. six 128-byte lines
          for (int o = 16; o < 10 * per; o += 16) prefetch_l2(a + o);
          prefetch_l2(a + 10 * per - 1);
        }
      }
    }
  }

  // lane -> cell: half-warps cover rows {0,2} / {1,3} of a z-layer so that the 54-word cell
  // stride maps the 16 lanes of a 64-bit shared access onto 16 distinct bank pairs
  const int lane = tid & 31;
  const int cx = lane & 7;
  const int cy = ((lane >> 3) & 1) * 2 + (lane >> 4);
  const int cz = tid >> 5;
  const int gx = x0 + cx, gy = y0 + cy, gz = z0 + cz;
  const bool active = gx < Nx && gy < TF.lim[1] && gz < TF.lim[2];

  // ---- per-cell coefficients (overlaps the TMA latency).  All coefficient loads are issued
  // before the first use so that the L2 round trips overlap instead of adding up; everything up
  // to the reciprocals is branch-free so that the six face set-ups interleave. -------------------
  double A0[3], csL[3], coL[3], csR[3], coR[3];
  double creact = 0.0;
  bool constrained = false;
  double bs[3] = {0.0, 0.0, 0.0}, bo[3] = {0.0, 0.0, 0.0};  // HAS_B: own velocity and b_d of the upper d-neighbour
  int kinds = 0;                                             // HAS_B: face kinds, 2 bits each, face 2 d + side
  if (active) {
    const int cell = gx + Nx * (gy + Ny * gz);
    const int stride[3] = {1, Nx, Nx * Ny};
    const bool onb[3][2] = {{gx == 0, gx == Nx - 1}, {gy == 0, gy == Ny - 1}, {gz == 0, gz == Nz - 1}};
    double a[3], ao[3][2];
    int kind[3][2];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      a[d] = load_adiag<AMODE>(P, cell, d);
#pragma unroll
      for (int side = 0; side < 2; side++) {
        kind[d][side] = onb[d][side] ? 1 : 0;
        ao[d][side] = load_adiag<AMODE>(P, onb[d][side] ? cell : cell + (side ? stride[d] : -stride[d]), d);
      }
    }
    if (HAS_C) creact = __ldg(P.c + cell) * F.scale;
    if (HAS_B) {
#pragma unroll
      for (int d = 0; d < 3; d++) {
        bs[d] = __ldg(P.b + (long long)cell * 3 + d);
        bo[d] = onb[d][1] ? bs[d] : __ldg(P.b + (long long)(cell + stride[d]) * 3 + d);
      }
    }
    const bool any_b = onb[0][0] | onb[0][1] | onb[1][0] | onb[1][1] | onb[2][0] | onb[2][1];
    if (any_b) {  // rare: cells on the box surface
      // boundary-face numbers of pdelab_b200.h: tangential coordinates lexicographic, lower direction fastest
      const long long bf[3] = {gy + (long long)Ny * gz, gx + (long long)Nx * gz, gx + (long long)Nx * gy};
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int side = 0; side < 2; side++)
          if (onb[d][side]) {
            if (P.side_kind[d][side] == PDB200_SIDE_PROCESSOR) {
              kind[d][side] = 2;
              constrained = true;
            } else if (P.bctype) {
              const int bt = P.bctype[P.bf_off[d][side] + bf[d]];
              kind[d][side] = bt == PDB200_BC_DIRICHLET ? 1 : (HAS_B && bt == PDB200_BC_OUTFLOW ? 3 : 2);
            }
          }
    }
    if (HAS_B) {
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int side = 0; side < 2; side++) kinds |= kind[d][side] << (2 * (2 * d + side));
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
      A0[d] = a[d] * F.ih2[d];
      face_coef<WEIGHTS_ON>(kind[d][0] == 3 ? 2 : kind[d][0], a[d], ao[d][0], F.ih2[d], csL[d], coL[d]);
      face_coef<WEIGHTS_ON>(kind[d][1] == 3 ? 2 : kind[d][1], a[d], ao[d][1], F.ih2[d], csR[d], coR[d]);
    }
  }

  // ---- shared-memory addresses of the cell and its six face neighbours -------------------------
  const int so = R0 + ((cz * ROWY + cy + 1) * ROWX + cx + 2) * NLOC;
  const int xl = so - NLOC, xr = so + NLOC;
  const int yl = so - ROWX * NLOC, yr = so + ROWX * NLOC;
  const int zl = cz > 0 ? so - ROWY * ROWX * NLOC : R3 + zhalo_row(cy) + cx * NLOC;
  const int zr = cz < TZ - 1 ? so + ROWY * ROWX * NLOC : R4 + zhalo_row(cy) + cx * NLOC;
  // a z-halo layer outside the vector: every thread zeroes the slot only it reads (the TMA box does this by itself)
  if (active && ((cz == 0 && !zlo_in) || (cz == TZ - 1 && !zup_in))) {
    double* __restrict__ slot = tile + (cz == 0 && !zlo_in ? zl : zr);
#pragma unroll
    for (int i = 0; i < NLOC; i++) slot[i] = 0.0;
  }

  mbar_wait(&bar, 0);
  if (BND && (pmask & 3)) {  // the y-halo row of this warp's layer comes from the receive buffer (read by its y-sweep only)
    if (gz < Nz) fused_patch_y(TF, tile, pmask, gz, Nx, x0, ncx, cz, lane);
    __syncwarp();
  }

  double t[NLOC];
  // convection coefficients of direction d from the velocities and the face kinds
  auto conv = [&](int d) {
    Conv1D V = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (HAS_B) {
      const double ih = F.ih[d];
      V.cb = bs[d] * ih;
      const int kL = (kinds >> (4 * d)) & 3, kR = (kinds >> (4 * d + 2)) & 3;
      // lower face: this cell is the larger-index (inside) cell, n = -e_d, beta = -b_d (:426-438)
      const double betaL = -bs[d];
      const bool selfL = betaL >= 0.0;
      V.cuLs = (kL == 3 || (kL != 2 && selfL)) ? betaL * ih : 0.0;
      V.cuLo = (kL == 0 && !selfL) ? betaL * ih : 0.0;
      // upper face: interior -> the neighbour is the inside cell, its normal is -e_d and its velocity counts
      const double betaR = bo[d];
      const bool selfR = kR == 0 ? !(-betaR >= 0.0) : betaR >= 0.0;
      V.cuRs = (kR == 3 || (kR != 2 && selfR)) ? betaR * ih : 0.0;
      V.cuRo = (kR == 0 && !selfR) ? betaR * ih : 0.0;
      if ((kL == 3 && betaL < -1e-30) || (kR == 3 && betaR < -1e-30)) *TF.err = 1;  // :802-806
    }
    return V;
  };
  // the z-sweep first: it is the only one that reads other warps' layers
  if (active)
    sweep<9, true, HAS_C, WEIGHTS_ON, HAS_B>(tile + so, t, tile + zl, tile + zr, F, creact, A0[2], F.ih2[2], csL[2], coL[2], csR[2],
                                             coR[2], conv(2));
  __syncthreads();  // from here on a warp reads and writes its own layer only
  if (BND && tid == 0) fused_ack(TF, pmask);  // all halo data is in shared memory or consumed: the receive buffers are free
  if (active) {
    sweep<1, false, HAS_C, WEIGHTS_ON, HAS_B>(tile + so, t, tile + xl, tile + xr, F, creact, A0[0], F.ih2[0], csL[0], coL[0], csR[0],
                                              coR[0], conv(0));
    sweep<3, false, HAS_C, WEIGHTS_ON, HAS_B>(tile + so, t, tile + yl, tile + yr, F, creact, A0[1], F.ih2[1], csL[1], coL[1], csR[1],
                                              coR[1], conv(1));
  }
  __syncwarp();  // every lane is done reading the layer: it becomes the warp's output stage
  double* const stage = tile + R0 + cz * LAYER;
  const uint32_t rowbytes = (uint32_t)ncx * NLOC * 8;
  const bool rowlane = lane < TY && y0 + lane < Ny && gz < Nz;  // lane r sends row r of the layer, if inside the vector
  if (TF.r0 && lane == 0) {
    // residual form R(x) = J x + R(0) (the operator is affine): the rows of the cached R(0) travel into the layer
    // behind the stage while the mass sweeps run; they are added to y by reduce-adds, no thread touches them
    const int nr = gz < Nz ? nrow : 0;
    mbar_expect_tx(&wbar[cz], nr * rowbytes);
    for (int r = 0; r < nr; r++)
      bulk_load(stage + ZSIZE + zhalo_row(r), TF.r0 + (((long long)gz * Ny + y0 + r) * Nx + x0) * NLOC, rowbytes, &wbar[cz]);
  }
  if (active) {
    mass_sweep<1>(t);
    mass_sweep<3>(t);
    mass_sweep<9>(t);
    if (constrained) {  // constraints/p0.hh:31-41 + constrain_residual (jacobianapplyengine.hh:249-254)
#pragma unroll
      for (int i = 0; i < NLOC; i++) t[i] = 0.0;
    }
  }
  {  // the row stores write the whole box (clipped to the vector): cells of the box outside the
     // tiled range are ghost rows, which are zero
    double* __restrict__ dst = stage + zhalo_row(cy) + cx * NLOC;
#pragma unroll
    for (int i = 0; i < NLOC; i++) dst[i] = active ? t[i] : 0.0;
    if (TF.accumulate && active && constrained) {  // constrained rows are SET to zero, not incremented
      double* __restrict__ row = TF.out + (long long)(gx + Nx * (gy + Ny * gz)) * NLOC;
#pragma unroll
      for (int i = 0; i < NLOC; i++) row[i] = 0.0;
    }
  }
  fence_proxy_async();
  __syncwarp();
  if (rowlane) {
    double* dst = TF.out + (((long long)gz * Ny + y0 + lane) * Nx + x0) * NLOC;
    if (TF.accumulate)
      bulk_reduce_add(dst, stage + zhalo_row(lane), rowbytes);
    else
      bulk_store(dst, stage + zhalo_row(lane), rowbytes);
    if (TF.r0) {
      mbar_wait(&wbar[cz], 0);
      bulk_reduce_add(dst, stage + ZSIZE + zhalo_row(lane), rowbytes);
    }
  }

  // ---- rows of the ghost layers in y / z next to this tile := 0 (constraints/p0.hh:31-41 +
  // constrain_residual); only tiles at the ends of the tiled range take this path ----------------
  const bool gzl = TF.org[2] && z0 == TF.org[2], gzu = TF.lim[2] < Nz && z0 + TZ >= TF.lim[2];
  const bool gyl = TF.org[1] && y0 == TF.org[1], gyu = TF.lim[1] < Ny && y0 + TY >= TF.lim[1];
  bool zissued = false;
  if ((gzl | gzu | gyl | gyu) && cz == 0) {
    // warp 0 zeroes one x-row in the (dead) lower z-halo region and sends it to every ghost row next to the tile as a
    // bulk store.  (A loop of plain stores over these rows cost 5-6 % of the whole launch on a 128 x 130 x 130 box.)
    double* __restrict__ zsrc = tile + R3;
    for (int i = lane; i < XROW; i += 32) zsrc[i] = 0.0;
    fence_proxy_async();
    __syncwarp();
    const int ya = gyl ? 0 : y0, yb = gyu ? Ny : min(y0 + TY, TF.lim[1]);
    const int za = gzl ? 0 : z0, zb = gzu ? Nz : min(z0 + TZ, TF.lim[2]);
    const int ny = yb - ya, cnt = ny * (zb - za);
    for (int i = lane; i < cnt; i += 32) {
      const int zz = za + i / ny, yy = ya + i % ny;
      const bool ghost = (TF.org[1] && yy == 0) || (TF.lim[1] < Ny && yy == Ny - 1) || (TF.org[2] && zz == 0) ||
                         (TF.lim[2] < Nz && zz == Nz - 1);
      if (!ghost) continue;
      bulk_store(TF.out + (((long long)zz * Ny + yy) * Nx + x0) * NLOC, zsrc, rowbytes);
      zissued = true;
    }
  }
  if (rowlane || zissued) tma_store_commit_and_wait();
}

#ifndef PDB200_FAST_MINB
#define PDB200_FAST_MINB 3  // tuning aid (tools/build_variant.sh): resident CTAs per SM the register cap is set for
#endif
template <int AMODE, bool HAS_C, bool WEIGHTS_ON, bool HAS_B, bool FUSED>
__global__ void __launch_bounds__(TX* TY* TZ, PDB200_FAST_MINB)
    dg_fast_q2_3d_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_pf,
                         const double* __restrict__ xin, const __grid_constant__ DevParams P,
                         const __grid_constant__ FastConst F, const __grid_constant__ TileFrame TF) {
  if (FUSED) {  // one-launch step of the overlapping partition: 1-D grid = push blocks, then the tiles in list order
    if ((int)blockIdx.x < TF.npush) {
      fused_push(TF, xin, blockIdx.x);
      return;
    }
    const unsigned t = blockIdx.x - TF.npush;
    const bool head = t < (unsigned)TF.box_start[1], tail = t >= (unsigned)TF.tail_start;
    if (head | tail) {
      // an interior tile: the lower part of the interior box runs first, the upper part last (the tiles next to a
      // processor side sit in between: late enough for the neighbours' layers to have arrived, and not in the tail of
      // the launch).  Decoded from scalars of the constant bank only, so that the tile coordinates stay in uniform
      // registers like blockIdx of the 3-D launch: the tile body runs at the 168-register cap, and per-thread
      // coordinates cost it 9 % (A/B on one box).
      const unsigned u = head ? t : t - (unsigned)TF.tail_start;
      const unsigned q = fast_div(u, TF.div_m[0][0], TF.div_s[0][0][0], TF.div_s[0][0][1]);
      const unsigned r = fast_div(q, TF.div_m[0][1], TF.div_s[0][1][0], TF.div_s[0][1][1]);
      tile_body<AMODE, HAS_C, WEIGHTS_ON, HAS_B, true, false>(tm_rows, tm_pf, xin, P, F, TF, TF.box[0][0] + (int)(u - q * (unsigned)TF.box[0][3]),
                                                              TF.box[0][1] + (int)(q - r * (unsigned)TF.box[0][4]),
                                                              (head ? TF.box[0][2] : TF.tail_z) + (int)r, 0);
    } else {  // a tile next to a processor side
      int bx, by, bz;
      fused_tile(TF, (int)t, bx, by, bz);
      const int y0 = TF.org[1] + by * TY, z0 = TF.org[2] + bz * TZ;
      const int pmask = (y0 == 1 && P.side_kind[1][0] == PDB200_SIDE_PROCESSOR ? 1 : 0) |
                        (y0 + TY == P.N[1] - 1 && P.side_kind[1][1] == PDB200_SIDE_PROCESSOR ? 2 : 0) |
                        (z0 == 1 && P.side_kind[2][0] == PDB200_SIDE_PROCESSOR ? 4 : 0) |
                        (z0 + TZ == P.N[2] - 1 && P.side_kind[2][1] == PDB200_SIDE_PROCESSOR ? 8 : 0);
      tile_body<AMODE, HAS_C, WEIGHTS_ON, HAS_B, true, true>(tm_rows, tm_pf, xin, P, F, TF, bx, by, bz, pmask);
    }
  } else {
    tile_body<AMODE, HAS_C, WEIGHTS_ON, HAS_B, false, false>(tm_rows, tm_pf, xin, P, F, TF, blockIdx.x + TF.off[0],
                                                             blockIdx.y + TF.off[1], blockIdx.z + TF.off[2], 0);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct FastPlan {
  FastConst F;
  EncodeFn encode = nullptr;
  struct Maps {
    const void* ptr = nullptr;
    CUtensorMap rows, core;  // rows: the input box of a tile; core: its cells without halo (L2 prefetch)
  };
  std::vector<Maps> cache;  // tensor maps embed the global address: keep the most recent few
  double* scratch = nullptr;
  long long scratch_n = 0;
};

bool dg_fast_supported(const DevParams& P) {
  // a cell-wise constant velocity is a Kronecker term too (HAS_B variants)
  return P.dg && P.basis == PDB200_BASIS_LAGRANGE && P.dim == 3 && P.k == 2 && P.m >= 3 && P.pw == 0 && P.a_mode != PDB200_A_FULL &&
         P.N[0] % 2 == 0;
}

FastPlan* dg_fast_plan_create(const DevParams& P, const Kron1D& K) {
  FastPlan* plan = new FastPlan;
  FastConst& F = plan->F;
  F.scale = P.vol / 27000.0;
  for (int i = 0; i < 3; i++) {
    F.E0[i] = K.E0[i] * F.scale;
    F.E1[i] = K.E1[i] * F.scale;
    F.m0[i] = K.m0[i] * F.scale;
    F.mk[i] = K.mk[i] * F.scale;
    F.q0[i] = K.q0[i] * F.scale;
    F.q1[i] = K.q1[i] * F.scale;
    F.ih2[i] = 1.0 / (P.h[i] * P.h[i]);
    F.ih[i] = 1.0 / P.h[i];
  }
  F.alpha_pen = P.alpha * P.k * (P.k + P.dim - 1);
  F.theta = P.theta;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PDB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) throw Error("cuTensorMapEncodeTiled is not available in this driver");
  plan->encode = (EncodeFn)fn;
#define PDB_SET_SMEM(AM, HC, WO)                                                                                 \
  PDB_CUDA(cudaFuncSetAttribute(dg_fast_q2_3d_kernel<AM, HC, WO, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                SMEM_BYTES));                                                                     \
  PDB_CUDA(cudaFuncSetAttribute(dg_fast_q2_3d_kernel<AM, HC, WO, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                SMEM_BYTES));                                                                     \
  PDB_CUDA(cudaFuncSetAttribute(dg_fast_q2_3d_kernel<AM, HC, WO, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                SMEM_BYTES));                                                                     \
  PDB_CUDA(cudaFuncSetAttribute(dg_fast_q2_3d_kernel<AM, HC, WO, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                SMEM_BYTES));
#define PDB_SET_SMEM4(AM) PDB_SET_SMEM(AM, false, false) PDB_SET_SMEM(AM, false, true) PDB_SET_SMEM(AM, true, false) PDB_SET_SMEM(AM, true, true)
  PDB_SET_SMEM4(PDB200_A_IDENTITY) PDB_SET_SMEM4(PDB200_A_SCALAR) PDB_SET_SMEM4(PDB200_A_DIAGONAL)
#undef PDB_SET_SMEM4
#undef PDB_SET_SMEM
  return plan;
}

void dg_fast_plan_destroy(FastPlan* plan) {
  if (!plan) return;
  if (plan->scratch) cudaFree(plan->scratch);
  delete plan;
}

static void encode_map(FastPlan* plan, CUtensorMap* m, const void* ptr, const DevParams& P, int bx, int by, int bz) {
  cuuint64_t gdim[4] = {54, (cuuint64_t)P.N[0] / 2, (cuuint64_t)P.N[1], (cuuint64_t)P.N[2]};
  cuuint64_t gstr[3] = {432, 216ull * P.N[0], 216ull * P.N[0] * P.N[1]};
  cuuint32_t box[4] = {54, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = plan->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

static FastPlan::Maps& get_maps(FastPlan* plan, const void* ptr, const DevParams& P) {
  for (auto& m : plan->cache)
    if (m.ptr == ptr) return m;
  if ((uintptr_t)ptr % 16 != 0) throw Error("fast DG kernel: vectors must be 16-byte aligned");
  if (plan->cache.size() >= 16) plan->cache.erase(plan->cache.begin());
  FastPlan::Maps m;
  m.ptr = ptr;
  encode_map(plan, &m.rows, ptr, P, ROWX / 2, ROWY, TZ);
  encode_map(plan, &m.core, ptr, P, TX / 2, TY, TZ);
  plan->cache.push_back(m);
  return plan->cache.back();
}

// Tiles whose face halo reads a ghost cell layer (a side of kind PDB200_SIDE_PROCESSOR) are the
// "boundary" part; the rest, a box of tiles, is the "interior" part, which can run while the halo
// exchange is in flight.  Returns the interior tile box [lo, hi) per direction.
static void tile_frame(const DevParams& P, TileFrame& F, int nt[3]) {
  const int T[3] = {TX, TY, TZ};
  F = TileFrame{};  // fz = nullptr: not a fused launch
  for (int d = 0; d < 3; d++) {
    const bool glo = d > 0 && P.side_kind[d][0] == PDB200_SIDE_PROCESSOR;
    const bool ghi = d > 0 && P.side_kind[d][1] == PDB200_SIDE_PROCESSOR;
    F.off[d] = 0;
    F.org[d] = glo ? 1 : 0;
    F.lim[d] = P.N[d] - (ghi ? 1 : 0);
    nt[d] = std::max(0, (F.lim[d] - F.org[d] + T[d] - 1) / T[d]);
  }
}

static void interior_tile_box(const DevParams& P, const TileFrame& F, const int nt[3], int lo[3], int hi[3]) {
  const int T[3] = {TX, TY, TZ};
  for (int d = 0; d < 3; d++) {
    lo[d] = 0;
    hi[d] = nt[d];
    // tile t covers cells [org + t T, org + (t+1) T) and reads one more layer on either side
    if (P.side_kind[d][0] == PDB200_SIDE_PROCESSOR) lo[d] = 1;  // tile 0 reads the ghost layer 0
    if (P.side_kind[d][1] == PDB200_SIDE_PROCESSOR) {
      int t = 0;
      while (t < nt[d] && F.org[d] + (t + 1) * T[d] < P.N[d] - 1) t++;  // first tile that reaches layer N-1
      hi[d] = t;
    }
    if (hi[d] < lo[d]) hi[d] = lo[d];
  }
}

// number of tile layers along z and the cell layers [z0, z1) that tile layers [lo, hi) write
int dg_fast_ztiles(const DevParams& P) {
  TileFrame TF;
  int nt[3];
  tile_frame(P, TF, nt);
  return nt[2];
}
void dg_fast_ztile_layers(const DevParams& P, int lo, int hi, int* z0, int* z1) {
  TileFrame TF;
  int nt[3];
  tile_frame(P, TF, nt);
  *z0 = lo <= 0 ? 0 : TF.org[2] + lo * TZ;                     // the first window also owns the lower ghost layer
  *z1 = hi >= nt[2] ? P.N[2] : std::min(TF.org[2] + hi * TZ, P.N[2]);  // the last one the upper ghost layer
}

int launch_dg_fast_fused(FastPlan* plan, const DevParams& P, const double* x, double* y, const FusedTable* table,
                         const FusedTable& th, unsigned long long epoch, cudaStream_t s, int* errflag);

int launch_dg_fast(FastPlan* plan, const DevParams& P, const double* x, double* y, const double* r0, bool overwrite,
                   int part, cudaStream_t s, int* errflag, int ztile_lo, int ztile_hi) {
  if (r0 && overwrite) throw Error("the residual form accumulates (r += J x + R(0))");
  if (part != PDB200_PART_ALL && !overwrite) throw Error("partial application needs the overwrite form");
  double* out = y;  // accumulate semantics (y += J z [+ R(0)]) through the TMA reduce-add store
  const FastPlan::Maps mx = get_maps(plan, x, P);
  if ((uintptr_t)out % 16 != 0 || (r0 && (uintptr_t)r0 % 16 != 0)) throw Error("fast DG kernel: vectors must be 16-byte aligned");
  TileFrame TF;
  int nt[3];
  tile_frame(P, TF, nt);
  TF.out = out;
  TF.err = errflag;
  TF.accumulate = overwrite ? 0 : 1;
  TF.r0 = r0;
  {
    static const int pf_env = [] {
      const char* e = getenv("PDB200_FAST_PREFETCH");
      // L2 prefetch distance in tiles.  It bought 1-2 % with the first shared-memory layout (444 = three tiles per SM
      // ahead, profiles/r01_prefetch_sweep.txt) and costs ~1 % with the present one (A/B on one box, round 2): off.
      return e ? atoi(e) : 0;
    }();
    TF.pf = pf_env;

  }
  // boxes of tiles to launch: {offset, extent}
  int boxes[7][6], nboxes = 0;
  auto add = [&](int ox, int oy, int oz, int ex, int ey, int ez) {
    if (ex <= 0 || ey <= 0 || ez <= 0) return;
    const int b[6] = {ox, oy, oz, ex, ey, ez};
    for (int i = 0; i < 6; i++) boxes[nboxes][i] = b[i];
    nboxes++;
  };
  {
    // tuning aid: the tile-list (1-D grid) variant of the kernel on a box without processor sides
    static const bool linear = [] { const char* e = getenv("PDB200_FAST_LINEAR"); return e && e[0] == '1'; }();
    if (linear && part == PDB200_PART_ALL && overwrite && !r0 && ztile_lo <= 0 && ztile_hi >= nt[2]) {
      static const FusedTable none{};
      return launch_dg_fast_fused(plan, P, x, y, nullptr, none, 0, s, errflag);
    }
  }
  if (part == PDB200_PART_ALL) {
    const int zlo = std::max(0, ztile_lo), zhi = std::min(nt[2], ztile_hi);  // optional window of tile layers
    add(0, 0, zlo, nt[0], nt[1], zhi - zlo);
  } else {
    int lo[3], hi[3];
    interior_tile_box(P, TF, nt, lo, hi);
    if (part == PDB200_PART_INTERIOR) {
      add(lo[0], lo[1], lo[2], hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]);
    } else {  // the complement as up to six slabs
      add(0, 0, 0, nt[0], nt[1], lo[2]);
      add(0, 0, hi[2], nt[0], nt[1], nt[2] - hi[2]);
      add(0, 0, lo[2], nt[0], lo[1], hi[2] - lo[2]);
      add(0, hi[1], lo[2], nt[0], nt[1] - hi[1], hi[2] - lo[2]);
      add(0, lo[1], lo[2], lo[0], hi[1] - lo[1], hi[2] - lo[2]);
      add(hi[0], lo[1], lo[2], nt[0] - hi[0], hi[1] - lo[1], hi[2] - lo[2]);
    }
  }
  int launches = 0;
#define PDB_LAUNCH(AM, HC, WO)                                                                                            \
  do {                                                                                                                     \
    if (P.b) dg_fast_q2_3d_kernel<AM, HC, WO, true, false><<<grid, TX * TY * TZ, SMEM_BYTES, s>>>(mx.rows, mx.core, x, P, plan->F, TF); \
    else dg_fast_q2_3d_kernel<AM, HC, WO, false, false><<<grid, TX * TY * TZ, SMEM_BYTES, s>>>(mx.rows, mx.core, x, P, plan->F, TF);  \
  } while (0)
#define PDB_LAUNCH_A(AM)                                \
  do {                                                  \
    if (P.c && P.weights_on) PDB_LAUNCH(AM, true, true);        \
    else if (P.c) PDB_LAUNCH(AM, true, false);          \
    else if (P.weights_on) PDB_LAUNCH(AM, false, true); \
    else PDB_LAUNCH(AM, false, false);                  \
  } while (0)
  const int am = P.a_mode == PDB200_A_IDENTITY || P.a_mode == PDB200_A_SCALAR ? P.a_mode : PDB200_A_DIAGONAL;
  for (int b = 0; b < nboxes; b++) {
    for (int d = 0; d < 3; d++) TF.off[d] = boxes[b][d];
    const dim3 grid(boxes[b][3], boxes[b][4], boxes[b][5]);
    if (am == PDB200_A_IDENTITY) PDB_LAUNCH_A(PDB200_A_IDENTITY);
    else if (am == PDB200_A_SCALAR) PDB_LAUNCH_A(PDB200_A_SCALAR);
    else PDB_LAUNCH_A(PDB200_A_DIAGONAL);
    launches++;
  }
#undef PDB_LAUNCH_A
#undef PDB_LAUNCH
  PDB_CUDA(cudaGetLastError());
  return launches;
}

// The fused step needs every ghost layer exactly one cell outside a tile (never inside a tile box): the owned extent of
// a split direction is a multiple of the tile extent; x is never split.
bool dg_fast_fused_supported(const DevParams& P) {
  if (!dg_fast_supported(P)) return false;
  if (P.side_kind[0][0] == PDB200_SIDE_PROCESSOR || P.side_kind[0][1] == PDB200_SIDE_PROCESSOR) return false;
  const int T[3] = {TX, TY, TZ};
  bool any = false;
  for (int d = 1; d < 3; d++) {
    const bool lo = P.side_kind[d][0] == PDB200_SIDE_PROCESSOR, hi = P.side_kind[d][1] == PDB200_SIDE_PROCESSOR;
    any |= lo | hi;
    if (hi && (P.N[d] - 1 - (lo ? 1 : 0)) % T[d] != 0) return false;
    if ((lo || hi) && P.N[d] - (lo ? 1 : 0) - (hi ? 1 : 0) < T[d]) return false;
  }
  return any;
}

int launch_dg_fast_fused(FastPlan* plan, const DevParams& P, const double* x, double* y, const FusedTable* table,
                         const FusedTable& th, unsigned long long epoch, cudaStream_t s, int* errflag) {
  if (table && !dg_fast_fused_supported(P)) throw Error("fused overlapping step: unsupported partition");
  const FastPlan::Maps mx = get_maps(plan, x, P);
  if ((uintptr_t)y % 16 != 0) throw Error("fast DG kernel: vectors must be 16-byte aligned");
  TileFrame TF;
  int nt[3];
  tile_frame(P, TF, nt);
  TF.out = y;
  TF.err = errflag;
  TF.accumulate = 0;
  TF.r0 = nullptr;
  {
    const char* e = getenv("PDB200_FAST_PREFETCH");
    TF.pf = e ? atoi(e) : 0;
  }
  TF.fz = table;
  TF.epoch = epoch;
  {
    const char* e = getenv("PDB200_PUSH_BLOCKS");
    TF.push_blocks = e ? std::max(1, atoi(e)) : 32;
  }
  int nside = 0;
  for (int i = 0; i < 6; i++)
    if (th.s[i].active) TF.push_side[nside++] = i;
  TF.npush = nside * TF.push_blocks;
  int lo[3], hi[3];
  interior_tile_box(P, TF, nt, lo, hi);
  TF.nboxes = 0;
  int count = 0;
  auto add = [&](int ox, int oy, int oz, int ex, int ey, int ez) {
    if (TF.nboxes > 0 && (ex <= 0 || ey <= 0 || ez <= 0)) return;  // box 0 (interior) always exists, possibly empty
    const int b[6] = {ox, oy, oz, std::max(ex, 0), std::max(ey, 0), std::max(ez, 0)};
    for (int i = 0; i < 6; i++) TF.box[TF.nboxes][i] = b[i];
    TF.box_start[TF.nboxes] = count;
    count += b[3] * b[4] * b[5];
    TF.nboxes++;
  };
  // the interior box is cut along z: the lower part runs first, the upper part last
  const bool has_interior = hi[0] > lo[0] && hi[1] > lo[1] && hi[2] > lo[2];
  const int zcut = has_interior ? lo[2] + (hi[2] - lo[2] + 1) / 2 : lo[2];
  add(lo[0], lo[1], lo[2], hi[0] - lo[0], hi[1] - lo[1], zcut - lo[2]);
  if (!has_interior) {  // degenerate: the whole box is boundary
    TF.box[0][3] = TF.box[0][4] = TF.box[0][5] = 1;  // extents must not be zero for the decode; count stays 0
    add(0, 0, 0, nt[0], nt[1], nt[2]);
    TF.tail_start = count;
    TF.tail_z = 0;
  } else {
    add(0, 0, 0, nt[0], nt[1], lo[2]);
    add(0, 0, hi[2], nt[0], nt[1], nt[2] - hi[2]);
    add(0, 0, lo[2], nt[0], lo[1], hi[2] - lo[2]);
    add(0, hi[1], lo[2], nt[0], nt[1] - hi[1], hi[2] - lo[2]);
    TF.tail_start = count;
    TF.tail_z = zcut;
    add(lo[0], lo[1], zcut, hi[0] - lo[0], hi[1] - lo[1], hi[2] - zcut);  // (may be empty)
  }
  TF.box_start[TF.nboxes] = count;
  for (int b = 0; b < TF.nboxes; b++)
    for (int i = 0; i < 2; i++) {
      const unsigned d = (unsigned)std::max(1, TF.box[b][3 + i]);
      unsigned l = 0;
      while ((1ull << l) < d) l++;
      TF.div_m[b][i] = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
      TF.div_s[b][i][0] = (unsigned char)std::min(l, 1u);
      TF.div_s[b][i][1] = (unsigned char)(l > 0 ? l - 1 : 0);
    }
  if (count != nt[0] * nt[1] * nt[2]) throw Error("fused overlapping step: tile list does not cover the box");
  // tiles reading the receive buffer of a side: the first / last tile layer of that direction
  for (int i = 0; i < 6; i++) TF.side_tiles[i] = 0;
  for (int d = 1; d < 3; d++)
    for (int sd = 0; sd < 2; sd++)
      if (P.side_kind[d][sd] == PDB200_SIDE_PROCESSOR) TF.side_tiles[2 * d + sd] = nt[0] * nt[d == 1 ? 2 : 1];
  const unsigned grid = (unsigned)(TF.npush + count);
#define PDB_LAUNCH(AM, HC, WO)                                                                                            \
  do {                                                                                                                     \
    if (P.b) dg_fast_q2_3d_kernel<AM, HC, WO, true, true><<<grid, TX * TY * TZ, SMEM_BYTES, s>>>(mx.rows, mx.core, x, P, plan->F, TF); \
    else dg_fast_q2_3d_kernel<AM, HC, WO, false, true><<<grid, TX * TY * TZ, SMEM_BYTES, s>>>(mx.rows, mx.core, x, P, plan->F, TF);  \
  } while (0)
#define PDB_LAUNCH_A(AM)                                \
  do {                                                  \
    if (P.c && P.weights_on) PDB_LAUNCH(AM, true, true);        \
    else if (P.c) PDB_LAUNCH(AM, true, false);          \
    else if (P.weights_on) PDB_LAUNCH(AM, false, true); \
    else PDB_LAUNCH(AM, false, false);                  \
  } while (0)
  const int am = P.a_mode == PDB200_A_IDENTITY || P.a_mode == PDB200_A_SCALAR ? P.a_mode : PDB200_A_DIAGONAL;
  if (am == PDB200_A_IDENTITY) PDB_LAUNCH_A(PDB200_A_IDENTITY);
  else if (am == PDB200_A_SCALAR) PDB_LAUNCH_A(PDB200_A_SCALAR);
  else PDB_LAUNCH_A(PDB200_A_DIAGONAL);
#undef PDB_LAUNCH_A
#undef PDB_LAUNCH
  PDB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace pdb

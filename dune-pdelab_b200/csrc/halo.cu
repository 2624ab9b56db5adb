// halo.cu — pack / unpack of one cell layer of a QkDG vector for the overlapping partition.
//
// Replaces the per-entity gather/scatter of GFSDataHandle + CopyGatherScatter
// (gridfunctionspace/genericdatahandle.hh:130-260) used by the overlapping backends
// (boilerplate/pdelab.hh:872-880): with overlap 1 the data sent to the neighbour across side
// (dir, side) is the owned cell layer at distance 1 from that box face, and the data received
// fills the ghost layer at distance 0.  A DG cell is n contiguous doubles, so the copy is a
// strided block copy; consecutive threads move consecutive doubles of a cell.
#include <algorithm>
#include <cstddef>
#include <cstring>

#include "common.cuh"
#include "host_tables.h"

namespace pdb {
namespace {

__global__ void halo_copy_kernel(const DevParams P, double* __restrict__ x, double* __restrict__ buf, int dir,
                                 int layer, int pack, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const long long f = i / P.n;  // cell number inside the layer (lexicographic tangential order)
    const int k = (int)(i - f * P.n);
    int c[3] = {0, 0, 0};
    long long ff = f;
    for (int d = 0; d < P.dim; d++)
      if (d != dir) {
        c[d] = (int)(ff % P.N[d]);
        ff /= P.N[d];
      }
    c[dir] = layer;
    const long long cell = cell_index(P.N, c[0], c[1], c[2]);
    if (pack)
      buf[i] = x[cell * P.n + k];
    else
      x[cell * P.n + k] = buf[i];
  }
}

}  // namespace

void launch_halo_copy(const DevParams& P, double* x, double* buf, int dir, int side, bool pack, cudaStream_t s) {
  if (dir < 0 || dir >= P.dim || side < 0 || side > 1) throw Error("invalid (dir, side)");
  if (P.N[dir] < 3) throw Error("halo exchange needs at least 3 cell layers in the exchange direction");
  // pack: owned layer next to the ghost layer; unpack: the ghost layer itself
  const int layer = pack ? (side ? P.N[dir] - 2 : 1) : (side ? P.N[dir] - 1 : 0);
  const long long total = (P.ncells / P.N[dir]) * P.n;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148 * 16);
  halo_copy_kernel<<<blocks, threads, 0, s>>>(P, x, buf, dir, layer, pack ? 1 : 0, total);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace pdb

// ---------------------------------------------------------------------------------------------
// Peer-to-peer halo exchange over NVLink: every rank owns a MAILBOX in device memory (receive
// buffers for its processor sides + flags), shared with its face neighbours through CUDA IPC.
//   push   (sender):   wait until the neighbour has consumed the previous epoch (ack), copy the owned
//                      boundary layer of x straight into the neighbour's receive buffer with remote
//                      stores, __threadfence_system(), then the last CTA publishes ready = epoch.
//   unpack (receiver): spin on the local ready flag, copy the receive buffer into the ghost layer
//                      of x, then the last CTA sends ack = epoch to the sender.
// No NCCL call, no host round trip: two kernels per exchange regardless of the number of sides,
// and they run on a high-priority side stream while the interior tiles are computed.
// Replaces the CopyDataHandle communication of boilerplate/pdelab.hh:872-880 /
// gridfunctionspace/genericdatahandle.hh on the reference's overlapping partition.

namespace pdb {

namespace {

struct MailboxHeader {
  unsigned long long ready[6];      // [my side] epoch of the data the neighbour across that side has delivered
  unsigned long long ack[6];        // [my side] last epoch the neighbour across that side has consumed
  unsigned long long buf_off[6];    // byte offset of the receive buffer of each side
  unsigned long long layer_doubles[6];  // doubles received across each side
  unsigned long long send_doubles[6];   // conforming Qk: doubles sent across each side (k vs k+1 planes)
  unsigned long long magic;
};
constexpr unsigned long long MAILBOX_MAGIC = 0x70646232303068ull;  // "pdb200h"
constexpr int P2P_BLOCKS = 64;  // CTAs per side
constexpr unsigned long long SPIN_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;

// conforming Qk: a slab of lattice planes lo_d .. lo_d + n_d - 1 over the full tangential extent of the local box
struct QkBox {
  int lo[3], n[3];
  long long count;
};

struct SideDesc {
  int active, dir, layer_src, layer_dst;
  QkBox qsend, qrecv;              // conforming Qk: the planes sent across this side / received from it
  long long total;                 // doubles in the layer
  long long chunk, stride;         // contiguous run and its repeat stride, in doubles (see copy_layer)
  double* peer_buf;                // receive buffer in the NEIGHBOUR's mailbox for its side (dir, 1-side)
  unsigned long long* peer_ready;  // neighbour's ready[(dir, 1-side)]
  unsigned long long* peer_ack;    // neighbour's ack[(dir, 1-side)]
  const double* my_buf;            // my receive buffer for this side
  unsigned long long* my_ready;
  unsigned long long* my_ack;
};
struct SideTable {
  SideDesc s[6];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// A cell layer normal to `dir` is a set of contiguous chunks: the block of all lower directions
// (chunk = n * prod_{d<dir} N_d doubles), repeated with stride chunk * N_dir.  z-layers are one
// contiguous block, y-layers one chunk per z.  Element i of the packed layer (lexicographic
// tangential order, lower directions fastest) lives at  layer*chunk + (i / chunk)*stride + i % chunk.
template <typename T>
__device__ __forceinline__ long long layer_elem(long long i, long long chunk, long long stride, long long layer_off) {
  const long long c = i / chunk;
  return layer_off + c * stride + (i - c * chunk);
}

template <typename T>
__device__ __forceinline__ void copy_layer(T* __restrict__ dst, const T* __restrict__ src, long long total, long long chunk,
                                           long long stride, long long layer_off, bool strided_src) {
  // grid-stride, four independent loads in flight per thread
  const long long step = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * step < total; i += 4 * step) {
    T v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long j = i + u * step;
      v[u] = src[strided_src ? layer_elem<T>(j, chunk, stride, layer_off) : j];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long j = i + u * step;
      dst[strided_src ? j : layer_elem<T>(j, chunk, stride, layer_off)] = v[u];
    }
  }
  for (; i < total; i += step) {
    const T v = src[strided_src ? layer_elem<T>(i, chunk, stride, layer_off) : i];
    dst[strided_src ? i : layer_elem<T>(i, chunk, stride, layer_off)] = v;
  }
}

// Waiting is done by ONE warp (lane s watches side s), in its own tiny kernel, so that no wide
// copy kernel sits on SM resources while it spins: the copy kernels below are launched behind it
// on the same stream and run only once the data / the free mailbox is there.
//   which = 0: my_ack >= want (the neighbour has emptied its receive buffer), 1: my_ready >= want
__global__ void p2p_wait_kernel(const SideTable T, int which, unsigned long long want, int* err) {
  const int s = threadIdx.x;
  if (s < 6 && T.s[s].active) {
    const unsigned long long* flag = which ? T.s[s].my_ready : T.s[s].my_ack;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < want) {
      if (globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) {
        atomicExch(err, 1);
        break;
      }
      __nanosleep(100);
    }
  }
}

// grid = (blocks per side, 6)
template <int VEC>
__global__ void p2p_push_kernel(const DevParams P, const SideTable T, const double* __restrict__ x,
                                unsigned long long epoch, unsigned int* counters, int* err) {
  const SideDesc& S = T.s[blockIdx.y];
  if (!S.active) return;
  if (VEC == 2)
    copy_layer<double2>((double2*)S.peer_buf, (const double2*)x, S.total / 2, S.chunk / 2, S.stride / 2,
                        (long long)S.layer_src * (S.chunk / 2), true);
  else
    copy_layer<double>(S.peer_buf, x, S.total, S.chunk, S.stride, (long long)S.layer_src * S.chunk, true);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&counters[blockIdx.y], 1u);
    if (done == gridDim.x - 1) {
      counters[blockIdx.y] = 0;
      __threadfence_system();
      st_release_sys(S.peer_ready, epoch);
    }
  }
}

template <int VEC>
__global__ void p2p_unpack_kernel(const DevParams P, const SideTable T, double* __restrict__ x,
                                  unsigned long long epoch, unsigned int* counters, int* err) {
  const SideDesc& S = T.s[blockIdx.y];
  if (!S.active) return;
  if (VEC == 2)
    copy_layer<double2>((double2*)x, (const double2*)S.my_buf, S.total / 2, S.chunk / 2, S.stride / 2,
                        (long long)S.layer_dst * (S.chunk / 2), false);
  else
    copy_layer<double>(x, S.my_buf, S.total, S.chunk, S.stride, (long long)S.layer_dst * S.chunk, false);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(&counters[6 + blockIdx.y], 1u);
    if (done == gridDim.x - 1) {
      counters[6 + blockIdx.y] = 0;
      st_release_sys(S.peer_ack, epoch);
    }
  }
}

// ---- conforming Qk on the same mailboxes ---------------------------------------------------------------------
// Every lattice point has ONE owner: the interface plane between two ranks' owned cells belongs to the lower rank
// (the rule of gridfunctionspace/genericdatahandle.hh:894-947 on a Cartesian partition).  Across its upper side a
// rank sends the k+1 lattice planes of its last owned cell layer and receives the k planes beyond the interface,
// across its lower side it sends k planes and receives k+1.  The planes are scattered over the sub-entity groups of
// the container; the container index of a lattice point is closed-form arithmetic (qk_lattice_index), so pack and
// unpack need no index arrays.  Directions are exchanged one after the other over the full tangential extent: edge
// and corner neighbours are reached in two / three hops, without messages of their own.
__device__ __forceinline__ long long qk_box_index(const QkLayout& L, const QkBox& B, long long i) {
  int l[3] = {0, 0, 0};
  for (int d = 0; d < L.dim; d++) {
    l[d] = B.lo[d] + (int)(i % B.n[d]);
    i /= B.n[d];
  }
  return qk_lattice_index(L, l);
}

// grid = (blocks per side, 2): the two sides of direction `dir`
__global__ void qk_push_kernel(const QkLayout L, const SideTable T, int dir, const double* __restrict__ x,
                               unsigned long long epoch, unsigned int* counters) {
  const int sidx = 2 * dir + blockIdx.y;
  const SideDesc& S = T.s[sidx];
  if (!S.active) return;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S.qsend.count; i += step)
    S.peer_buf[i] = x[qk_box_index(L, S.qsend, i)];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&counters[sidx], 1u);
    if (done == gridDim.x - 1) {
      counters[sidx] = 0;
      __threadfence_system();
      st_release_sys(S.peer_ready, epoch);
    }
  }
}

__global__ void qk_unpack_kernel(const QkLayout L, const SideTable T, int dir, double* __restrict__ x,
                                 unsigned long long epoch, unsigned int* counters) {
  const int sidx = 2 * dir + blockIdx.y;
  const SideDesc& S = T.s[sidx];
  if (!S.active) return;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S.qrecv.count; i += step)
    x[qk_box_index(L, S.qrecv, i)] = S.my_buf[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(&counters[6 + sidx], 1u);
    if (done == gridDim.x - 1) {
      counters[6 + sidx] = 0;
      st_release_sys(S.peer_ack, epoch);
    }
  }
}

// one warp waits for the two sides of a direction (see p2p_wait_kernel)
__global__ void qk_wait_kernel(const SideTable T, int dir, int which, unsigned long long want, int* err) {
  const int s = 2 * dir + threadIdx.x;
  if (threadIdx.x < 2 && T.s[s].active) {
    const unsigned long long* flag = which ? T.s[s].my_ready : T.s[s].my_ack;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < want) {
      if (globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) {
        atomicExch(err, 1);
        break;
      }
      __nanosleep(100);
    }
  }
}

// x := 0 on the lattice points this rank does not own (the slabs below / above the owned range of every direction)
struct QkBoxes6 {
  QkBox b[6];
};
__global__ void qk_zero_boxes_kernel(const QkLayout L, const QkBoxes6 B, double* __restrict__ x) {
  const QkBox& box = B.b[blockIdx.y];
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < box.count; i += step)
    x[qk_box_index(L, box, i)] = 0.0;
}

}  // namespace

struct P2PHalo {
  bool qk = false;                   // conforming Qk: lattice-plane exchange, direction by direction
  QkLayout L;
  QkBoxes6 nonowned;                 // conforming Qk: the lattice points of the local box this rank does not own
  unsigned char* mailbox = nullptr;  // device memory of this rank (cudaMalloc, IPC-exported)
  size_t mailbox_bytes = 0;
  MailboxHeader header;              // host copy
  void* peer_base[6] = {};           // mapped mailboxes of the neighbours
  SideTable table;
  unsigned int* counters = nullptr;  // [12]
  int* err = nullptr;
  unsigned long long epoch = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {};
  int nactive = 0;
  bool vec2 = true;  // every chunk is an even number of doubles: 16-byte accesses
  FusedTable fused_host;             // the one-launch step of dg_fast.cu (QkDG)
  FusedTable* fused_dev = nullptr;
  bool fused_built = false;
};

// lattice slab of direction d, planes a..b, full extent elsewhere
static QkBox qk_slab(const DevParams& P, int d, int a, int b) {
  QkBox B;
  B.count = 1;
  for (int e = 0; e < 3; e++) {
    B.lo[e] = 0;
    B.n[e] = e < P.dim ? P.k * P.N[e] + 1 : 1;
    if (e == d) {
      B.lo[e] = a;
      B.n[e] = std::max(0, b - a + 1);
    }
    B.count *= B.n[e];
  }
  return B;
}

P2PHalo* p2p_create(const DevParams& P, pdb200_ipc_handle* mine) {
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(pdb200_ipc_handle), "handle size");
  P2PHalo* H = new P2PHalo;
  std::memset(&H->header, 0, sizeof(H->header));
  std::memset(&H->table, 0, sizeof(H->table));
  std::memset(&H->nonowned, 0, sizeof(H->nonowned));
  H->qk = !P.dg;
  if (H->qk) H->L = make_qk_layout(P);
  size_t off = (sizeof(MailboxHeader) + 255) / 256 * 256;
  for (int d = 0; d < P.dim; d++)
    for (int s = 0; s < 2; s++) {
      if (P.side_kind[d][s] != PDB200_SIDE_PROCESSOR) continue;
      if (P.N[d] < 3) throw Error("halo exchange needs at least 3 cell layers in the exchange direction");
      unsigned long long n = (unsigned long long)(P.ncells / P.N[d]) * P.n;
      if (H->qk) {
        const int k = P.k, nd = P.N[d];
        SideDesc& S = H->table.s[2 * d + s];
        // upper side: I own the interface plane; lower side: the neighbour does
        S.qsend = s ? qk_slab(P, d, k * (nd - 2), k * (nd - 1)) : qk_slab(P, d, k + 1, 2 * k);
        S.qrecv = s ? qk_slab(P, d, k * (nd - 1) + 1, k * nd) : qk_slab(P, d, 0, k);
        H->nonowned.b[2 * d + s] = S.qrecv;   // exactly the planes received are the ones not owned
        n = (unsigned long long)S.qrecv.count;
        H->header.send_doubles[2 * d + s] = (unsigned long long)S.qsend.count;
      }
      H->header.buf_off[2 * d + s] = off;
      H->header.layer_doubles[2 * d + s] = n;
      off += (n * sizeof(double) + 255) / 256 * 256;
    }
  H->header.magic = MAILBOX_MAGIC;
  H->mailbox_bytes = off;
  PDB_CUDA(cudaMalloc(&H->mailbox, off));
  PDB_CUDA(cudaMemset(H->mailbox, 0, off));
  PDB_CUDA(cudaMemcpy(H->mailbox, &H->header, sizeof(H->header), cudaMemcpyHostToDevice));
  PDB_CUDA(cudaMalloc(&H->counters, 12 * sizeof(unsigned int)));
  PDB_CUDA(cudaMemset(H->counters, 0, 12 * sizeof(unsigned int)));
  PDB_CUDA(cudaMalloc(&H->err, sizeof(int)));
  PDB_CUDA(cudaMemset(H->err, 0, sizeof(int)));
  int lo = 0, hi = 0;
  PDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  PDB_CUDA(cudaStreamCreateWithPriority(&H->stream, cudaStreamNonBlocking, hi));
  for (auto& e : H->ev) PDB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaIpcMemHandle_t ih;
  PDB_CUDA(cudaIpcGetMemHandle(&ih, H->mailbox));
  std::memset(mine, 0, sizeof(*mine));
  std::memcpy(mine->bytes, &ih, sizeof(ih));
  PDB_CUDA(cudaDeviceSynchronize());
  return H;
}

void p2p_connect(P2PHalo* H, const DevParams& P, int dir, int side, const pdb200_ipc_handle* peer) {
  if (dir < 0 || dir >= P.dim || side < 0 || side > 1) throw Error("invalid (dir, side)");
  if (P.side_kind[dir][side] != PDB200_SIDE_PROCESSOR) throw Error("p2p_connect: not a processor side");
  const int s = 2 * dir + side, sp = 2 * dir + (1 - side);  // my side, the neighbour's side facing me
  cudaIpcMemHandle_t ih;
  std::memcpy(&ih, peer->bytes, sizeof(ih));
  void* base = nullptr;
  PDB_CUDA(cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess));
  H->peer_base[s] = base;
  MailboxHeader ph;
  PDB_CUDA(cudaMemcpy(&ph, base, sizeof(ph), cudaMemcpyDeviceToHost));
  if (ph.magic != MAILBOX_MAGIC) throw Error("p2p_connect: the peer handle is not a pdelab_b200 mailbox");
  if (H->qk ? (ph.layer_doubles[sp] != H->header.send_doubles[s] || ph.send_doubles[sp] != H->header.layer_doubles[s])
            : ph.layer_doubles[sp] != H->header.layer_doubles[s])
    throw Error("p2p_connect: the neighbour's layer size differs (inconsistent partition)");
  unsigned char* pb = (unsigned char*)base;
  SideDesc& S = H->table.s[s];
  S.active = 1;
  S.dir = dir;
  S.layer_src = side ? P.N[dir] - 2 : 1;
  S.layer_dst = side ? P.N[dir] - 1 : 0;
  S.total = (long long)H->header.layer_doubles[s];
  S.chunk = P.n;
  for (int d = 0; d < dir; d++) S.chunk *= P.N[d];
  S.stride = S.chunk * P.N[dir];
  if (S.chunk % 2) H->vec2 = false;
  S.peer_buf = (double*)(pb + ph.buf_off[sp]);
  S.peer_ready = (unsigned long long*)(pb + offsetof(MailboxHeader, ready)) + sp;
  S.peer_ack = (unsigned long long*)(pb + offsetof(MailboxHeader, ack)) + sp;
  S.my_buf = (const double*)(H->mailbox + H->header.buf_off[s]);
  S.my_ready = (unsigned long long*)(H->mailbox + offsetof(MailboxHeader, ready)) + s;
  S.my_ack = (unsigned long long*)(H->mailbox + offsetof(MailboxHeader, ack)) + s;
  H->nactive++;
}

void p2p_destroy(P2PHalo* H) {
  if (!H) return;
  cudaDeviceSynchronize();
  for (void* b : H->peer_base)
    if (b) cudaIpcCloseMemHandle(b);
  if (H->mailbox) cudaFree(H->mailbox);
  if (H->counters) cudaFree(H->counters);
  if (H->err) cudaFree(H->err);
  if (H->fused_dev) cudaFree(H->fused_dev);
  if (H->stream) cudaStreamDestroy(H->stream);
  for (auto& e : H->ev)
    if (e) cudaEventDestroy(e);
  delete H;
}

static void p2p_require_connected(P2PHalo* H, const DevParams& P) {
  int need = 0;
  for (int d = 0; d < P.dim; d++)
    for (int s = 0; s < 2; s++) need += P.side_kind[d][s] == PDB200_SIDE_PROCESSOR;
  if (H->nactive != need) throw Error("p2p halo: not every processor side is connected");
}

// owner -> ghost copy of a conforming Qk vector: per direction wait(ack), push, wait(ready), unpack
static int qk_exchange(P2PHalo* H, const DevParams& P, double* x, cudaStream_t s) {
  p2p_require_connected(H, P);
  if (H->nactive == 0) return 0;
  H->epoch++;
  int launches = 0;
  for (int d = 0; d < P.dim; d++) {
    if (!H->table.s[2 * d].active && !H->table.s[2 * d + 1].active) continue;
    qk_wait_kernel<<<1, 32, 0, s>>>(H->table, d, 0, H->epoch - 1, H->err);
    qk_push_kernel<<<dim3(P2P_BLOCKS, 2), 256, 0, s>>>(H->L, H->table, d, x, H->epoch, H->counters);
    qk_wait_kernel<<<1, 32, 0, s>>>(H->table, d, 1, H->epoch, H->err);
    qk_unpack_kernel<<<dim3(P2P_BLOCKS, 2), 256, 0, s>>>(H->L, H->table, d, x, H->epoch, H->counters);
    launches += 4;
  }
  PDB_CUDA(cudaGetLastError());
  return launches;
}

// owner -> ghost copy of x across all connected sides (QkDG: two kernels; conforming Qk: direction by direction)
int p2p_exchange(P2PHalo* H, const DevParams& P, double* x, cudaStream_t s) {
  if (H->qk) return qk_exchange(H, P, x, s);
  return p2p_push(H, P, x, s) + p2p_wait_unpack(H, P, x, s);
}

// x := 0 on everything this rank does not own: the ghost cell layers (QkDG) / the lattice planes it receives (Qk)
int p2p_zero_ghosts(P2PHalo* H, const DevParams& P, double* x, cudaStream_t s) {
  if (!H->qk) return launch_halo_zero(P, x, s);
  if (H->nactive == 0) return 0;
  qk_zero_boxes_kernel<<<dim3(P2P_BLOCKS, 6), 256, 0, s>>>(H->L, H->nonowned, x);
  PDB_CUDA(cudaGetLastError());
  return 1;
}

bool p2p_is_qk(const P2PHalo* H) { return H->qk; }

int p2p_push(P2PHalo* H, const DevParams& P, const double* x, cudaStream_t s) {
  if (H->qk) throw Error("p2p_push: conforming spaces exchange direction by direction (p2p_exchange)");
  p2p_require_connected(H, P);
  if (H->nactive == 0) return 0;
  H->epoch++;
  p2p_wait_kernel<<<1, 32, 0, s>>>(H->table, 0, H->epoch - 1, H->err);
  if (H->vec2 && (uintptr_t)x % 16 == 0)
    p2p_push_kernel<2><<<dim3(P2P_BLOCKS, 6), 256, 0, s>>>(P, H->table, x, H->epoch, H->counters, H->err);
  else
    p2p_push_kernel<1><<<dim3(P2P_BLOCKS, 6), 256, 0, s>>>(P, H->table, x, H->epoch, H->counters, H->err);
  PDB_CUDA(cudaGetLastError());
  return 2;
}

int p2p_wait_unpack(P2PHalo* H, const DevParams& P, double* x, cudaStream_t s) {
  if (H->nactive == 0) return 0;
  p2p_wait_kernel<<<1, 32, 0, s>>>(H->table, 1, H->epoch, H->err);
  if (H->vec2 && (uintptr_t)x % 16 == 0)
    p2p_unpack_kernel<2><<<dim3(P2P_BLOCKS, 6), 256, 0, s>>>(P, H->table, x, H->epoch, H->counters, H->err);
  else
    p2p_unpack_kernel<1><<<dim3(P2P_BLOCKS, 6), 256, 0, s>>>(P, H->table, x, H->epoch, H->counters, H->err);
  PDB_CUDA(cudaGetLastError());
  return 2;
}

void p2p_check(P2PHalo* H) {
  int e = 0;
  PDB_CUDA(cudaMemcpy(&e, H->err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) {
    cudaMemset(H->err, 0, sizeof(int));
    throw Error("p2p halo exchange timed out waiting for a neighbour");
  }
}

const FusedTable* p2p_fused_table(P2PHalo* H, const DevParams& P) {
  if (H->qk || !H->vec2 || H->nactive == 0) return nullptr;
  if (!H->fused_built) {
    p2p_require_connected(H, P);
    std::memset(&H->fused_host, 0, sizeof(FusedTable));
    for (int i = 0; i < 6; i++) {
      const SideDesc& S = H->table.s[i];
      FusedSide& F = H->fused_host.s[i];
      F.active = S.active;
      if (!S.active) continue;
      F.total2 = S.total / 2;
      F.chunk2 = S.chunk / 2;
      F.stride2 = S.stride / 2;
      F.src_off2 = (long long)S.layer_src * (S.chunk / 2);
      F.peer_buf = (double2*)S.peer_buf;
      F.peer_ready = S.peer_ready;
      F.peer_ack = S.peer_ack;
      F.my_buf = S.my_buf;
      F.my_ready = S.my_ready;
      F.my_ack = S.my_ack;
    }
    H->fused_host.counters = H->counters;
    H->fused_host.err = H->err;
    PDB_CUDA(cudaMalloc(&H->fused_dev, sizeof(FusedTable)));
    PDB_CUDA(cudaMemcpy(H->fused_dev, &H->fused_host, sizeof(FusedTable), cudaMemcpyHostToDevice));
    H->fused_built = true;
  }
  return H->fused_dev;
}
const FusedTable& p2p_fused_table_host(P2PHalo* H) { return H->fused_host; }
unsigned long long p2p_next_epoch(P2PHalo* H) { return ++H->epoch; }

cudaStream_t p2p_stream(P2PHalo* H) { return H->stream; }
cudaEvent_t p2p_event(P2PHalo* H, int i) { return H->ev[i]; }

// ---------------------------------------------------------------------------------------------
// All-ranks reduction over peer-mapped mailboxes: the global sum of OverlappingScalarProduct::dot
// (backend/istl/ovlpistlsolverbackend.hh:103-108, gridView().comm().sum) without NCCL and without a host
// round trip.  Every rank owns a small mailbox (CUDA IPC, mapped by all peers) with one value slot and one
// flag per rank, double-buffered by the parity of the epoch.  One kernel of one CTA per reduction:
//   1. sum the block partials of the local inner product(s) in the fixed order every consumer uses;
//   2. thread t stores the local sums into rank t's mailbox (remote store), fences, publishes flag = epoch;
//   3. thread t spins on the local flag of rank t (time-out into the error flag, never a hang);
//   4. the sums of all ranks are added in rank order — the same order on every rank, so all ranks hold
//      bit-identical scalars — and written back as partial 0 (the other partials are zeroed), which is what
//      the consuming vector kernels re-add.
// A rank can run at most one reduction ahead of a peer (it needs the peer's flag of the current epoch to
// finish), so two buffers suffice.
namespace {

constexpr int COMM_MAX_RANKS = 16;
constexpr unsigned long long COMM_MAGIC = 0x70646232303063ull;  // "pdb200c"
struct CommBox {
  unsigned long long magic;
  unsigned long long flag[2][COMM_MAX_RANKS];
  double val[2][COMM_MAX_RANKS][2];
};
struct CommTable {
  int rank, size;
  CommBox* box[COMM_MAX_RANKS];  // box[rank] is this rank's own mailbox
};

__global__ void __launch_bounds__(256) comm_allreduce_kernel(const CommTable T, double* P1, double* P2, int nb,
                                                             unsigned long long epoch, int* err) {
  __shared__ double ws[2][8];
  __shared__ double tot[2];
  double v1 = 0.0, v2 = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) {
    v1 += P1[i];
    if (P2) v2 += P2[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v1 += __shfl_down_sync(0xffffffffu, v1, o);
    v2 += __shfl_down_sync(0xffffffffu, v2, o);
  }
  if ((threadIdx.x & 31) == 0) ws[0][threadIdx.x >> 5] = v1, ws[1][threadIdx.x >> 5] = v2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) a += ws[0][w], b += ws[1][w];
    tot[0] = a, tot[1] = b;
  }
  __syncthreads();
  const int par = (int)(epoch & 1);
  const int t = threadIdx.x;
  if (t < T.size) {
    CommBox* dst = T.box[t];
    volatile double* slot = dst->val[par][T.rank];
    slot[0] = tot[0];
    slot[1] = tot[1];
    __threadfence_system();
    st_release_sys(&dst->flag[par][T.rank], epoch);
    const unsigned long long* flag = &T.box[T.rank]->flag[par][t];
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < epoch) {
      if (globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) {
        atomicExch(err, 1);
        break;
      }
      __nanosleep(50);
    }
  }
  __syncthreads();
  if (t == 0) {
    const CommBox* me = T.box[T.rank];
    double g1 = 0.0, g2 = 0.0;
    for (int r = 0; r < T.size; r++) {
      const volatile double* slot = me->val[par][r];
      g1 += slot[0];
      g2 += slot[1];
    }
    P1[0] = g1;
    if (P2) P2[0] = g2;
  }
  for (int i = 1 + t; i < nb; i += 256) {
    P1[i] = 0.0;
    if (P2) P2[i] = 0.0;
  }
}

// zero one cell layer (the ghost layer of a processor side)
__global__ void halo_zero_kernel(double* __restrict__ x, long long total, long long chunk, long long stride, long long layer_off) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step)
    x[layer_elem<double>(i, chunk, stride, layer_off)] = 0.0;
}

}  // namespace

struct PeerComm {
  CommBox* box = nullptr;  // own mailbox (device memory, IPC-exported)
  CommTable table;
  void* peer_base[COMM_MAX_RANKS] = {};
  int connected = 1;
  unsigned long long epoch = 0;
  int* err = nullptr;
};

PeerComm* comm_create(int rank, int size, pdb200_ipc_handle* mine) {
  if (size < 1 || size > COMM_MAX_RANKS || rank < 0 || rank >= size) throw Error("comm_create: rank / size out of range (at most 16 ranks)");
  PeerComm* C = new PeerComm;
  std::memset(&C->table, 0, sizeof(C->table));
  PDB_CUDA(cudaMalloc(&C->box, sizeof(CommBox)));
  PDB_CUDA(cudaMemset(C->box, 0, sizeof(CommBox)));
  const unsigned long long magic = COMM_MAGIC;
  PDB_CUDA(cudaMemcpy(&C->box->magic, &magic, sizeof(magic), cudaMemcpyHostToDevice));
  PDB_CUDA(cudaMalloc(&C->err, sizeof(int)));
  PDB_CUDA(cudaMemset(C->err, 0, sizeof(int)));
  C->table.rank = rank;
  C->table.size = size;
  C->table.box[rank] = C->box;
  cudaIpcMemHandle_t ih;
  PDB_CUDA(cudaIpcGetMemHandle(&ih, C->box));
  std::memset(mine, 0, sizeof(*mine));
  std::memcpy(mine->bytes, &ih, sizeof(ih));
  PDB_CUDA(cudaDeviceSynchronize());
  return C;
}

void comm_connect(PeerComm* C, int peer_rank, const pdb200_ipc_handle* peer) {
  if (peer_rank < 0 || peer_rank >= C->table.size || peer_rank == C->table.rank) throw Error("comm_connect: invalid peer rank");
  if (C->peer_base[peer_rank]) throw Error("comm_connect: peer already connected");
  cudaIpcMemHandle_t ih;
  std::memcpy(&ih, peer->bytes, sizeof(ih));
  void* base = nullptr;
  PDB_CUDA(cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess));
  unsigned long long magic = 0;
  PDB_CUDA(cudaMemcpy(&magic, base, sizeof(magic), cudaMemcpyDeviceToHost));
  if (magic != COMM_MAGIC) {
    cudaIpcCloseMemHandle(base);
    throw Error("comm_connect: the peer handle is not a pdelab_b200 reduction mailbox");
  }
  C->peer_base[peer_rank] = base;
  C->table.box[peer_rank] = (CommBox*)base;
  C->connected++;
}

void comm_destroy(PeerComm* C) {
  if (!C) return;
  cudaDeviceSynchronize();
  for (void* b : C->peer_base)
    if (b) cudaIpcCloseMemHandle(b);
  if (C->box) cudaFree(C->box);
  if (C->err) cudaFree(C->err);
  delete C;
}

int comm_size(const PeerComm* C) { return C->table.size; }

void comm_allreduce_partials(PeerComm* C, double* P1, double* P2, int nb, cudaStream_t s) {
  if (C->connected != C->table.size) throw Error("peer reduction: not every rank is connected");
  C->epoch++;
  comm_allreduce_kernel<<<1, 256, 0, s>>>(C->table, P1, P2, nb, C->epoch, C->err);
  PDB_CUDA(cudaGetLastError());
}

void comm_check(PeerComm* C) {
  int e = 0;
  PDB_CUDA(cudaMemcpy(&e, C->err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) {
    cudaMemset(C->err, 0, sizeof(int));
    throw Error("peer reduction timed out waiting for a rank");
  }
}

// x := 0 on the ghost layers of all processor sides (set_constrained_dofs(cc, 0.0, y) for the P0 parallel
// constraints, backend/istl/ovlpistlsolverbackend.hh:48-49; constraints/p0.hh:31-41)
int launch_halo_zero(const DevParams& P, double* x, cudaStream_t s) {
  int launches = 0;
  for (int d = 0; d < P.dim; d++)
    for (int side = 0; side < 2; side++) {
      if (P.side_kind[d][side] != PDB200_SIDE_PROCESSOR) continue;
      long long chunk = P.n;
      for (int e = 0; e < d; e++) chunk *= P.N[e];
      const long long stride = chunk * P.N[d];
      const long long total = (P.ncells / P.N[d]) * P.n;
      const long long layer = side ? P.N[d] - 1 : 0;
      const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
      halo_zero_kernel<<<blocks, 256, 0, s>>>(x, total, chunk, stride, layer * chunk);
      launches++;
    }
  PDB_CUDA(cudaGetLastError());
  return launches;
}

// index gather / scatter (pack / unpack of the conforming-Qk ghost exchange)
__global__ void gather_kernel(const double* __restrict__ x, const long long* __restrict__ idx, long long n,
                              double* __restrict__ buf) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    buf[i] = x[idx[i]];
}
__global__ void scatter_kernel(const double* __restrict__ buf, const long long* __restrict__ idx, long long n,
                               double* __restrict__ x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[idx[i]] = buf[i];
}
void launch_gather(const double* x, const long long* idx, long long n, double* buf, cudaStream_t s) {
  if (n <= 0) return;
  gather_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(x, idx, n, buf);
  PDB_CUDA(cudaGetLastError());
}
void launch_scatter(const double* buf, const long long* idx, long long n, double* x, cudaStream_t s) {
  if (n <= 0) return;
  scatter_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(buf, idx, n, x);
  PDB_CUDA(cudaGetLastError());
}

}  // namespace pdb

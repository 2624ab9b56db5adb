// halo.cu — pack / unpack of one cell layer of a QkDG vector for the overlapping partition.
//
// Replaces the per-entity gather/scatter of GFSDataHandle + CopyGatherScatter
// (gridfunctionspace/genericdatahandle.hh:130-260) used by the overlapping backends
// (boilerplate/pdelab.hh:872-880): with overlap 1 the data sent to the neighbour across side
// (dir, side) is the owned cell layer at distance 1 from that box face, and the data received
// fills the ghost layer at distance 0.  A DG cell is n contiguous doubles, so the copy is a
// strided block copy; consecutive threads move consecutive doubles of a cell.
#include "common.cuh"

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

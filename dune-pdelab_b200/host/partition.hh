// partition.hh — C++ host side of the multi-GPU path: the overlapping Cartesian partition of a YaspGrid, the peer-to-peer
// halo exchanger and the overlapping Krylov back-end over the C ABI (one process per GPU).
//
// What it mirrors in the reference (paths relative to dune/pdelab/): an overlapping YaspGrid (overlap >= 1) splits the
// cell index box into a Cartesian processor grid (dune-grid's Torus); every rank stores its interior block plus `overlap`
// ghost cell layers towards each neighbour and assembles on interior + overlap cells (gridoperator/default/assembler.hh:116);
// consistency of input vectors is restored by CopyDataHandle communication (boilerplate/pdelab.hh:872-880,
// gridfunctionspace/genericdatahandle.hh), rows of cells at a processor boundary are constrained (constraints/p0.hh:31-41),
// and the ISTLBackend_OVLP_* back-ends run a Krylov loop over OverlappingOperator / OverlappingScalarProduct
// (backend/istl/ovlpistlsolverbackend.hh:30-134, 477-560).
//
// The index arithmetic here is the same as python/pdelab_b200/partition.py (tests/test_cpp_partition.py compares the two
// rank by rank).  The set-up handshake — every rank's 64-byte mailbox handle to its neighbours / to all ranks — is the one
// collective a host program has to supply (MPI_Allgather in a DUNE program): it is passed in as a call-back, so this header
// depends on nothing but the C ABI.
#ifndef PDELAB_B200_HOST_PARTITION_HH
#define PDELAB_B200_HOST_PARTITION_HH

#include <array>
#include <functional>
#include <tuple>
#include <vector>

#include "gridoperator.hh"

namespace Dune {
namespace PDELab {
namespace B200 {

// Cartesian processor grid, as cubic as possible, larger factors last.  By default direction 0 is not split
// (2 -> 1x1x2, 4 -> 1x2x2, 8 -> 1x2x4): the x-rows of a QkDG vector stay contiguous and the local cell count in x stays
// even, which the TMA tiling of the fast kernel needs; split_x gives YaspGrid's most-cubic grid (2x2x2 at 8 ranks).
inline std::array<int, 3> processor_grid(int world, int dim = 3, bool split_x = false) {
  const int ndim = split_x ? dim : dim - 1;
  if (world < 1 || dim < 2 || dim > 3) throw Exception("processor_grid: world >= 1 and dim in {2, 3}");
  std::vector<int> best;
  int best_spread = -1;
  std::vector<int> cur;
  std::function<void(int, int)> rec = [&](int rem, int k) {
    if (k == 1) {
      std::vector<int> cand(cur);
      cand.push_back(rem);
      std::sort(cand.begin(), cand.end());
      const int spread = cand.back() - cand.front();
      if (best_spread < 0 || spread < best_spread || (spread == best_spread && cand < best)) {
        best_spread = spread;
        best = cand;
      }
      return;
    }
    for (int f = 1; f <= rem; f++)
      if (rem % f == 0) {
        cur.push_back(f);
        rec(rem / f, k - 1);
        cur.pop_back();
      }
  };
  rec(world, ndim);
  std::array<int, 3> grid{1, 1, 1};
  for (int i = 0; i < ndim; i++) grid[split_x ? i : i + 1] = best[i];
  return grid;
}

// One rank's view of the overlapping partition (owned block + ghost layers, local numbering lexicographic over the
// extended box like a YaspGrid rank).
template <int dim>
class OverlappingPartition {
 public:
  OverlappingPartition(const std::array<int, dim>& global_cells, const std::array<int, 3>& procs, int rank, int overlap = 1,
                       const FieldVector<double, dim>& lower = FieldVector<double, dim>(0.0),
                       const FieldVector<double, dim>& upper = FieldVector<double, dim>(1.0))
      : global_cells(global_cells), procs(procs), rank(rank), overlap(overlap) {
    world = 1;
    for (int d = 0; d < dim; d++) world *= procs[d];
    if (rank < 0 || rank >= world) throw Exception("OverlappingPartition: rank out of range");
    int r = rank;  // rank = px + Px (py + Py pz): lexicographic torus coordinates, x fastest
    for (int d = 0; d < dim; d++) {
      coords[d] = r % procs[d];
      r /= procs[d];
    }
    for (auto& s : side_kind) s = {PDB200_SIDE_DOMAIN, PDB200_SIDE_DOMAIN};
    for (auto& n : neighbour) n = {-1, -1};
    for (int d = 0; d < dim; d++) {
      const long long n = global_cells[d], p = procs[d], i = coords[d];
      owned_lo[d] = (int)((n * i) / p);
      owned_hi[d] = (int)((n * (i + 1)) / p);
      local_lo[d] = owned_lo[d];
      local_hi[d] = owned_hi[d];
      if (i > 0) {
        local_lo[d] -= overlap;
        side_kind[d][0] = PDB200_SIDE_PROCESSOR;
        auto c = coords;
        c[d] -= 1;
        neighbour[d][0] = rank_of(c);
      }
      if (i < p - 1) {
        local_hi[d] += overlap;
        side_kind[d][1] = PDB200_SIDE_PROCESSOR;
        auto c = coords;
        c[d] += 1;
        neighbour[d][1] = rank_of(c);
      }
      owned_cells[d] = owned_hi[d] - owned_lo[d];
      local_cells[d] = local_hi[d] - local_lo[d];
      const double h = (upper[d] - lower[d]) / global_cells[d];
      local_lower[d] = lower[d] + h * local_lo[d];
      local_upper[d] = lower[d] + h * local_hi[d];
    }
  }
  // fixed work per rank / fixed total work
  static OverlappingPartition weak(const std::array<int, dim>& cells_per_rank, int world, int rank, int overlap = 1) {
    const auto procs = processor_grid(world, dim);
    std::array<int, dim> glob;
    for (int d = 0; d < dim; d++) glob[d] = cells_per_rank[d] * procs[d];
    return OverlappingPartition(glob, procs, rank, overlap);
  }
  static OverlappingPartition strong(const std::array<int, dim>& global_cells, int world, int rank, int overlap = 1) {
    return OverlappingPartition(global_cells, processor_grid(world, dim), rank, overlap);
  }

  int rank_of(const std::array<int, dim>& c) const {
    int r = 0, stride = 1;
    for (int d = 0; d < dim; d++) {
      r += stride * c[d];
      stride *= procs[d];
    }
    return r;
  }
  // (direction, side, neighbour rank) for every processor side of this rank
  std::vector<std::tuple<int, int, int>> exchanges() const {
    std::vector<std::tuple<int, int, int>> e;
    for (int d = 0; d < dim; d++)
      for (int s = 0; s < 2; s++)
        if (neighbour[d][s] >= 0) e.emplace_back(d, s, neighbour[d][s]);
    return e;
  }
  long long num_local_cells() const {
    long long n = 1;
    for (int d = 0; d < dim; d++) n *= local_cells[d];
    return n;
  }
  // local cell index (lexicographic over the extended box) -> global lexicographic cell index / owned?
  long long global_cell(long long local) const {
    long long g = 0, stride = 1;
    for (int d = 0; d < dim; d++) {
      const long long c = local % local_cells[d] + local_lo[d];
      local /= local_cells[d];
      g += stride * c;
      stride *= global_cells[d];
    }
    return g;
  }
  bool is_owned(long long local) const {
    for (int d = 0; d < dim; d++) {
      const long long c = local % local_cells[d] + local_lo[d];
      local /= local_cells[d];
      if (c < owned_lo[d] || c >= owned_hi[d]) return false;
    }
    return true;
  }
  // the rank's YaspGrid: local box with its processor sides marked (what GridOperator::init hands to pdb200_create)
  YaspGrid<dim> localGrid() const {
    YaspGrid<dim> g(local_lower, local_upper, local_cells);
    for (int d = 0; d < dim; d++) g.side_kind[d] = side_kind[d];
    return g;
  }

  std::array<int, dim> global_cells;
  std::array<int, 3> procs;
  int rank, overlap, world;
  std::array<int, dim> coords{}, owned_lo{}, owned_hi{}, local_lo{}, local_hi{}, owned_cells{}, local_cells{};
  std::array<std::array<int, 2>, 3> side_kind{}, neighbour{};
  FieldVector<double, dim> local_lower{}, local_upper{};
};

// the one collective of the set-up: every rank contributes 64 bytes, every rank receives all of them in rank order
// (MPI_Allgather(mine, 64, MPI_BYTE, all, 64, MPI_BYTE, comm) in a DUNE program)
using AllGatherHandles = std::function<std::vector<pdb200_ipc_handle>(const pdb200_ipc_handle& mine)>;

// Owner -> ghost copy through peer-mapped mailboxes (csrc/halo.cu): replaces the CopyDataHandle communication.
template <class GO, int dim>
class P2PHaloExchanger {
 public:
  P2PHaloExchanger(const GO& go, const OverlappingPartition<dim>& part, const AllGatherHandles& allgather) : go_(go) {
    pdb200_ipc_handle mine;
    check(pdb200_halo_p2p_create(go.handle(), &mine), "P2PHaloExchanger");
    const auto all = allgather(mine);
    if ((int)all.size() != part.world) throw Exception("P2PHaloExchanger: allgather must return one handle per rank");
    for (const auto& e : part.exchanges())
      check(pdb200_halo_p2p_connect(go.handle(), std::get<0>(e), std::get<1>(e), &all[std::get<2>(e)]), "P2PHaloExchanger");
  }
  // device vectors over the local box
  void exchange(double* x) const { check(pdb200_halo_exchange_p2p(go_.handle(), x), "halo exchange"); }
  // y = J x with the exchange hidden behind the interior tiles (OnTheFlyOperator::apply on the partition)
  void apply(double* x, double* y) const { check(pdb200_onthefly_apply_p2p(go_.handle(), x, y), "apply"); }

 protected:
  const GO& go_;
};

// Krylov solvers on the overlapping partition: the role of ISTLBackend_OVLP_{BCGS,CG}_* (ovlpistlsolverbackend.hh:477-560),
// matrix-free or with the rank's assembled matrix, device-resident (pdb200_solve_ovlp).
template <class GO, int dim, int SOLVER = PDB200_SOLVER_BICGSTAB, int PRECOND = PDB200_PRECOND_NONE>
class OverlappingSolverBackend : public P2PHaloExchanger<GO, dim> {
 public:
  OverlappingSolverBackend(const GO& go, const OverlappingPartition<dim>& part, const AllGatherHandles& allgather,
                           unsigned maxiter = 5000, int verbose = 1)
      : P2PHaloExchanger<GO, dim>(go, part, allgather), maxiter_(maxiter), verbose_(part.rank == 0 ? verbose : 0) {
    pdb200_ipc_handle mine;
    check(pdb200_comm_create(go.handle(), part.rank, part.world, &mine), "OverlappingSolverBackend");
    const auto all = allgather(mine);
    if ((int)all.size() != part.world) throw Exception("OverlappingSolverBackend: allgather must return one handle per rank");
    for (int r = 0; r < part.world; r++)
      if (r != part.rank) check(pdb200_comm_connect(go.handle(), r, &all[r]), "OverlappingSolverBackend");
  }
  // apply(z, r, reduction): matrix-free; apply(values, z, r, reduction): the rank's assembled CSR values (device pointers)
  void apply(double* z, double* r, double reduction) { run(nullptr, z, r, reduction); }
  void apply(const double* values, double* z, double* r, double reduction) { run(values, z, r, reduction); }
  // OverlappingScalarProduct::norm of a device vector in the unique representation is part of the solve; for host-side
  // scalars: comm().sum
  double sum(double v) const {
    check(pdb200_comm_sum(this->go_.handle(), &v, 1), "comm().sum");
    return v;
  }
  const LinearSolverResult<double>& result() const { return res_; }

 private:
  void run(const double* values, double* z, double* r, double reduction) {
    pdb200_solve_result s;
    check(pdb200_solve_ovlp(this->go_.handle(), SOLVER, PRECOND, values, PDB200_LAYOUT_CSR, z, r, reduction, maxiter_, &s),
          "ISTLBackend_OVLP::apply");
    res_.converged = s.converged != 0;
    res_.iterations = s.iterations;
    res_.elapsed = s.elapsed;
    res_.reduction = s.reduction;
    res_.conv_rate = s.conv_rate;
    if (verbose_ > 0)
      std::printf("=== overlapping device Krylov: %u iterations, reduction %.3e, %.4f s\n", s.iterations, s.reduction, s.elapsed);
  }
  unsigned maxiter_;
  int verbose_;
  LinearSolverResult<double> res_;
};

}  // namespace B200
}  // namespace PDELab
}  // namespace Dune

#endif  // PDELAB_B200_HOST_PARTITION_HH

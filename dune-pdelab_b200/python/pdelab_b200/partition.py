"""Overlapping Cartesian partition of a YaspGrid and the ghost-layer exchange for QkDG vectors.

Host-side logic only (index arithmetic + torch.distributed plumbing); the data movement on the
device is done by the pack/unpack kernels of the C ABI (csrc/halo.cu).

What it mirrors in the reference: an overlapping YaspGrid (overlap >= 1) splits the cell index
box into a Cartesian processor grid; every rank stores its interior block plus `overlap` ghost
cell layers towards each neighbour and the GridOperator assembles on interior + overlap cells
without any communication (gridoperator/default/assembler.hh:116).  Consistency of
the input vector is restored before the operator by an owner -> ghost copy
(Dune::PDELab::CopyDataHandle over InteriorBorder_All_Interface, boilerplate/pdelab.hh:872-880,
gridfunctionspace/genericdatahandle.hh); rows of cells touching the processor boundary are
zeroed through P0ParallelConstraints (constraints/p0.hh:31-41).  ConvectionDiffusionDG couples
cells through faces only, so the six face neighbours suffice.
"""
import numpy as np

from . import abi


def processor_grid(world, dim=3, split_x=False):
    """Cartesian processor grid, as cubic as possible, larger factors last.

    By default direction 0 is not split (2 -> 1x1x2, 4 -> 1x2x2, 8 -> 1x2x4): the x-rows of a
    QkDG vector stay contiguous and the local cell count in x stays even, which the TMA tiling of
    the fast kernel needs (csrc/dg_fast.cu); for cubic per-rank boxes every face costs the same
    message size, and a rank still has at most 3 neighbours at 8 ranks, exactly like 2x2x2.
    split_x=True gives YaspGrid's default most-cubic grid (2x2x2 at 8 ranks)."""
    ndim = dim if split_x else dim - 1
    best = None

    def rec(rem, k, cur):
        nonlocal best
        if k == 1:
            cand = sorted(cur + [rem])
            score = (max(cand) - min(cand), cand)
            if best is None or score < best[0]:
                best = (score, cand)
            return
        for f in range(1, rem + 1):
            if rem % f == 0:
                rec(rem // f, k - 1, cur + [f])

    rec(world, ndim, [])
    grid = best[1]
    return tuple(grid) if split_x else (1,) + tuple(grid)


class OverlappingPartition:
    """One rank's view of the overlapping partition."""

    def __init__(self, global_cells, procs, rank, overlap=1, lower=None, upper=None):
        self.dim = len(global_cells)
        self.global_cells = tuple(int(v) for v in global_cells)
        self.procs = tuple(int(v) for v in procs)
        self.world = int(np.prod(self.procs))
        self.rank = int(rank)
        self.overlap = int(overlap)
        lower = tuple(lower) if lower is not None else (0.0,) * self.dim
        upper = tuple(upper) if upper is not None else (1.0,) * self.dim
        assert 0 <= rank < self.world
        # rank = px + Px*(py + Py*pz)  (lexicographic torus coordinates, x fastest)
        c, r = [], self.rank
        for d in range(self.dim):
            c.append(r % self.procs[d])
            r //= self.procs[d]
        self.coords = tuple(c)
        self.owned_lo, self.owned_hi, self.local_lo, self.local_hi = [], [], [], []
        self.side_kind = [[abi.SIDE_DOMAIN, abi.SIDE_DOMAIN] for _ in range(3)]
        self.neighbour = [[None, None] for _ in range(3)]
        for d in range(self.dim):
            n, p, i = self.global_cells[d], self.procs[d], self.coords[d]
            lo, hi = (n * i) // p, (n * (i + 1)) // p
            self.owned_lo.append(lo)
            self.owned_hi.append(hi)
            llo, lhi = lo, hi
            if i > 0:
                llo -= self.overlap
                self.side_kind[d][0] = abi.SIDE_PROCESSOR
                self.neighbour[d][0] = self.rank_of(tuple(cc - (1 if dd == d else 0) for dd, cc in enumerate(self.coords)))
            if i < p - 1:
                lhi += self.overlap
                self.side_kind[d][1] = abi.SIDE_PROCESSOR
                self.neighbour[d][1] = self.rank_of(tuple(cc + (1 if dd == d else 0) for dd, cc in enumerate(self.coords)))
            self.local_lo.append(llo)
            self.local_hi.append(lhi)
        self.owned_cells = tuple(h - l for l, h in zip(self.owned_lo, self.owned_hi))
        self.local_cells = tuple(h - l for l, h in zip(self.local_lo, self.local_hi))
        hs = [(upper[d] - lower[d]) / self.global_cells[d] for d in range(self.dim)]
        self.local_lower = tuple(lower[d] + hs[d] * self.local_lo[d] for d in range(self.dim))
        self.local_upper = tuple(lower[d] + hs[d] * self.local_hi[d] for d in range(self.dim))

    @classmethod
    def weak(cls, cells_per_rank, world, rank, overlap=1):
        procs = processor_grid(world, len(cells_per_rank))
        glob = tuple(c * p for c, p in zip(cells_per_rank, procs))
        return cls(glob, procs, rank, overlap)

    @classmethod
    def strong(cls, global_cells, world, rank, overlap=1):
        return cls(global_cells, processor_grid(world, len(global_cells)), rank, overlap)

    def rank_of(self, coords):
        r, stride = 0, 1
        for d in range(self.dim):
            r += stride * coords[d]
            stride *= self.procs[d]
        return r

    # index helpers -------------------------------------------------------------------------
    def local_cell_grid(self):
        """Global lexicographic cell index of every local cell, shape local_cells[::-1]."""
        axes = [np.arange(self.local_lo[d], self.local_hi[d]) for d in range(self.dim)]
        idx = np.zeros(self.local_cells[::-1], dtype=np.int64)
        stride = 1
        for d in range(self.dim):
            shape = [1] * self.dim
            shape[self.dim - 1 - d] = -1
            idx = idx + stride * axes[d].reshape(shape)
            stride *= self.global_cells[d]
        return idx

    def owned_mask(self):
        """Boolean array over local cells (shape local_cells[::-1]): owned (interior) cells."""
        m = np.ones(self.local_cells[::-1], dtype=bool)
        for d in range(self.dim):
            ax = self.dim - 1 - d
            sel = np.zeros(self.local_cells[d], dtype=bool)
            sel[self.owned_lo[d] - self.local_lo[d]: self.owned_hi[d] - self.local_lo[d]] = True
            shape = [1] * self.dim
            shape[ax] = -1
            m &= sel.reshape(shape)
        return m

    def exchanges(self):
        """(direction, side, neighbour rank) for every processor side of this rank."""
        return [(d, s, self.neighbour[d][s]) for d in range(self.dim) for s in range(2)
                if self.neighbour[d][s] is not None]


class HaloExchanger:
    """Owner -> ghost copy of a QkDG vector across the six face neighbours (overlap = 1).

    pack/unpack default to the CUDA kernels behind the C ABI (pdb200_halo_pack / _unpack); the
    transport is torch.distributed point-to-point (NCCL over NVLink on the GPU box, gloo in the
    CPU tests, which inject their own pack/unpack).
    """

    def __init__(self, go, part, device, pack=None, unpack=None, layer_size=None, dist=None):
        import torch
        if dist is None:
            import torch.distributed as dist
        assert part.overlap == 1, "the pack/unpack kernels move one cell layer"
        self.dist, self.part, self.torch = dist, part, torch
        self.pack = pack or go.halo_pack
        self.unpack = unpack or go.halo_unpack
        size = layer_size or go.halo_layer_size
        self.send, self.recv = {}, {}
        for d, s, _ in part.exchanges():
            n = size(d)
            self.send[(d, s)] = torch.empty(n, dtype=torch.float64, device=device)
            self.recv[(d, s)] = torch.empty(n, dtype=torch.float64, device=device)
        self.bytes_per_exchange = sum(t.numel() * 8 for t in self.send.values())

    def exchange(self, x):
        """Fill the ghost layers of x with the neighbours' owned values."""
        dist = self.dist
        ops = []
        for d, s, nbr in self.part.exchanges():
            self.pack(x, d, s, self.send[(d, s)])
            ops.append(dist.P2POp(dist.isend, self.send[(d, s)], nbr))
            ops.append(dist.P2POp(dist.irecv, self.recv[(d, s)], nbr))
        if not ops:
            return
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for d, s, _ in self.part.exchanges():
            self.unpack(x, d, s, self.recv[(d, s)])


def qk_container_index(cells, k, P):
    """Container index of the conforming Qk DOF at lattice point(s) P (int array [..., dim],
    0 <= P_d <= k * cells_d): the closed form of csrc/host_tables.h (QkLayout, qk_lattice_index) —
    Q1: vertex-lexicographic; Q2: vertices | edges | faces | cells blocks, sub-entities grouped by
    extension bitset (SURVEY.md §8a "index-formula detail")."""
    cells = tuple(int(c) for c in cells)
    dim = len(cells)
    P = np.asarray(P, dtype=np.int64)
    if k == 1:
        idx, stride = np.zeros(P.shape[:-1], dtype=np.int64), 1
        for d in range(dim):
            idx = idx + stride * P[..., d]
            stride *= cells[d] + 1
        return idx
    assert k == 2
    size = lambda s: int(np.prod([cells[d] if (s >> d) & 1 else cells[d] + 1 for d in range(dim)]))
    count = [0] * (dim + 1)
    group_off = [0] * (1 << dim)
    for edim in range(dim + 1):
        for s in range(1 << dim):
            if bin(s).count("1") == edim:
                group_off[s] = count[edim]
                count[edim] += size(s)
    block_off = np.concatenate([[0], np.cumsum(count)])[:dim + 1]
    ext = P & 1
    s = np.zeros(P.shape[:-1], dtype=np.int64)
    for d in range(dim):
        s |= ext[..., d] << d
    edim = ext.sum(axis=-1)
    idx, stride = np.zeros(P.shape[:-1], dtype=np.int64), np.ones(P.shape[:-1], dtype=np.int64)
    for d in range(dim):
        idx = idx + stride * (P[..., d] >> 1)
        stride = stride * np.where(ext[..., d] == 1, cells[d], cells[d] + 1)
    return block_off[edim] + np.asarray(group_off, dtype=np.int64)[s] + idx


class QkHaloExchanger:
    """Owner -> ghost copy of a conforming Qk vector on the overlapping partition (overlap = 1).

    Every lattice point has one owner: the points of the interface plane between two ranks' owned
    cells belong to the LOWER rank (the rule of genericdatahandle.hh:894-947 restricted to a
    Cartesian partition).  Towards its upper neighbour a rank sends the k+1 lattice planes of its
    last owned cell layer (incl. the interface) and receives the k planes beyond the interface;
    towards its lower neighbour it sends k planes and receives k+1.  Directions are exchanged one
    after the other over the full tangential extent, so edge and corner neighbours are reached
    without extra messages.  After the exchange the vector is consistent on the whole extended box;
    the operator then computes complete rows for every point of the closure of the owned cells (the
    boundary of the extended box is constrained like a Dirichlet boundary).

    gather/scatter default to the index kernels behind the C ABI (pdb200_gather_dofs / _scatter_dofs);
    the transport is torch.distributed point-to-point."""

    def __init__(self, go, part, degree, device, gather=None, scatter=None, dist=None):
        import torch
        if dist is None:
            import torch.distributed as dist
        assert part.overlap == 1
        self.dist, self.part, self.k = dist, part, int(degree)
        self.gather = gather or go.gather_dofs
        self.scatter = scatter or go.scatter_dofs
        k, lc = self.k, part.local_cells
        self.plan = []   # per direction: list of (neighbour, send idx, recv idx, send buf, recv buf)
        for d in range(part.dim):
            entries = []
            for s in range(2):
                nbr = part.neighbour[d][s]
                if nbr is None:
                    continue
                n = lc[d]
                if s == 1:   # upper neighbour: I own the interface plane
                    send_r, recv_r = (k * (n - 2), k * (n - 1)), (k * (n - 1) + 1, k * n)
                else:        # lower neighbour owns the interface plane
                    send_r, recv_r = (k + 1, 2 * k), (0, k)
                mk = lambda r: torch.from_numpy(self._box_indices(d, r)).to(device)
                si, ri = mk(send_r), mk(recv_r)
                entries.append((nbr, si, ri, torch.empty(si.numel(), dtype=torch.float64, device=device),
                                torch.empty(ri.numel(), dtype=torch.float64, device=device)))
            if entries:
                self.plan.append(entries)
        self.bytes_per_exchange = sum(e[3].numel() * 8 for entries in self.plan for e in entries)

    def _box_indices(self, d, rng):
        """container indices of the lattice points with rng[0] <= p_d <= rng[1], full extent elsewhere"""
        part, k = self.part, self.k
        axes = [np.arange(rng[0], rng[1] + 1) if dd == d else np.arange(k * part.local_cells[dd] + 1)
                for dd in range(part.dim)]
        grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
        P = np.stack([g.reshape(-1) for g in grids], axis=-1)
        return qk_container_index(part.local_cells, k, P).astype(np.int64)

    def owned_point_mask(self):
        """Boolean mask over the local container: lattice points this rank OWNS (each global point
        exactly once over all ranks): points of the closure of the owned cells minus the interface
        planes towards lower neighbours."""
        part, k = self.part, self.k
        axes = [np.arange(k * part.local_cells[dd] + 1) for dd in range(part.dim)]
        grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
        P = np.stack([g.reshape(-1) for g in grids], axis=-1)
        own = np.ones(P.shape[0], dtype=bool)
        for d in range(part.dim):
            lo = k + 1 if part.neighbour[d][0] is not None else 0
            hi = k * (part.local_cells[d] - 1) if part.neighbour[d][1] is not None else k * part.local_cells[d]
            own &= (P[:, d] >= lo) & (P[:, d] <= hi)
        mask = np.zeros(P.shape[0], dtype=bool)
        mask[qk_container_index(part.local_cells, k, P)] = own
        return mask

    def global_point_index(self, global_cells):
        """global container index of every local DOF (array over the local container)"""
        part, k = self.part, self.k
        axes = [np.arange(k * part.local_cells[dd] + 1) for dd in range(part.dim)]
        grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
        P = np.stack([g.reshape(-1) for g in grids], axis=-1)
        G = P + k * np.asarray(part.local_lo, dtype=np.int64)
        out = np.zeros(P.shape[0], dtype=np.int64)
        out[qk_container_index(part.local_cells, k, P)] = qk_container_index(global_cells, k, G)
        return out

    def exchange(self, x):
        dist = self.dist
        for entries in self.plan:      # one direction after the other: corners travel in two hops
            ops = []
            for nbr, si, ri, sb, rb in entries:
                self.gather(x, si, sb)
                ops.append(dist.P2POp(dist.isend, sb, nbr))
                ops.append(dist.P2POp(dist.irecv, rb, nbr))
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            for nbr, si, ri, sb, rb in entries:
                self.scatter(rb, ri, x)


def exchange_cell_field(field, part, dist):
    """Owner -> ghost copy of a per-cell field (torch tensor of shape local_cells[::-1] + trailing
    dims), used once at set-up for the coefficient arrays.  Plain tensor slicing: not a hot path."""
    import torch
    # one direction after the other over the full tangential extent: edge and corner ghost cells (needed
    # by the conforming spaces) are filled in two / three hops
    for d in range(part.dim):
        ops, recvs = [], []
        for s in range(2):
            nbr = part.neighbour[d][s]
            if nbr is None:
                continue
            ax = part.dim - 1 - d
            n = part.local_cells[d]
            src = 1 if s == 0 else n - 2
            dst = 0 if s == 0 else n - 1
            sendbuf = field.select(ax, src).contiguous()
            recvbuf = torch.empty_like(sendbuf)
            ops.append(dist.P2POp(dist.isend, sendbuf, nbr))
            ops.append(dist.P2POp(dist.irecv, recvbuf, nbr))
            recvs.append((ax, dst, recvbuf))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for ax, dst, buf in recvs:
            field.select(ax, dst).copy_(buf)
    return field


class P2PHaloExchanger:
    """Owner -> ghost copy through peer-mapped mailboxes (csrc/halo.cu): no NCCL call and no host
    round trip per exchange.  torch.distributed is used ONCE, at set-up, to hand every rank's IPC
    handle to its face neighbours (any backend: nccl on the GPU box, gloo in tests)."""

    def __init__(self, go, part, dist=None):
        if dist is None:
            import torch.distributed as dist
        self.go, self.part = go, part
        mine = go.halo_p2p_create()
        handles = [None] * part.world
        dist.all_gather_object(handles, mine)
        for d, s, nbr in part.exchanges():
            go.halo_p2p_connect(d, s, handles[nbr])
        dist.barrier()  # every rank has mapped its neighbours before the first push

    def exchange(self, x):
        self.go.halo_exchange_p2p(x)

    def apply(self, x, y):
        """y = J x with the exchange overlapped with the interior tiles."""
        return self.go.apply_p2p(x, y)


class OverlappingSolverBackend(P2PHaloExchanger):
    """Krylov solvers on the overlapping partition — the role of the reference's ISTLBackend_OVLP_* classes
    (backend/istl/ovlpistlsolverbackend.hh:477-560: OverlappingOperator + OverlappingScalarProduct + BiCGSTAB / CG)
    with one process per GPU.  On top of the halo mailboxes every rank maps every other rank's reduction mailbox
    (one all_gather at set-up); per solve there is no torch.distributed / NCCL call at all."""

    def __init__(self, go, part, dist=None, solver=None, precond=None, maxiter=5000):
        from . import abi
        if dist is None:
            import torch.distributed as dist
        super().__init__(go, part, dist)
        mine = go.comm_create(part.rank, part.world)
        handles = [None] * part.world
        dist.all_gather_object(handles, mine)
        for r in range(part.world):
            if r != part.rank:
                go.comm_connect(r, handles[r])
        dist.barrier()
        self.solver = abi.SOLVER_BICGSTAB if solver is None else solver
        self.precond = abi.PRECOND_NONE if precond is None else precond
        self.maxiter = maxiter
        self.result = None

    def apply(self, *args):
        """apply(z, r, reduction): matrix-free;  apply(values, z, r, reduction): the rank's assembled matrix
        (the call signatures of the reference's back-ends, ovlpistlsolverbackend.hh:520-545)."""
        if len(args) == 3:
            z, r, reduction = args
            values = None
        else:
            values, z, r, reduction = args
        self.result = self.go.solve_ovlp(z, r, reduction, solver=self.solver, precond=self.precond, values=values,
                                         maxiter=self.maxiter)
        return self.result

    def norm(self, v):
        """OverlappingScalarProduct::norm of a vector in the unique representation (ghost rows zero)."""
        import numpy as np
        own = float((v * v).sum())
        return float(np.sqrt(self.go.comm_sum(np.array([own]))[0]))

"""Overlapping partition + ghost-layer exchange: host logic on CPU with gloo, world_size 2 and 4.

The distributed result on owned cells must equal the single-domain oracle on the global grid
(SURVEY.md §8e "Parity definition").  On CPU the pack/unpack kernels are replaced by torch
indexing injected into HaloExchanger; the operator is the CPU oracle (this is a test of the
partition/exchange logic, not of the CUDA path — tests/test_gpu_multi.py covers that)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdelab_b200 import abi
from pdelab_b200.partition import HaloExchanger, OverlappingPartition, exchange_cell_field, processor_grid


def test_processor_grid():
    assert processor_grid(1) == (1, 1, 1)
    assert processor_grid(2) == (1, 1, 2)
    assert processor_grid(4) == (1, 2, 2)
    assert processor_grid(8) == (1, 2, 4)
    assert processor_grid(8, split_x=True) == (2, 2, 2)
    assert processor_grid(4, dim=2, split_x=True) == (2, 2)


@pytest.mark.parametrize("cells,procs", [((8, 6, 4), (1, 2, 2)), ((6, 6, 6), (2, 2, 2)), ((7, 5), (2, 2)), ((4, 4, 9), (1, 1, 3))])
def test_owned_cells_tile_the_grid(cells, procs):
    world = int(np.prod(procs))
    seen = np.zeros(cells[::-1], dtype=int)
    for r in range(world):
        p = OverlappingPartition(cells, procs, r)
        gidx = p.local_cell_grid()
        own = p.owned_mask()
        np.add.at(seen.reshape(-1), gidx[own], 1)
        # ghost layers only towards neighbours; the processor sides are flagged
        for d in range(len(cells)):
            for s in range(2):
                has_nbr = p.neighbour[d][s] is not None
                assert (p.side_kind[d][s] == abi.SIDE_PROCESSOR) == has_nbr
                if has_nbr:
                    q = OverlappingPartition(cells, procs, p.neighbour[d][s])
                    assert q.neighbour[d][1 - s] == r
        assert tuple(h - l for l, h in zip(p.local_lo, p.local_hi)) == p.local_cells
    assert np.all(seen == 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _layer_index(part, n, d, layer):
    """flat DOF indices of cell layer `layer` normal to d, lexicographic tangential order"""
    lc = part.local_cells
    axes = [np.arange(lc[dd]) if dd != d else np.array([layer]) for dd in range(part.dim)]
    grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
    cell = np.zeros_like(grids[0])
    stride = 1
    for dd in range(part.dim):
        cell = cell + stride * grids[dd]
        stride *= lc[dd]
    return (cell.reshape(-1, 1) * n + np.arange(n)).reshape(-1)


def _worker(rank, world, port, cells, degree, out):
    sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle"), os.path.dirname(__file__)]
    from oracle import Oracle
    from problems import kappa_field, mt_vector
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim = len(cells)
        part = OverlappingPartition.strong(cells, world, rank) if dim == 3 else \
            OverlappingPartition(cells, processor_grid(world, 2, split_x=True), rank)
        n = (degree + 1) ** dim
        ncg = int(np.prod(cells))
        zg = mt_vector(ncg * n).reshape(ncg, n)
        kg = kappa_field(ncg)
        gidx = part.local_cell_grid().reshape(-1)
        own = part.owned_mask().reshape(-1)
        # owned values from the global vector, ghosts poisoned: the exchange must fill them
        z = torch.full((gidx.size, n), float("nan"), dtype=torch.float64)
        z[own] = torch.from_numpy(zg[gidx[own]])
        z = z.reshape(-1)
        kap = torch.full((gidx.size,), float("nan"), dtype=torch.float64)
        kap[own] = torch.from_numpy(kg[gidx[own]])
        exchange_cell_field(kap.view(part.local_cells[::-1]), part, dist)

        def pack(x, d, s, buf):
            layer = 1 if s == 0 else part.local_cells[d] - 2
            buf.copy_(x[torch.from_numpy(_layer_index(part, n, d, layer))])

        def unpack(x, d, s, buf):
            layer = 0 if s == 0 else part.local_cells[d] - 1
            x[torch.from_numpy(_layer_index(part, n, d, layer))] = buf

        halo = HaloExchanger(None, part, "cpu", pack=pack, unpack=unpack,
                             layer_size=lambda d: int(np.prod(part.local_cells)) // part.local_cells[d] * n, dist=dist)
        halo.exchange(z)
        # face neighbours only: corner/edge ghosts may stay NaN in the vector but are never read by owned rows
        zl = torch.nan_to_num(z, nan=1e300).numpy()
        kl = torch.nan_to_num(kap, nan=1.0).numpy()
        spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QKDG, degree=degree, lower=part.local_lower,
                               upper=part.local_upper, alpha=3.0, a_mode=abi.A_SCALAR, A=kl, side_kind=part.side_kind)
        y = Oracle(spec).jacobian_apply(zl).reshape(-1, n)
        # rows of cells touching a processor side are zero (P0ParallelConstraints): those are ghosts
        gspec = abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=degree, alpha=3.0, a_mode=abi.A_SCALAR, A=kg)
        want = Oracle(gspec).jacobian_apply(zg.reshape(-1)).reshape(-1, n)
        err = np.abs(y[own] - want[gidx[own]]).max() / np.abs(want).max()
        ghosts_zero = bool(np.all(y[~own] == 0.0)) if (~own).any() else True
        out.put((rank, float(err), ghosts_zero, int(own.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cells,degree", [(2, (4, 3, 6), 2), (4, (4, 6, 6), 1), (4, (6, 6), 2)])
def test_distributed_apply_matches_global_oracle(world, cells, degree):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cells, degree, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[3] for r in res) == int(np.prod(cells))
    for rank, err, ghosts_zero, _ in res:
        assert err < 1e-12, (rank, err)
        assert ghosts_zero


# ---- conforming Qk on the overlapping partition (QkHaloExchanger) -----------------------------------

def test_qk_container_index_matches_the_library_numbering():
    """partition.qk_container_index restates csrc/host_tables.h; the oracle restates it too
    (cell_dof_indices): the two must agree DOF by DOF."""
    sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle")]
    from oracle import Oracle
    from pdelab_b200.partition import qk_container_index
    for cells, k in (((3, 2), 1), ((3, 2), 2), ((2, 3, 2), 1), ((3, 2, 2), 2)):
        spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=k)
        orc = Oracle(spec)
        dim, n1 = len(cells), k + 1
        for e in range(spec.ncells):
            c = np.unravel_index(e, cells[::-1])[::-1]
            loc = np.array(np.unravel_index(np.arange(n1 ** dim), (n1,) * dim)[::-1]).T
            P = loc + k * np.asarray(c)
            assert np.array_equal(qk_container_index(cells, k, P), orc.cell_dof_indices(e).astype(np.int64))


def _qk_worker(rank, world, port, cells, degree, out):
    sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle"), os.path.dirname(__file__)]
    from oracle import Oracle
    from problems import kappa_field, mt_vector
    from pdelab_b200.partition import QkHaloExchanger
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim = len(cells)
        part = OverlappingPartition.strong(cells, world, rank) if dim == 3 else \
            OverlappingPartition(cells, processor_grid(world, 2, split_x=True), rank)
        ncg = int(np.prod(cells))
        kg = kappa_field(ncg)
        gspec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=degree, a_mode=abi.A_SCALAR, A=kg)
        zg = mt_vector(gspec.num_dofs)
        want = Oracle(gspec).jacobian_apply(zg)
        gidx = part.local_cell_grid().reshape(-1)
        own_c = part.owned_mask().reshape(-1)
        kap = torch.full((gidx.size,), float("nan"), dtype=torch.float64)
        kap[own_c] = torch.from_numpy(kg[gidx[own_c]])
        exchange_cell_field(kap.view(part.local_cells[::-1]), part, dist)

        def gather(x, idx, buf):
            buf.copy_(x[idx])

        def scatter(buf, idx, x):
            x[idx] = buf

        halo = QkHaloExchanger(None, part, degree, "cpu", gather=gather, scatter=scatter, dist=dist)
        own = halo.owned_point_mask()
        gp = halo.global_point_index(cells)
        # owned values from the global vector, everything else poisoned: the exchange must fill the box
        z = torch.full((own.size,), float("nan"), dtype=torch.float64)
        z[torch.from_numpy(own)] = torch.from_numpy(zg[gp[own]])
        halo.exchange(z)
        consistent = bool(np.array_equal(z.numpy(), zg[gp]))
        kl = torch.nan_to_num(kap, nan=1.0).numpy()
        spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QK, degree=degree, lower=part.local_lower,
                               upper=part.local_upper, a_mode=abi.A_SCALAR, A=kl, side_kind=part.side_kind)
        y = Oracle(spec).jacobian_apply(np.nan_to_num(z.numpy(), nan=1e300))
        err = np.abs(y[own] - want[gp[own]]).max() / np.abs(want).max()
        out.put((rank, float(err), consistent, int(own.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cells,degree", [(2, (4, 3, 6), 2), (4, (3, 6, 6), 1), (4, (6, 6), 2), (4, (5, 7), 1)])
def test_distributed_qk_apply_matches_global_oracle(world, cells, degree):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_qk_worker, args=(r, world, port, cells, degree, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ndofs = int(np.prod([degree * c + 1 for c in cells]))
    assert sum(r[3] for r in res) == ndofs          # every lattice point has exactly one owner
    for rank, err, consistent, _ in res:
        assert consistent, rank                      # the whole extended box carries the global values
        assert err < 1e-12, (rank, err)


# ---- overlapping Krylov solver: the scheme of pdb200_solve_ovlp restated on the CPU (gloo) --------------------
# Vectors are ghost-free outside the operator: the apply makes its input consistent (halo exchange), evaluates the
# local rows (ghost rows come out zero), and drops the input's ghosts again; every inner product is then the disjoint
# dot product of OverlappingScalarProduct (backend/istl/ovlpistlsolverbackend.hh:103-108) plus one all_reduce.

class _FakeGridOperator:
    """Records the mailbox handshake of OverlappingSolverBackend (no device here)."""

    def __init__(self, rank):
        self.rank, self.calls = rank, []

    def halo_p2p_create(self):
        return b"H%03d" % self.rank + bytes(60)

    def halo_p2p_connect(self, d, s, handle):
        self.calls.append(("halo", d, s, bytes(handle[:4])))

    def comm_create(self, rank, size):
        self.calls.append(("comm_create", rank, size))
        return b"C%03d" % rank + bytes(60)

    def comm_connect(self, peer, handle):
        self.calls.append(("comm", peer, bytes(handle[:4])))


def _cg_worker(rank, world, port, cells, out):
    sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle"), os.path.dirname(__file__)]
    from oracle import Oracle
    from pdelab_b200.partition import OverlappingSolverBackend
    from problems import kappa_field, mt_vector
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        degree, n = 1, 8
        part = OverlappingPartition.strong(cells, world, rank)
        # the set-up handshake: halo mailboxes to the face neighbours, reduction mailboxes to every other rank
        fake = _FakeGridOperator(rank)
        OverlappingSolverBackend(fake, part, dist)
        want_calls = [("halo", d, s, b"H%03d" % nbr) for d, s, nbr in part.exchanges()]
        want_calls += [("comm_create", rank, world)] + [("comm", r, b"C%03d" % r) for r in range(world) if r != rank]
        handshake_ok = fake.calls == want_calls

        ncg = int(np.prod(cells))
        bg = mt_vector(ncg * n, seed=3).reshape(ncg, n)
        kg = kappa_field(ncg)
        gidx = part.local_cell_grid().reshape(-1)
        own = part.owned_mask().reshape(-1)
        spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QKDG, degree=degree, lower=part.local_lower,
                               upper=part.local_upper, alpha=3.0, a_mode=abi.A_SCALAR, A=kg[gidx], side_kind=part.side_kind)
        orc = Oracle(spec)

        def pack(x, d, s, buf):
            layer = 1 if s == 0 else part.local_cells[d] - 2
            buf.copy_(x[torch.from_numpy(_layer_index(part, n, d, layer))])

        def unpack(x, d, s, buf):
            layer = 0 if s == 0 else part.local_cells[d] - 1
            x[torch.from_numpy(_layer_index(part, n, d, layer))] = buf

        halo = HaloExchanger(None, part, "cpu", pack=pack, unpack=unpack,
                             layer_size=lambda d: int(np.prod(part.local_cells)) // part.local_cells[d] * n, dist=dist)
        ghost = np.repeat(~own, n)

        def apply(v):                      # OverlappingOperator::apply on a ghost-free vector
            t = torch.from_numpy(v.copy())
            halo.exchange(t)
            y = orc.jacobian_apply(t.numpy())
            assert np.all(y[ghost] == 0.0)
            return y

        def dot(a, b):                     # OverlappingScalarProduct::dot
            s = torch.tensor([float(a @ b)], dtype=torch.float64)
            dist.all_reduce(s)
            return float(s)

        b = np.zeros((gidx.size, n))
        b[own] = bg[gidx[own]]
        b = b.reshape(-1)
        x = np.zeros_like(b)
        r = b - apply(x)
        p = r.copy()
        rr = dot(r, r)
        r0 = np.sqrt(rr)
        its = 0
        while np.sqrt(rr) > 1e-10 * r0 and its < 500:
            q = apply(p)
            lam = rr / dot(p, q)
            x += lam * p
            r -= lam * q
            rr_new = dot(r, r)
            p = r + (rr_new / rr) * p
            rr = rr_new
            its += 1
        assert np.all(x[ghost] == 0.0) and np.all(r[ghost] == 0.0) and np.all(p[ghost] == 0.0)
        out.put((rank, its, gidx[own], x.reshape(-1, n)[own], handshake_ok))
    finally:
        dist.destroy_process_group()


def test_overlapping_cg_scheme_reproduces_the_undivided_solve():
    """world_size 2 over gloo: the ghost-free CG of pdb200_solve_ovlp is the CG of the undivided problem."""
    sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle")]
    from oracle import Oracle
    from problems import kappa_field, mt_vector
    world, cells, n = 2, (4, 3, 6), 8
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, cells, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ncg = int(np.prod(cells))
    xg = np.zeros((ncg, n))
    for _, _, gi, xi, ok in res:
        xg[gi] = xi
        assert ok
    assert len({r[1] for r in res}) == 1
    # the same CG on the undivided grid
    orc = Oracle(abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=1, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa_field(ncg)))
    b = mt_vector(ncg * n, seed=3)
    x = np.zeros_like(b)
    r = b.copy()
    p = r.copy()
    rr = r @ r
    r0 = np.sqrt(rr)
    its = 0
    while np.sqrt(rr) > 1e-10 * r0 and its < 500:
        q = orc.jacobian_apply(p)
        lam = rr / (p @ q)
        x += lam * p
        r -= lam * q
        rr_new = r @ r
        p = r + (rr_new / rr) * p
        rr = rr_new
        its += 1
    assert abs(res[0][1] - its) <= 1, (res[0][1], its)
    assert np.abs(xg.reshape(-1) - x).max() / np.abs(x).max() < 1e-8

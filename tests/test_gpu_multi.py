"""Overlapping partition on the CUDA path: interior/boundary split of the fast kernel and the
peer-to-peer (CUDA IPC mailbox) halo exchange, world_size 2 and 4.  All ranks share cuda:0 when the
box has fewer GPUs than ranks (IPC between processes works on one device, NCCL would not), the
set-up handshake runs over gloo.  Parity definition of SURVEY.md §8e: owned rows of the
distributed result equal the single-domain oracle on the global grid to 1e-12."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from pdelab_b200 import abi

pytestmark = pytest.mark.gpu


def test_interior_plus_boundary_equals_all(cuda_lib):
    from pdelab_b200.capi import GridOperator
    from problems import kappa_field
    cells = (12, 10, 13)
    nc = int(np.prod(cells))
    for side_kind in ([[0, 0], [0, 1], [1, 1]], [[0, 0], [1, 0], [0, 1]], [[0, 0], [0, 0], [0, 0]]):
        spec = abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kappa_field(nc), side_kind=side_kind)
        go = GridOperator(spec)
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.rand(spec.num_dofs, dtype=torch.float64, device="cuda", generator=g)
        y_all = torch.full_like(x, float("nan"))
        y_split = torch.full_like(x, float("nan"))
        go.apply(x, y_all)
        go.apply_part(x, y_split, abi.PART_INTERIOR)
        torch.cuda.synchronize()
        n_int = int((~torch.isnan(y_split)).sum())
        go.apply_part(x, y_split, abi.PART_BOUNDARY)
        torch.cuda.synchronize()
        assert go.last_kernel() == "dg_fast_q2_3d"
        assert torch.equal(y_all, y_split)            # same kernel, same tiles: bit-identical
        if any(any(r) for r in side_kind):
            assert 0 < n_int < spec.num_dofs          # the interior part is a strict, non-empty subset
        else:
            assert n_int == spec.num_dofs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cells, out):
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [os.path.join(here, "..", "oracle"), here, os.path.join(here, "..", "dune-pdelab_b200", "python")]
        import torch.distributed as dist
        from oracle import Oracle
        from pdelab_b200.capi import GridOperator
        from pdelab_b200.partition import OverlappingPartition, P2PHaloExchanger
        from problems import kappa_field, mt_vector
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ndev = torch.cuda.device_count()
        dev = rank % ndev
        torch.cuda.set_device(dev)
        part = OverlappingPartition.strong(cells, world, rank)
        n = 27
        ncg = int(np.prod(cells))
        zg = mt_vector(ncg * n).reshape(ncg, n)
        kg = kappa_field(ncg)
        gidx = part.local_cell_grid().reshape(-1)
        own = part.owned_mask().reshape(-1)
        z = np.full((gidx.size, n), 1e300)          # ghosts poisoned: the exchange must fill them
        z[own] = zg[gidx[own]]
        spec = abi.ProblemSpec(part.local_cells, degree=2, lower=part.local_lower, upper=part.local_upper, alpha=3.0,
                               a_mode=abi.A_SCALAR, A=kg[gidx], side_kind=part.side_kind, device=dev)
        go = GridOperator(spec)
        halo = P2PHaloExchanger(go, part, dist)
        zd = torch.from_numpy(z.reshape(-1)).cuda()
        yd = torch.full_like(zd, float("nan"))
        want = Oracle(abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kg)).jacobian_apply(
            zg.reshape(-1)).reshape(-1, n)
        errs = []
        for it in range(3):                          # several epochs: flags/acks must keep working
            halo.apply(zd, yd)
            go.synchronize()
            y = yd.cpu().numpy().reshape(-1, n)
            errs.append(float(np.abs(y[own] - want[gidx[own]]).max() / np.abs(want).max()))
        kern = go.last_kernel()
        fused = kern.endswith("+halo")               # one-launch step: x's ghost layers are neither read nor written
        lc = part.local_cells
        coords = np.unravel_index(np.arange(gidx.size), lc[::-1])   # (z, y, x)
        outside = np.zeros(gidx.size, dtype=int)
        for d in range(3):
            c = coords[2 - d] + part.local_lo[d]
            outside += ~((part.owned_lo[d] <= c) & (c < part.owned_hi[d]))
        face_ghost = outside == 1
        zl = zd.cpu().numpy().reshape(-1, n)
        if fused:
            ok_ghost = bool(face_ghost.any()) and bool(np.all(zl[face_ghost] == 1e300))
        else:                                        # multi-launch schedule: face ghosts hold the neighbour's owned values
            ok_ghost = bool(face_ghost.any()) and bool(np.array_equal(zl[face_ghost], zg[gidx[face_ghost]]))
        # plain exchange entry point too, alternating with the apply on the same mailboxes
        # (edge/corner ghosts carry whatever the neighbour's own ghosts held: never read by owned rows)
        zd2 = torch.from_numpy(z.reshape(-1)).cuda()
        halo.exchange(zd2)
        go.synchronize()
        same = bool(np.array_equal(zd2.cpu().numpy().reshape(-1, n)[face_ghost], zg[gidx[face_ghost]]))
        halo.apply(zd, yd)
        go.synchronize()
        y = yd.cpu().numpy().reshape(-1, n)
        errs.append(float(np.abs(y[own] - want[gidx[own]]).max() / np.abs(want).max()))
        out.put((rank, max(errs), ok_ghost and same, kern, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, 1.0, False, "", traceback.format_exc()))


# owned extents that are multiples of the 8x4x4 tile run the one-launch step ("+halo"), the others the multi-launch one
@pytest.mark.parametrize("world,cells", [(2, (8, 6, 12)), (4, (8, 12, 10)), (2, (8, 8, 16)), (4, (16, 16, 8)), (4, (8, 8, 24)),
                                         (3, (8, 4, 12))])
def test_p2p_halo_apply_matches_global_oracle(cuda_lib, world, cells):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cells, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    fused = set()
    for rank, err, ok, kern, tb in res:
        assert tb is None, tb
        assert err < 1e-12, (rank, err)
        assert ok, rank
        assert kern in ("dg_fast_q2_3d", "dg_fast_q2_3d+halo")
        fused.add(kern.endswith("+halo"))
    # (a rank with only a LOWER processor side needs no alignment, so the other cases mix both schedules: the flag
    # protocol on the mailboxes is the same)
    if cells in ((8, 8, 16), (16, 16, 8), (8, 8, 24), (8, 4, 12)):
        assert fused == {True}, (cells, fused)


# ---- BASELINE-scale partition: 256^3 cells in total over 4 ranks against the UNDIVIDED oracle ----------

def _worker_big(rank, world, port, cells, tmpdir, out):
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [here, os.path.join(here, "..", "dune-pdelab_b200", "python")]
        import torch.distributed as dist
        from pdelab_b200.capi import GridOperator
        from pdelab_b200.partition import OverlappingPartition, P2PHaloExchanger
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        part = OverlappingPartition.strong(cells, world, rank)
        n = 27
        zg = np.load(os.path.join(tmpdir, "z.npy"), mmap_mode="r")        # [ncells, 27]
        kg = np.load(os.path.join(tmpdir, "kappa.npy"), mmap_mode="r")
        want = np.load(os.path.join(tmpdir, "want.npy"), mmap_mode="r")
        gidx = part.local_cell_grid().reshape(-1)
        own = part.owned_mask().reshape(-1)
        z = np.full((gidx.size, n), 1e300)          # ghosts poisoned: the exchange must fill them
        z[own] = zg[gidx[own]]
        spec = abi.ProblemSpec(part.local_cells, degree=2, lower=part.local_lower, upper=part.local_upper, alpha=3.0,
                               a_mode=abi.A_SCALAR, A=np.ascontiguousarray(kg[gidx]), side_kind=part.side_kind, device=dev)
        go = GridOperator(spec)
        halo = P2PHaloExchanger(go, part, dist)
        zd = torch.from_numpy(z.reshape(-1)).cuda()
        del z
        yd = torch.full_like(zd, float("nan"))
        halo.apply(zd, yd)
        go.synchronize()
        y = yd.cpu().numpy().reshape(-1, n)
        w = want[gidx[own]]
        err = float(np.abs(y[own] - w).max() / np.abs(w).max())
        out.put((rank, err, go.last_kernel(), int(own.sum()), None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        out.put((rank, 1.0, "", 0, traceback.format_exc()))


def test_256_cubed_over_four_ranks_matches_the_undivided_oracle(cuda_lib, tmp_path):
    """SURVEY 8e parity definition at BASELINE scale: DG k=2 on 256^3 cells (453 M DOFs) split 1x2x2 with one ghost
    layer; the owned rows of every rank equal the single-domain oracle on the global grid to 1e-12."""
    import torch.multiprocessing as mp
    from oracle import Oracle
    from problems import kappa_field, mt_vector
    cells, world = (256, 256, 256), 4
    ncg = int(np.prod(cells))
    zg = mt_vector(ncg * 27)
    kg = kappa_field(ncg)
    try:
        threads = max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        threads = os.cpu_count() or 1
    want = Oracle(abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kg)).jacobian_apply(zg, threads=threads)
    np.save(tmp_path / "z.npy", zg.reshape(ncg, 27))
    np.save(tmp_path / "kappa.npy", kg)
    np.save(tmp_path / "want.npy", want.reshape(ncg, 27))
    del zg, want
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_big, args=(r, world, port, cells, str(tmp_path), out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    owned = 0
    for rank, err, kern, nown, tb in res:
        assert tb is None, tb
        assert err < 1e-12, (rank, err)
        assert kern == "dg_fast_q2_3d+halo"    # the one-launch step: 128 owned layers are tile-aligned
        owned += nown
    assert owned == ncg          # every cell has exactly one owner


# ---- conforming Qk on the overlapping partition: gather / scatter kernels + QkHaloExchanger ----------

class _HostStagedDist:
    """torch.distributed look-alike that stages CUDA buffers through the host, so that several ranks can
    share one GPU over gloo in this test (on a multi-GPU box bench/production use NCCL directly)."""

    def __init__(self, dist):
        self.dist = dist
        self.isend, self.irecv = "isend", "irecv"

    def P2POp(self, op, buf, peer):
        return (op, buf, peer)

    def batch_isend_irecv(self, ops):
        real, back = [], []
        for op, buf, peer in ops:
            host = buf.cpu() if op == "isend" else torch.empty(buf.shape, dtype=buf.dtype)
            real.append(self.dist.P2POp(self.dist.isend if op == "isend" else self.dist.irecv, host, peer))
            if op == "irecv":
                back.append((buf, host))
        reqs = self.dist.batch_isend_irecv(real)

        class _Done:
            def __init__(s, r, last):
                s.r, s.last = r, last

            def wait(s):
                s.r.wait()
                if s.last:
                    for buf, host in back:
                        buf.copy_(host)
        return [_Done(r, i == len(reqs) - 1) for i, r in enumerate(reqs)]


def _qk_worker(rank, world, port, cells, degree, out):
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [os.path.join(here, "..", "oracle"), here, os.path.join(here, "..", "dune-pdelab_b200", "python")]
        import torch.distributed as dist
        from oracle import Oracle
        from pdelab_b200.capi import GridOperator
        from pdelab_b200.partition import OverlappingPartition, QkHaloExchanger
        from problems import kappa_field, mt_vector
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        part = OverlappingPartition.strong(cells, world, rank)
        ncg = int(np.prod(cells))
        kg = kappa_field(ncg)
        gspec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=degree, a_mode=abi.A_SCALAR, A=kg)
        zg = mt_vector(gspec.num_dofs)
        want = Oracle(gspec).jacobian_apply(zg)
        gidx = part.local_cell_grid().reshape(-1)
        spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QK, degree=degree, lower=part.local_lower,
                               upper=part.local_upper, a_mode=abi.A_SCALAR, A=kg[gidx], side_kind=part.side_kind, device=dev)
        go = GridOperator(spec)
        halo = QkHaloExchanger(go, part, degree, torch.device("cuda", dev), dist=_HostStagedDist(dist))
        own = halo.owned_point_mask()
        gp = halo.global_point_index(cells)
        z = np.full(own.size, 1e300)                 # everything not owned is poisoned
        z[own] = zg[gp[own]]
        zd = torch.from_numpy(z).cuda()
        yd = torch.full_like(zd, float("nan"))
        halo.exchange(zd)
        go.apply(zd, yd)
        go.synchronize()
        consistent = bool(np.array_equal(zd.cpu().numpy(), zg[gp]))
        y = yd.cpu().numpy()
        err = float(np.abs(y[own] - want[gp[own]]).max() / np.abs(want).max())
        out.put((rank, err, consistent, go.last_kernel(), int(own.sum()), None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, 1.0, False, "", 0, traceback.format_exc()))


@pytest.mark.parametrize("world,cells,degree", [(2, (8, 6, 10), 2), (4, (6, 10, 8), 1)])
def test_qk_halo_apply_matches_global_oracle(cuda_lib, world, cells, degree):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_qk_worker, args=(r, world, port, cells, degree, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sum(r[4] for r in res) == int(np.prod([degree * c + 1 for c in cells]))
    for rank, err, consistent, kern, _, tb in res:
        assert tb is None, tb
        assert consistent, rank
        assert err < 1e-12, (rank, err)
        assert kern == "fem_kron"


# ---- overlapping Krylov solvers: OverlappingOperator + OverlappingScalarProduct on the device ---------------

def _solve_worker(rank, world, port, cells, solver, precond, assembled, out):
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [os.path.join(here, "..", "oracle"), here, os.path.join(here, "..", "dune-pdelab_b200", "python")]
        import torch.distributed as dist
        from oracle import Oracle
        from pdelab_b200.capi import GridOperator
        from pdelab_b200.partition import OverlappingPartition, OverlappingSolverBackend
        from problems import kappa_field, mt_vector
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        part = OverlappingPartition.strong(cells, world, rank)
        n = 27
        ncg = int(np.prod(cells))
        bg = mt_vector(ncg * n, seed=3).reshape(ncg, n)
        kg = kappa_field(ncg)
        gidx = part.local_cell_grid().reshape(-1)
        own = part.owned_mask().reshape(-1)
        b = np.full((gidx.size, n), 1e300)           # ghost rows of the right-hand side are ignored
        b[own] = bg[gidx[own]]
        spec = abi.ProblemSpec(part.local_cells, degree=2, lower=part.local_lower, upper=part.local_upper, alpha=3.0,
                               a_mode=abi.A_SCALAR, A=kg[gidx], side_kind=part.side_kind, device=dev)
        go = GridOperator(spec)
        ls = OverlappingSolverBackend(go, part, dist, solver=solver, precond=precond, maxiter=2000)
        # comm().sum: every rank contributes (rank + 1, 1)
        s = go.comm_sum(np.array([rank + 1.0, 1.0]))
        sum_ok = s[0] == world * (world + 1) / 2 and s[1] == world
        bd = torch.from_numpy(b.reshape(-1)).cuda()
        zd = torch.zeros_like(bd)
        if assembled:
            rowptr, colidx = go.fill_pattern()
            values = torch.zeros(colidx.size, dtype=torch.float64, device="cuda")
            go.jacobian(torch.zeros_like(bd), values, fresh=True)
            res = ls.apply(values, zd, bd, 1e-9)
        else:
            res = ls.apply(zd, bd, 1e-9)
        go.synchronize()
        z = zd.cpu().numpy().reshape(-1, n)
        # gather the owned parts into the global vector and check  J z = b  with the single-domain oracle
        parts = [None] * world
        dist.all_gather_object(parts, (gidx[own], z[own]))
        zg = np.zeros((ncg, n))
        for gi, zi in parts:
            zg[gi] = zi
        gspec = abi.ProblemSpec(cells, degree=2, alpha=3.0, a_mode=abi.A_SCALAR, A=kg)
        jz = Oracle(gspec).jacobian_apply(zg.reshape(-1)).reshape(-1, n)
        err = float(np.linalg.norm(jz - bg) / np.linalg.norm(bg))
        # the returned solution is consistent: face ghosts hold the neighbour's owned values
        lc = part.local_cells
        coords = np.unravel_index(np.arange(gidx.size), lc[::-1])
        outside = np.zeros(gidx.size, dtype=int)
        for d in range(3):
            c = coords[2 - d] + part.local_lo[d]
            outside += ~((part.owned_lo[d] <= c) & (c < part.owned_hi[d]))
        fg = outside == 1
        consistent = bool(np.array_equal(z[fg], zg[gidx[fg]]))
        # reference count: the same solver on the undivided grid (rank 0 only)
        ref_it = None
        if rank == 0:
            gg = GridOperator(gspec.replace(device=dev))
            zz = torch.zeros(ncg * n, dtype=torch.float64, device="cuda")
            bb = torch.from_numpy(bg.reshape(-1).copy()).cuda()
            vals = None
            if assembled:
                rp, ci = gg.fill_pattern()
                vals = torch.zeros(ci.size, dtype=torch.float64, device="cuda")
                gg.jacobian(torch.zeros_like(bb), vals, fresh=True)
            ref_it = gg.solve(zz, bb, 1e-9, solver=solver, precond=precond, values=vals, maxiter=2000)["iterations"]
        out.put((rank, err, res, sum_ok and consistent, ref_it, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, 1.0, {}, False, None, traceback.format_exc()))


@pytest.mark.parametrize("world,cells,solver,precond,assembled", [
    (2, (8, 6, 12), abi.SOLVER_CG, abi.PRECOND_BLOCK_JACOBI, False),
    (2, (8, 6, 12), abi.SOLVER_BICGSTAB, abi.PRECOND_NONE, False),
    (4, (8, 12, 10), abi.SOLVER_CG, abi.PRECOND_JACOBI, False),
    (2, (4, 4, 8), abi.SOLVER_BICGSTAB, abi.PRECOND_JACOBI, True),
])
def test_overlapping_solver_matches_global_problem(cuda_lib, world, cells, solver, precond, assembled):
    """ISTLBackend_OVLP-style solve on 2 and 4 ranks: the gathered solution solves the undivided problem (oracle J),
    every rank reports the same iteration count, and that count is the single-domain solver's (same Krylov recurrences,
    sums in a different order)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solve_worker, args=(r, world, port, cells, solver, precond, assembled, out))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    its = set()
    ref_it = None
    for rank, err, r, ok, rit, tb in res:
        assert tb is None, tb
        assert r["converged"] == 1, (rank, r)
        assert err < 1e-7, (rank, err)
        assert ok, rank
        its.add(r["iterations"])
        ref_it = rit if rit is not None else ref_it
    assert len(its) == 1, its
    assert abs(its.pop() - ref_it) <= max(2, ref_it // 10), (res[0][2], ref_it)


# ---- conforming Qk on the peer-to-peer mailboxes (csrc/halo.cu: lattice planes, direction by direction) and in the
# overlapping Krylov solvers (owner mask = the planes a rank receives) -------------------------------------------

def _qk_p2p_worker(rank, world, port, cells, degree, solve, out):
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [os.path.join(here, "..", "oracle"), here, os.path.join(here, "..", "dune-pdelab_b200", "python")]
        import torch.distributed as dist
        from oracle import Oracle
        from pdelab_b200.capi import GridOperator
        from pdelab_b200.partition import OverlappingPartition, OverlappingSolverBackend, P2PHaloExchanger, QkHaloExchanger
        from problems import kappa_field, mt_vector
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        part = OverlappingPartition.strong(cells, world, rank)
        ncg = int(np.prod(cells))
        kg = kappa_field(ncg)
        gspec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=degree, a_mode=abi.A_SCALAR, A=kg)
        gidx = part.local_cell_grid().reshape(-1)
        spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QK, degree=degree, lower=part.local_lower,
                               upper=part.local_upper, a_mode=abi.A_SCALAR, A=kg[gidx], side_kind=part.side_kind, device=dev)
        go = GridOperator(spec)
        # index bookkeeping only (owner mask, local -> global numbering); no communication through this object
        book = QkHaloExchanger(go, part, degree, torch.device("cuda", dev), dist=dist)
        own, gp = book.owned_point_mask(), book.global_point_index(cells)
        if solve is None:
            halo = P2PHaloExchanger(go, part, dist)
            zg = mt_vector(gspec.num_dofs)
            want = Oracle(gspec).jacobian_apply(zg)
            z = np.full(own.size, 1e300)                 # everything not owned is poisoned
            z[own] = zg[gp[own]]
            zd = torch.from_numpy(z).cuda()
            errs, consistent = [], True
            for it in range(3):                          # several epochs: flags / acks keep working
                zd.copy_(torch.from_numpy(z))
                yd = torch.full_like(zd, float("nan"))
                if it == 1:
                    halo.exchange(zd)                    # the plain exchange entry point ...
                    go.apply(zd, yd)
                else:
                    halo.apply(zd, yd)                   # ... and y = J x with the exchange in front
                go.synchronize()
                consistent &= bool(np.array_equal(zd.cpu().numpy(), zg[gp]))
                y = yd.cpu().numpy()
                errs.append(float(np.abs(y[own] - want[gp[own]]).max() / np.abs(want).max()))
            out.put((rank, max(errs), consistent, go.last_kernel(), int(own.sum()), None, None))
        else:
            solver, precond, assembled = solve
            ls = OverlappingSolverBackend(go, part, dist, solver=solver, precond=precond, maxiter=3000)
            bg = mt_vector(gspec.num_dofs, seed=3)
            gg = GridOperator(gspec.replace(device=dev))
            gcon = gg.constrained_dofs().astype(np.int64)
            bg[gcon] = 0.0                               # Dirichlet rows: zero defect (the update stays zero there)
            b = np.full(own.size, 1e300)                 # rows that are not owned are ignored
            b[own] = bg[gp[own]]
            bd = torch.from_numpy(b).cuda()
            zd = torch.zeros_like(bd)
            if assembled:
                rowptr, colidx = go.fill_pattern()
                values = torch.zeros(colidx.size, dtype=torch.float64, device="cuda")
                go.jacobian(torch.zeros_like(bd), values, fresh=True)
                res = ls.apply(values, zd, bd, 1e-9)
            else:
                res = ls.apply(zd, bd, 1e-9)
            go.synchronize()
            z = zd.cpu().numpy()
            parts = [None] * world
            dist.all_gather_object(parts, (gp[own], z[own]))
            zg = np.zeros(gspec.num_dofs)
            for gi, zi in parts:
                zg[gi] = zi
            jz = Oracle(gspec).jacobian_apply(zg)
            free = np.ones(zg.size, dtype=bool)
            free[gcon] = False
            err = float(np.linalg.norm((jz - bg)[free]) / np.linalg.norm(bg[free]))
            consistent = bool(np.array_equal(z, zg[gp]))   # consistent on the whole extended box on return
            ref_it = None
            if rank == 0:
                zz = torch.zeros(gspec.num_dofs, dtype=torch.float64, device="cuda")
                bb = torch.from_numpy(bg.copy()).cuda()
                vals = None
                if assembled:
                    rp, ci = gg.fill_pattern()
                    vals = torch.zeros(ci.size, dtype=torch.float64, device="cuda")
                    gg.jacobian(torch.zeros_like(bb), vals, fresh=True)
                ref_it = gg.solve(zz, bb, 1e-9, solver=solver, precond=precond, values=vals, maxiter=3000)["iterations"]
            out.put((rank, err, consistent, go.last_kernel(), int(own.sum()), res, ref_it))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        out.put((rank, 1.0, False, "", 0, traceback.format_exc(), None))


def _run_qk_p2p(world, cells, degree, solve):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_qk_p2p_worker, args=(r, world, port, cells, degree, solve, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert not isinstance(r[5], str), r[5]
    assert sum(r[4] for r in res) == int(np.prod([degree * c + 1 for c in cells]))   # every point has one owner
    return res


@pytest.mark.parametrize("world,cells,degree", [(2, (8, 6, 10), 2), (4, (6, 10, 8), 1), (4, (6, 9, 8), 2), (2, (9, 12), 2),
                                                (4, (10, 12), 1)])
def test_qk_p2p_mailbox_exchange_and_apply_match_global_oracle(cuda_lib, world, cells, degree):
    for rank, err, consistent, kern, _, _, _ in _run_qk_p2p(world, cells, degree, None):
        assert consistent, rank          # after the exchange the extended box equals the global vector bit for bit
        assert err < 1e-12, (rank, err)  # owned rows equal the undivided oracle
        assert kern == "fem_kron"


@pytest.mark.parametrize("world,cells,degree,solver,precond,assembled", [
    (2, (8, 6, 10), 2, abi.SOLVER_CG, abi.PRECOND_JACOBI, False),
    (4, (6, 10, 8), 1, abi.SOLVER_BICGSTAB, abi.PRECOND_NONE, False),
    (4, (6, 8, 8), 2, abi.SOLVER_CG, abi.PRECOND_NONE, False),
    (2, (6, 4, 8), 2, abi.SOLVER_BICGSTAB, abi.PRECOND_JACOBI, True),
    (2, (12, 10), 1, abi.SOLVER_CG, abi.PRECOND_JACOBI, True),
])
def test_qk_overlapping_solver_matches_global_problem(cuda_lib, world, cells, degree, solver, precond, assembled):
    """ISTLBackend_OVLP-style solve of a conforming Qk problem on 2 and 4 ranks (ovlpistlsolverbackend.hh:40-134 works
    for any GFS): the gathered solution solves the undivided problem on the unconstrained rows, it is consistent on the
    overlap on return, all ranks agree on the iteration count and it is the single-domain solver's."""
    res = _run_qk_p2p(world, cells, degree, (solver, precond, assembled))
    its, ref_it = set(), None
    for rank, err, consistent, kern, _, r, rit in res:
        assert r["converged"] == 1, (rank, r)
        assert err < 1e-7, (rank, err)
        assert consistent, rank
        its.add(r["iterations"])
        ref_it = rit if rit is not None else ref_it
    assert len(its) == 1, its
    assert abs(its.pop() - ref_it) <= max(2, ref_it // 10), (res[0][5], ref_it)

#!/usr/bin/env python
"""bench.py — DG Q2 3D jacobian_apply throughput (BASELINE.json metric) on 1..8 B200.

A "step" is one OnTheFlyOperator::apply (y = J z, backend/istl/seqistlsolverbackend.hh:66-76) of
the ConvectionDiffusionDG SIPG operator on a QkDG k=2 YaspGrid with 128^3 cells per GPU
(BASELINE.json configs[1]); N > 1 ranks form an overlapping Cartesian partition (weak scaling)
and exchange the ghost cell layer of z over NCCL before every apply.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells C]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port (the reference
cannot be built in this image) on the host cores for the same metric.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "dune-pdelab_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "dg_q2_3d_jacobian_apply_dof_per_s"
UNIT = "DOF/s"
ALPHA = 3.0  # SIPG, weightsOn, alpha=3 (test/matrixfree/matrix_free_linear.cc:105-108)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz, self.thread = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_threads():
    """All host threads this process may use.  Deliberately NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently turn the CPU arm into a one-core run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(cells, threads=None):
    """The oracle port timed on the host cores on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import Oracle
    from problems import dg_problem
    spec = dg_problem((cells,) * 3, degree=2, a="scalar", alpha=ALPHA)
    try:
        orc = Oracle(spec, native=True)   # -O3 -march=native build made on this box
        build = "g++ -O3 -march=native -fopenmp"
    except Exception:
        orc = Oracle(spec, native=False)
        build = "g++ -O2 -fopenmp"
    threads = threads or host_threads()   # passed explicitly (omp num_threads clause), see host_threads()
    z = np.random.default_rng(0).random(spec.num_dofs)
    y = np.zeros_like(z)
    orc.jacobian_apply(z[:], y, threads=threads)  # warm-up (page faults)
    best = 1e30
    for _ in range(2):
        y[:] = 0.0
        t0 = time.perf_counter()
        orc.jacobian_apply(z, y, threads=threads)
        best = min(best, time.perf_counter() - t0)
    return {"value": spec.num_dofs / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one jacobian_apply on {cells}^3 cells ({spec.num_dofs} DOFs), same operator and "
                      f"coefficients family, {build}, best of 2, {best:.3f} s"}


FP64_PEAK_TFLOPS = 34.75   # measured DFMA peak of this pool's B200 (profiles/r01_fp64_peak.txt); DMMA shares the pipe


def _time_events(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def other_configs(torch, dev, peak):
    """BASELINE.json configs[0], [2], [3] (cfg1, cfg3, cfg4 of SURVEY.md §8d) on one GPU, timed like the headline:
    CUDA events on the launching stream, vectors / matrices resident, warm-up first.  Algorithmic bytes per entry
    point as in SURVEY.md §8d / DESIGN.md §5 (stated in each record)."""
    from pdelab_b200 import abi
    from pdelab_b200.capi import GridOperator

    def rand(n, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        return torch.rand(n, dtype=torch.float64, device=dev, generator=g)

    def rec(ms, alg_bytes, units, unit, kernel, note, flops=None):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        r = {"ms": ms, "value": units / (ms * 1e-3), "unit": unit, "kernel": kernel,
             "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                          "algorithmic_bytes_per_launch": alg_bytes}, "note": note}
        if flops is not None:
            # the fp64 pipe issues one instruction (DFMA, DADD or DMUL alike) per slot: the fraction of the pipe's issue
            # slots the kernel fills is (fp64 instructions executed) x 2 / time against the DFMA peak in flop/s
            tf = flops["slots"] * 2.0 / (ms * 1e-3) / 1e12
            r["roofline_fp64"] = {"bound": "fp64", "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s (DFMA-equivalent)",
                                  "frac": tf / FP64_PEAK_TFLOPS, "fp64_instructions_per_launch": flops["slots"],
                                  "note": flops["note"] + "; peak = measured DFMA rate (profiles/r01_fp64_peak.txt)",
                                  "canonical_tflops": flops["canonical"] / (ms * 1e-3) / 1e12,
                                  "canonical_note": "SURVEY.md 8d count for a sum-factorised quadrature kernel (357 flop/DOF "
                                                    "at k=4); the Kronecker form executes about a third of it, so this "
                                                    "rate is NOT a pipe utilisation"}
        return r

    out = {}

    def fem(cells, k, reps):
        nc = int(np.prod(cells))
        nq = (k + 1) ** len(cells)
        kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
        spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=k, a_mode=abi.A_SCALAR, A=kappa, f=rand(nc * nq, 1))
        go = GridOperator(spec)
        go.set_stream(torch.cuda.current_stream().cuda_stream)
        n = spec.num_dofs
        x, r = rand(n, 2), torch.zeros(n, dtype=torch.float64, device=dev)
        res = {"dofs": n, "cells": list(cells)}
        ms = _time_events(torch, lambda: go.residual(x, r), reps)
        res["residual"] = rec(ms, 32.0 * n + 8.0 * nc, n, "DOF/s", go.last_kernel(),
                              "r += R(x): read x, read+write r, read cached R(0) (32 B/DOF) + kappa per cell")
        ms = _time_events(torch, lambda: go.apply(x, r), reps)
        res["jacobian_apply"] = rec(ms, 16.0 * n + 8.0 * nc, n, "DOF/s", go.last_kernel(), "y = J x: 16 B/DOF + kappa")
        nr, nnz = go.pattern_size()
        res["nnz"] = nnz
        rowptr = torch.empty(nr + 1, dtype=torch.int64, device=dev)
        colidx = torch.empty(nnz, dtype=torch.int32, device=dev)
        mreps = max(2, reps // 4)
        ms = _time_events(torch, lambda: go.fill_pattern(rowptr=rowptr, colidx=colidx, index32=True), mreps, warm=1)
        res["fill_pattern"] = rec(ms, 4.0 * nnz + 8.0 * (nr + 1), nnz, "nnz/s", "qk_interior_kernel<pattern> + row-gather kernels",
                                  "u32 colidx + u64 rowptr written once")
        vals = torch.empty(nnz, dtype=torch.float64, device=dev)
        ms = _time_events(torch, lambda: go.jacobian(x, vals, fresh=True), mreps, warm=1)
        res["jacobian"] = rec(ms, 8.0 * nnz + 8.0 * nc, nnz, "nnz/s", "qk_interior_values_kernel + qk_assemble_kernel",
                              "A = 0; jacobian(x, A): 8 B per stored non-zero + kappa")
        y = torch.empty(n, dtype=torch.float64, device=dev)
        ms = _time_events(torch, lambda: go.csr_mv(vals, x, y), mreps, warm=1)
        res["spmv"] = rec(ms, 8.0 * nnz + 16.0 * n, nnz, "nnz/s", "qk_mv_interior + qk_mv_kernel",
                          "y = A x: 8 B per non-zero (colidx is decoded arithmetically, never read) + vectors")
        del go, vals, colidx, rowptr
        torch.cuda.empty_cache()
        return res

    # cfg2 with convection and reaction (second variant): b = (1, 0.5, 0.25), c = 1, same kernel family
    cells = (128, 128, 128)
    nc = 128 ** 3
    kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
    bvec = torch.tensor([1.0, 0.5, 0.25], dtype=torch.float64, device=dev).repeat(nc, 1).contiguous()
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=2, alpha=ALPHA, a_mode=abi.A_SCALAR, A=kappa, b=bvec,
                           c=torch.ones(nc, dtype=torch.float64, device=dev))
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    n = spec.num_dofs
    x, r = rand(n, 2), torch.zeros(n, dtype=torch.float64, device=dev)
    ms = _time_events(torch, lambda: go.apply(x, r), 100, warm=10)
    out["cfg2_dg_k2_3d_128_convection"] = {
        "dofs": n, "cells": list(cells),
        "jacobian_apply": rec(ms, 16.0 * n + 40.0 * nc, n, "DOF/s", go.last_kernel(),
                              "y = J x with b = (1, 0.5, 0.25), c = 1: 16 B/DOF + kappa, b (3), c per cell")}
    del go, x, r, bvec
    torch.cuda.empty_cache()
    out["cfg1_q1_2d_256"] = fem((256, 256), 1, 200)
    out["cfg1_q1_2d_256"]["note"] = ("0.5 MB working set: L2-resident and launch-latency bound by construction "
                                     "(the reference's own CPU-runnable case)")
    # cfg3: DG k=4 64^3 sum-factorised residual
    cells, k = (64, 64, 64), 4
    nc, nloc = 64 ** 3, 125
    kappa = 10.0 ** (2.0 * rand(nc, 42) - 1.0)
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QKDG, degree=k, alpha=ALPHA, a_mode=abi.A_SCALAR, A=kappa,
                           f=rand(nc * nloc, 1))
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    n = spec.num_dofs
    x, r = rand(n, 2), torch.zeros(n, dtype=torch.float64, device=dev)
    c3 = {"dofs": n, "cells": list(cells)}
    # dg_kron_3d_kernel<4>: 1661 fp64 instructions per thread (1306 DFMA + 84 DADD + 271 DMUL, cuobjdump -sass), five
    # threads per 125-DOF cell -> 66.4 fp64 instructions per DOF
    k4 = {"slots": 1661.0 * 5.0 * nc, "canonical": 357.0 * n,
          "note": "1661 fp64 instructions per thread x 5 threads per cell (SASS count), 66.4 per DOF"}
    ms = _time_events(torch, lambda: go.residual(x, r), 20)
    c3["residual"] = rec(ms, 32.0 * n + 8.0 * nc, n, "DOF/s", go.last_kernel(),
                         "r += R(x) = J x + cached R(0): 32 B/DOF + kappa; compute-bound config", flops=k4)
    ms = _time_events(torch, lambda: go.apply(x, r), 20)
    c3["jacobian_apply"] = rec(ms, 16.0 * n + 8.0 * nc, n, "DOF/s", go.last_kernel(), "y = J x: 16 B/DOF + kappa", flops=k4)
    out["cfg3_dg_k4_3d_64"] = c3
    del go, x, r
    torch.cuda.empty_cache()
    out["cfg4_q2_3d_160"] = fem((160, 160, 160), 2, 8)
    return out


def strong_512(torch, dist, dev, world, rank, local_rank, steps=10):
    """BASELINE.json configs[4]: DG k=2 on 512^3 cells in total, split over the ranks (strong scaling)."""
    from pdelab_b200 import abi
    from pdelab_b200.capi import GridOperator
    from pdelab_b200.partition import OverlappingPartition, P2PHaloExchanger, exchange_cell_field
    G = 512
    part = OverlappingPartition.strong((G, G, G), world, rank, overlap=1)
    ncl = int(np.prod(part.local_cells))
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    kappa = 10.0 ** (2.0 * torch.rand(ncl, dtype=torch.float64, device=dev, generator=g) - 1.0)
    exchange_cell_field(kappa.view(part.local_cells[::-1]), part, dist)
    spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QKDG, degree=2, lower=part.local_lower,
                           upper=part.local_upper, method=abi.DG_SIPG, weights=abi.DG_WEIGHTS_ON, alpha=ALPHA,
                           a_mode=abi.A_SCALAR, A=kappa, side_kind=part.side_kind, device=local_rank)
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    halo = P2PHaloExchanger(go, part, dist)
    z = torch.rand(spec.num_dofs, dtype=torch.float64, device=dev, generator=g)
    y = torch.empty_like(z)
    for _ in range(3):
        halo.apply(z, y)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        halo.apply(z, y)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    go.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dofs = G ** 3 * 27
    res = {"global_cells": [G] * 3, "dofs": dofs, "partition": "x".join(str(v) for v in part.procs),
           "cells_per_gpu": list(part.owned_cells), "ms_per_step": t.item(), "value": dofs / (t.item() * 1e-3),
           "unit": UNIT, "steps": steps, "scaling": "strong", "kernel": go.last_kernel()}
    del halo, go, z, y, kappa
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = []
    base = None
    for _ in range(max(1, args.warmup > 0)):
        base = cpu_baseline(args.ref_cells)
    for _ in range(max(1, min(args.steps, 3))):
        base = cpu_baseline(args.ref_cells)
        t_all.append(base["value"])
    v = float(np.median(t_all))
    base["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * 27 * args.ref_cells ** 3 / v,  # a step = one apply on the sample
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ConvectionDiffusionDG SIPG QkDG k=2 3D jacobian_apply (CPU oracle port of the "
                               "reference algorithm; bounded sample)", "cells": [args.ref_cells] * 3},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pdelab_b200 import abi
    from pdelab_b200.capi import GridOperator
    from pdelab_b200.partition import OverlappingPartition, HaloExchanger, P2PHaloExchanger, exchange_cell_field
    from problems import kappa_field

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    C = args.cells
    if args.global_cells:   # strong scaling: BASELINE.json configs[4] (512^3 cells over 2/4/8 GPUs)
        G = args.global_cells
        part = OverlappingPartition.strong((G, G, G), world, rank, overlap=1)
    else:
        part = OverlappingPartition.weak((C, C, C), world, rank, overlap=1)
    ncl = int(np.prod(part.local_cells))
    # synthetic coefficient field kappa_e = 10^(2u-1), generated on the device per rank
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    kappa = 10.0 ** (2.0 * torch.rand(ncl, dtype=torch.float64, device=dev, generator=g) - 1.0)
    if world == 1 and ncl <= 200_000:
        kappa = torch.from_numpy(kappa_field(ncl)).to(dev)
    if world > 1:  # ghost cells carry the owner's coefficient (set-up, not timed)
        exchange_cell_field(kappa.view(part.local_cells[::-1]), part, dist)
    spec = abi.ProblemSpec(part.local_cells, space=abi.SPACE_QKDG, degree=2, lower=part.local_lower,
                           upper=part.local_upper, method=abi.DG_SIPG, weights=abi.DG_WEIGHTS_ON, alpha=ALPHA,
                           a_mode=abi.A_SCALAR, A=kappa, side_kind=part.side_kind, device=local_rank)
    go = GridOperator(spec)
    go.set_stream(torch.cuda.current_stream().cuda_stream)
    halo, halo_kind = None, "none"
    if world > 1:
        # default: peer-to-peer mailboxes over NVLink with the exchange hidden behind the interior
        # tiles (csrc/halo.cu); --halo nccl: pack -> NCCL send/recv -> unpack, then the full kernel
        if args.halo == "p2p":
            try:
                halo, halo_kind = P2PHaloExchanger(go, part, dist), "p2p-mailbox (CUDA IPC over NVLink), overlapped"
            except Exception as e:  # noqa: BLE001  (no peer access on this box: say so, use NCCL)
                print(f"rank {rank}: p2p halo unavailable ({e}); using NCCL", file=sys.stderr)
        if halo is None:
            halo, halo_kind = HaloExchanger(go, part, dev), "pack + NCCL send/recv + unpack, not overlapped"
    ndofs = spec.num_dofs
    owned_dofs = int(np.prod(part.owned_cells)) * 27
    if args.global_cells:
        C = None
    z = torch.rand(ndofs, dtype=torch.float64, device=dev, generator=g)
    y = torch.empty_like(z)

    def step():
        if halo is None:
            go.apply(z, y)
        elif isinstance(halo, P2PHaloExchanger):
            halo.apply(z, y)
        else:
            halo.exchange(z)
            go.apply(z, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = go.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    go.synchronize()  # raises if a peer-to-peer wait timed out
    ms_total = e0.elapsed_time(e1)
    launches = go.launch_count() - l0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_step = ms_total / args.steps
    value = owned_dofs * world / (ms_step * 1e-3)

    # duration of the dominant kernel for the roofline.  On one GPU a step IS one launch of that
    # kernel, so the timed region above is the measurement; with a halo exchange in the step the
    # kernel is timed alone (CUDA events on the launching stream) right after.
    if halo is None:
        ms_kernel = ms_step
    else:
        ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ek0.record()
        for _ in range(args.steps):
            go.apply(z, y)
        ek1.record()
        torch.cuda.synchronize()
        ms_kernel = ek0.elapsed_time(ek1) / args.steps
    # sustained regime: the same step back to back for ~1 s (the fp64-heavy kernel runs into
    # sw_power_cap after ~0.1 s: DESIGN.md §6 "Burst vs sustained"); reported next to the burst number
    sus_sampler = ClockSampler(local_rank)
    sus_steps = max(200, int(1.0 / max(ms_step * 1e-3, 1e-6)))
    sus_steps = int(min(sus_steps, 20000))
    for _ in range(sus_steps // 2):
        step()
    barrier()
    sus_sampler.start()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(sus_steps // 2):
        step()
    es1.record()
    barrier()
    sus_clocks = sus_sampler.stop()
    ts = torch.tensor([es0.elapsed_time(es1) / (sus_steps // 2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    sustained = {"ms_per_step": ts.item(), "value": owned_dofs * world / (ts.item() * 1e-3), "unit": UNIT,
                 "steps": sus_steps // 2, "clocks": sus_clocks}
    kernel_name = go.last_kernel()
    peak, peak_src = measured_peak()
    alg_bytes = 16.0 * ndofs + 8.0 * ncl  # 8 B read z + 8 B write y per DOF + 8 B kappa per cell
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                rec = json.load(f)
            if rec.get("cells") == list(part.local_cells) and rec.get("kernel") == kernel_name:
                traffic = rec.get("dram_bytes_per_launch")
        except Exception:
            pass

    # end to end through the C ABI with HOST buffers (pinned): H2D z, apply, D2H y inside the timed region
    e2e_steps = max(1, min(args.steps, 10 if not args.global_cells else 2))
    if args.no_e2e:   # side runs with very large vectors (strong scaling at 512^3): no pinned host copies
        e2e_steps = 0
        zh = yh = None
    else:
        zh = torch.empty(ndofs, dtype=torch.float64).pin_memory()
        yh = torch.empty(ndofs, dtype=torch.float64).pin_memory()
        zh.copy_(z)

    def e2e_step():
        if halo is None:
            go.apply(zh, yh)     # host pointers through the C ABI: H2D, kernel, D2H, synchronous on return
        else:
            z.copy_(zh, non_blocking=True)
            step()
            yh.copy_(y, non_blocking=True)
            torch.cuda.synchronize()

    e2e_value = None
    if e2e_steps:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = owned_dofs * world / te.item()

    # the other BASELINE.json configurations ride in the same line (N = 1: cfg1, cfg3, cfg4; N > 1: cfg5 strong)
    configs = strong = None
    if not args.no_configs and not args.global_cells:
        del zh, yh
        try:
            if world == 1:
                del z, y, go
                torch.cuda.empty_cache()
                configs = other_configs(torch, dev, peak)
            elif isinstance(halo, P2PHaloExchanger):
                del z, y
                torch.cuda.empty_cache()
                strong = strong_512(torch, dist, dev, world, rank, local_rank)
        except Exception as e:  # noqa: BLE001  (the headline line must survive a failure of the side measurements)
            if world == 1:
                configs = {"error": repr(e)}
            else:
                strong = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if args.global_cells else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": (f"ConvectionDiffusionDG SIPG QkDG k=2 on 3D YaspGrid {C}^3 cells per GPU" if C else
                             f"ConvectionDiffusionDG SIPG QkDG k=2 on 3D YaspGrid {args.global_cells}^3 cells in total") +
                            ": matrix-free jacobian_apply (OnTheFlyOperator::apply), fp64",
                "cells_per_gpu": list(part.owned_cells), "global_cells": list(part.global_cells), "dofs_per_gpu": owned_dofs,
                "partition": "x".join(str(v) for v in part.procs), "overlap": 1 if world > 1 else 0,
                "coefficients": "cell-wise scalar kappa=10^(2u-1), b=0, c=0, all-Dirichlet, alpha=3",
                "cache": f"input+output {2 * ndofs * 8 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                "kernel": kernel_name, "halo": halo_kind,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ndofs * 8, "d2h_bytes_per_step": ndofs * 8,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "sustained": sustained,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kernel_name, "kernel_ms": ms_kernel,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.ref_cells)
        if configs is not None:
            line["configs"] = configs
        if strong is not None:
            line["strong_512"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=128, help="cells per direction per GPU")
    ap.add_argument("--global-cells", type=int, default=0,
                    help="strong scaling: cells per direction of the WHOLE grid (e.g. 512), split over the ranks")
    ap.add_argument("--ref-cells", type=int, default=64, help="cells per direction of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/cfg3/cfg4 (N=1) and strong 512^3 (N>1) records")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (side runs only)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="ghost-layer exchange for N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

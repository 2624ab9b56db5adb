"""The C++ host partition (dune-pdelab_b200/host/partition.hh) against the Python one (pdelab_b200/partition.py), rank by
rank: processor grids, owned / local boxes, processor sides, neighbours, local -> global cell maps and owned masks.
CPU only (pure index arithmetic); the same program instantiates the C++ halo exchanger and overlapping solver back-end
so that they stay compile-checked against the C ABI."""
import json
import os
import subprocess

import numpy as np

from pdelab_b200.partition import OverlappingPartition, processor_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_partition.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_partition")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build():
    from pdelab_b200 import capi
    capi.load_library()
    deps = [SRC, os.path.join(ROOT, "dune-pdelab_b200", "host", "partition.hh"),
            os.path.join(ROOT, "dune-pdelab_b200", "host", "gridoperator.hh"), os.path.join(ROOT, "include", "pdelab_b200.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run([GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", EXE, SRC,
                    "-L", os.path.join(ROOT, "dune-pdelab_b200", "lib"), "-lpdelab_b200",
                    "-Wl,-rpath,$ORIGIN/../../../dune-pdelab_b200/lib"], check=True)
    return EXE


def test_cpp_partition_equals_python_partition():
    out = subprocess.run([_build()], capture_output=True, text=True, check=True).stdout
    recs = [json.loads(line) for line in out.splitlines()]
    assert len(recs) > 100
    checked = 0
    for r in recs:
        if r.get("split_x"):
            assert tuple(r["procs"]) == processor_grid(r["world"], 3, split_x=True) + (1,) * 0
            continue
        cells, world, rank = tuple(r["cells"]), r["world"], r["rank"]
        p = OverlappingPartition.weak(cells, world, rank) if r["weak"] else OverlappingPartition.strong(cells, world, rank)
        dim = r["dim"]
        assert tuple(r["procs"][:dim]) == p.procs[:dim] and all(v == 1 for v in r["procs"][dim:])
        for key in ("global_cells", "coords", "owned_lo", "owned_hi", "local_lo", "local_hi", "local_cells"):
            assert tuple(r[key]) == tuple(getattr(p, key)), (key, r, getattr(p, key))
        assert [list(s) for s in r["side_kind"]] == [list(s) for s in p.side_kind]
        assert [tuple(e) for e in r["exchanges"]] == p.exchanges()
        assert np.allclose(r["local_lower"], p.local_lower, rtol=0, atol=0) and np.allclose(r["local_upper"], p.local_upper, rtol=0, atol=0)
        g = p.local_cell_grid().reshape(-1).astype(np.uint64)
        own = p.owned_mask().reshape(-1)
        c = np.arange(g.size, dtype=np.uint64)
        with np.errstate(over="ignore"):
            h = int(((c + np.uint64(1)) * (g + np.uint64(7))).sum(dtype=np.uint64))
            ho = int(((c[own] + np.uint64(3)) * (g[own] + np.uint64(1))).sum(dtype=np.uint64))
        assert int(r["map_hash"]) == h and int(r["owned_hash"]) == ho and r["owned"] == int(own.sum())
        assert r["grid_cells0"] == p.local_cells[0]
        checked += 1
    assert checked == 3 * sum((1, 2, 3, 4, 6, 8, 12))


def test_cpp_time_stepping_tables_equal_the_python_tables():
    """host/onestep.hh and pdelab_b200/onestep.py restate instationary/onestepparameter.hh independently: same entries."""
    from pdelab_b200 import capi, onestep as osm
    capi.load_library()
    src = os.path.join(ROOT, "tests", "cpp", "test_tables.cc")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_tables")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run([GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, src,
                    "-L", os.path.join(ROOT, "dune-pdelab_b200", "lib"), "-lpdelab_b200",
                    "-Wl,-rpath,$ORIGIN/../../../dune-pdelab_b200/lib"], check=True)
    recs = [json.loads(line) for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()]
    assert len(recs) == 9
    for r in recs:
        m = osm.OneStepThetaParameter(0.5) if r["class"].startswith("OneStepTheta") else getattr(osm, r["class"])()
        assert m.s() == r["s"] and m.implicit() == r["implicit"] and m.name() == r["name"], r["class"]
        for k in range(1, m.s() + 1):
            assert [m.a(k, i) for i in range(k + 1)] == r["a"][k - 1], (r["class"], "a", k)
            assert [m.b(k, i) for i in range(k + 1)] == r["b"][k - 1], (r["class"], "b", k)
        assert [m.d(i) for i in range(m.s() + 1)] == r["d"], r["class"]


def test_product_basis_tables_match_independent_numpy_evaluation():
    """The basis-specific part of the CUDA path is its 1-D tables (the generic QkDG kernels are table-driven):
    host_fill_tables of csrc/host_tables.h for Lagrange / Legendre / Gauss-Lobatto, k = 1..4, at the Gauss points and at
    the cell ends, against numpy.polynomial (tests/numpy_assembly.py: basis_tables)."""
    import shutil
    from numpy_assembly import basis_tables
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(ROOT, "tests", "cpp", "test_host_tables.cu")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_tables")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run([nvcc, "-std=c++17", "-O1", "-o", exe, src], check=True)
    recs = [json.loads(line) for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()]
    assert len(recs) == 12
    for r in recs:
        k, n1 = r["k"], r["k"] + 1
        pts = np.array(r["xq"] + [0.0, 1.0])
        V, D = basis_tables(k, pts, r["basis"])
        P = np.array(r["P"]).reshape(len(pts), n1)
        DP = np.array(r["DP"]).reshape(len(pts), n1)
        assert np.abs(P - V).max() < 1e-13 * max(1.0, np.abs(V).max()), (r["basis"], k)
        assert np.abs(DP - D).max() < 1e-12 * max(1.0, np.abs(D).max()), (r["basis"], k)

"""The C++ host mirror (dune-pdelab_b200/host/gridoperator.hh): compiles against the C ABI with
plain g++ (no nvcc, no torch), and — on a GPU box — passes the restated reference tests of
tests/cpp/test_gridoperator.cc (every result also checked against the CPU oracle)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_gridoperator.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_gridoperator")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build():
    from pdelab_b200 import capi
    capi.load_library()                       # libpdelab_b200.so must exist (no fallback)
    import oracle                             # builds oracle/_build/liboracle.so if needed
    oracle.load()
    deps = [SRC, os.path.join(ROOT, "dune-pdelab_b200", "host", "gridoperator.hh"),
            os.path.join(ROOT, "include", "pdelab_b200.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run([GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", EXE, SRC,
                    "-L", os.path.join(ROOT, "dune-pdelab_b200", "lib"), "-lpdelab_b200",
                    "-L", os.path.join(ROOT, "oracle", "_build"), "-loracle",
                    "-Wl,-rpath,$ORIGIN/../../../dune-pdelab_b200/lib",
                    "-Wl,-rpath,$ORIGIN/../../../oracle/_build"], check=True)
    return EXE


def test_cpp_host_builds_and_fails_loudly_without_a_device():
    exe = _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (covered by the gpu test)")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2, out.stdout + out.stderr
    assert "no CUDA device available" in out.stderr


@pytest.mark.gpu
def test_cpp_host_reference_tests_on_gpu():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(out.stdout[-6000:])
    assert out.returncode == 0, out.stdout[-6000:] + out.stderr[-2000:]
    assert "ALL OK" in out.stdout


# ---- the multi-rank classes of host/partition.hh, run for real (fork, one rank per process, handles over socket pairs)
MR_SRC = os.path.join(ROOT, "tests", "cpp", "test_multirank.cc")
MR_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_multirank")


def _build_multirank():
    from pdelab_b200 import capi
    capi.load_library()
    import oracle
    oracle.load()
    deps = [MR_SRC, os.path.join(ROOT, "dune-pdelab_b200", "host", "gridoperator.hh"),
            os.path.join(ROOT, "dune-pdelab_b200", "host", "partition.hh"), os.path.join(ROOT, "include", "pdelab_b200.h")]
    if os.path.exists(MR_EXE) and all(os.path.getmtime(MR_EXE) >= os.path.getmtime(d) for d in deps):
        return MR_EXE
    os.makedirs(os.path.dirname(MR_EXE), exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run([GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(cuda, "include"), "-o", MR_EXE,
                    MR_SRC, "-L", os.path.join(ROOT, "dune-pdelab_b200", "lib"), "-lpdelab_b200",
                    "-L", os.path.join(ROOT, "oracle", "_build"), "-loracle", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                    "-Wl,-rpath,$ORIGIN/../../../dune-pdelab_b200/lib", "-Wl,-rpath,$ORIGIN/../../../oracle/_build",
                    "-Wl,-rpath," + os.path.join(cuda, "lib64")], check=True)
    return MR_EXE


def test_cpp_multirank_program_builds():
    _build_multirank()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_cpp_multirank_halo_exchanger_and_overlapping_solver(world):
    """P2PHaloExchanger and OverlappingSolverBackend of host/partition.hh with 2 and 4 processes on the device: owned rows
    against the undivided oracle (QkDG k=2 and conforming Q2), CG + Jacobi on the partition."""
    exe = _build_multirank()
    out = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=900)
    print(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-3000:]
    assert "ALL OK" in out.stdout

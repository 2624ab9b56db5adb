"""The C-ABI shared library loads and exports every symbol include/pdelab_b200.h declares.
No compute call is made here (CPU-only run); with no CUDA device the library must fail loudly,
never fall back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pdelab_b200 import abi, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pdelab_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libpdelab_b200.so lacks {name}"
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with the header"
    assert b"sm_100a" in lib.pdb200_version()


def test_problem_struct_layout_matches_header():
    """ctypes mirror of struct pdb200_problem: field order and size (LP64)."""
    names = [f[0] for f in abi.Problem._fields_]
    text = open(os.path.join(ROOT, "include", "pdelab_b200.h")).read()
    body = text[text.index("typedef struct pdb200_problem {"):text.index("} pdb200_problem;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decl = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[0-9]\])*;", body)
    assert decl == names
    # size and offsets as the C compiler lays the struct out
    import subprocess
    import tempfile
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"pdelab_b200.h\"\nint main(){printf(\"%zu\", sizeof(pdb200_problem));" + \
        "".join(f'printf(" %zu", offsetof(pdb200_problem, {n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "t.c"), os.path.join(tmp, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == C.sizeof(abi.Problem)
    assert out[1:] == [getattr(abi.Problem, n).offset for n in names]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    spec = abi.ProblemSpec((2, 2), degree=1)
    with pytest.raises(capi.PDELabError, match="no CUDA device"):
        capi.GridOperator(spec)


def test_product_package_does_not_touch_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may use oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "dune-pdelab_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                hit = re.search(r"^\s*(import|from)\s+oracle|liboracle|#\s*include\s*[<\"][^>\"]*oracle|oracle_[a-z_]+\s*\(",
                                text, flags=re.M)
                assert hit is None, (os.path.join(dirpath, fn), hit.group(0))


def test_python_constants_match_the_header_enums():
    """abi.py / onestep.py mirror the enums of include/pdelab_b200.h by value."""
    from pdelab_b200 import onestep as osm
    text = open(os.path.join(ROOT, "include", "pdelab_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enums = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", text, flags=re.S):
        for name, val in re.findall(r"(PDB200_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            enums[name] = int(val)
    assert len(enums) >= 35
    want = {
        "PDB200_SPACE_QKDG": abi.SPACE_QKDG, "PDB200_SPACE_QK": abi.SPACE_QK,
        "PDB200_DG_NIPG": abi.DG_NIPG, "PDB200_DG_SIPG": abi.DG_SIPG, "PDB200_DG_IIPG": abi.DG_IIPG,
        "PDB200_DG_WEIGHTS_ON": abi.DG_WEIGHTS_ON, "PDB200_DG_WEIGHTS_OFF": abi.DG_WEIGHTS_OFF,
        "PDB200_BC_DIRICHLET": abi.BC_DIRICHLET, "PDB200_BC_NEUMANN": abi.BC_NEUMANN, "PDB200_BC_OUTFLOW": abi.BC_OUTFLOW,
        "PDB200_BC_NONE": abi.BC_NONE,
        "PDB200_A_IDENTITY": abi.A_IDENTITY, "PDB200_A_SCALAR": abi.A_SCALAR, "PDB200_A_DIAGONAL": abi.A_DIAGONAL,
        "PDB200_A_FULL": abi.A_FULL,
        "PDB200_SIDE_DOMAIN": abi.SIDE_DOMAIN, "PDB200_SIDE_PROCESSOR": abi.SIDE_PROCESSOR,
        "PDB200_KERNEL_AUTO": abi.KERNEL_AUTO, "PDB200_KERNEL_GENERIC": abi.KERNEL_GENERIC, "PDB200_KERNEL_FAST": abi.KERNEL_FAST,
        "PDB200_BASIS_LAGRANGE": abi.BASIS_LAGRANGE, "PDB200_BASIS_LEGENDRE": abi.BASIS_LEGENDRE,
        "PDB200_BASIS_LOBATTO": abi.BASIS_LOBATTO,
        "PDB200_POINTWISE_A": abi.POINTWISE_A, "PDB200_POINTWISE_B": abi.POINTWISE_B,
        "PDB200_POINTWISE_C": abi.POINTWISE_C, "PDB200_POINTWISE_BCTYPE": abi.POINTWISE_BCTYPE,
        "PDB200_LAYOUT_CSR": abi.LAYOUT_CSR, "PDB200_LAYOUT_BCSR": abi.LAYOUT_BCSR,
        "PDB200_PART_ALL": abi.PART_ALL, "PDB200_PART_INTERIOR": abi.PART_INTERIOR, "PDB200_PART_BOUNDARY": abi.PART_BOUNDARY,
        "PDB200_SOLVER_BICGSTAB": abi.SOLVER_BICGSTAB, "PDB200_SOLVER_CG": abi.SOLVER_CG,
        "PDB200_PRECOND_NONE": abi.PRECOND_NONE, "PDB200_PRECOND_JACOBI": abi.PRECOND_JACOBI,
        "PDB200_PRECOND_BLOCK_JACOBI": abi.PRECOND_BLOCK_JACOBI, "PDB200_PRECOND_BLOCK_SOR": abi.PRECOND_BLOCK_SOR,
        "PDB200_PRECOND_BLOCK_SSOR": abi.PRECOND_BLOCK_SSOR,
        "PDB200_SOR_BACKWARD": abi.SOR_BACKWARD, "PDB200_SOR_KEEP_ITERATE": abi.SOR_KEEP_ITERATE,
        "PDB200_ONESTEP_DIVIDE_OPERATOR1_BY_DT": osm.OneStepGridOperator.DivideOperator1ByDT,
        "PDB200_ONESTEP_MULTIPLY_OPERATOR0_BY_DT": osm.OneStepGridOperator.MultiplyOperator0ByDT,
        "PDB200_ONESTEP_DO_NOT_ASSEMBLE_DT": osm.OneStepGridOperator.DoNotAssembleDT,
    }
    assert enums == want

"""Parity of the CUDA QkDG path with the CPU oracle (-m gpu, through the C ABI)."""
import numpy as np
import pytest

from pdelab_b200 import abi
from problems import dg_problem, mt_vector, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12  # north_star: relative tolerance 1e-12 in fp64 (norm-relative)


def _ops(spec):
    from oracle import Oracle
    from pdelab_b200.capi import GridOperator
    return GridOperator(spec), Oracle(spec)


GENERIC_CASES = [
    dict(cells=(5, 4), degree=1), dict(cells=(5, 4), degree=2, a="full", with_b=True, with_c=True),
    dict(cells=(4, 3), degree=3, a="diagonal", with_c=True), dict(cells=(3, 3), degree=4, a="full", with_b=True),
    dict(cells=(3, 2, 1), degree=1, a="full", with_b=True, with_c=True),
    dict(cells=(4, 3, 2), degree=2, a="full", with_b=True, with_c=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(4, 3, 2), degree=2, a="scalar"), dict(cells=(1, 1, 1), degree=2, a="diagonal"),
    dict(cells=(2, 2, 2), degree=3, a="full", with_b=True), dict(cells=(2, 2, 2), degree=4, a="diagonal", with_c=True),
    dict(cells=(4, 3, 2), degree=2, a="full", with_b=True, bc="mixed"),
    dict(cells=(4, 3, 2), degree=2, a="scalar", method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF),
    dict(cells=(4, 3, 2), degree=2, a="scalar", method=abi.DG_IIPG, intorderadd=1),
]


@pytest.mark.parametrize("case", GENERIC_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_generic_jacobian_apply_matches_oracle(cuda_lib, case):
    spec = dg_problem(kernel=abi.KERNEL_GENERIC, **case)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    y0 = mt_vector(spec.num_dofs, seed=7)          # results are accumulated into y
    y = go.jacobian_apply(z, y0.copy())
    assert go.last_kernel() == "dg_generic_jacobian_apply"
    assert rel_err(y, orc.jacobian_apply(z, y0.copy())) < TOL


@pytest.mark.parametrize("case", GENERIC_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_generic_residual_matches_oracle(cuda_lib, case):
    case = dict(case, with_f=True)
    if case.get("bc", "dirichlet") == "dirichlet":
        case["bc"] = "dirichlet_g"
    spec = dg_problem(kernel=abi.KERNEL_GENERIC, **case)
    go, orc = _ops(spec)
    x = mt_vector(spec.num_dofs)
    r0 = mt_vector(spec.num_dofs, seed=9)
    r = go.residual(x, r0.copy())
    assert rel_err(r, orc.residual(x, r0.copy())) < TOL


FAST_CASES = [
    dict(cells=(8, 4, 4)), dict(cells=(16, 8, 8), a="diagonal", with_c=True),
    dict(cells=(6, 5, 3), a="scalar", extent=(1.0, 0.7, 1.3)),            # partial tiles, anisotropic h
    dict(cells=(2, 1, 1), a="identity"), dict(cells=(10, 9, 7), a="diagonal", bc="mixed"),
    dict(cells=(12, 4, 5), a="scalar", method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF, alpha=1.0),
    dict(cells=(12, 4, 5), a="diagonal", method=abi.DG_IIPG),
    dict(cells=(24, 20, 12), a="scalar"),
]


@pytest.mark.parametrize("case", FAST_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_fast_jacobian_apply_matches_oracle(cuda_lib, case):
    spec = dg_problem(degree=2, kernel=abi.KERNEL_FAST, **case)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    want = orc.jacobian_apply(z)
    y = go.apply(z, np.full(spec.num_dofs, np.nan))          # OnTheFlyOperator: y = J z
    assert go.last_kernel() == "dg_fast_q2_3d"
    assert rel_err(y, want) < TOL
    y0 = mt_vector(spec.num_dofs, seed=7)
    y = go.jacobian_apply(z, y0.copy())                      # accumulate form
    assert rel_err(y, want + y0) < TOL


@pytest.mark.parametrize("case", [dict(cells=(8, 4, 4), a="scalar"), dict(cells=(10, 9, 7), a="diagonal", bc="mixed", with_c=True),
                                  dict(cells=(6, 5, 3), a="scalar", extent=(1.0, 0.7, 1.3))],
                         ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_fast_residual_matches_oracle(cuda_lib, case):
    """residual = J x + R(0) through the Kronecker kernel, R(0) cached per coefficient set."""
    case = dict(case, with_f=True)
    if case.get("bc", "dirichlet") == "dirichlet":
        case["bc"] = "dirichlet_g"
    spec = dg_problem(degree=2, kernel=abi.KERNEL_FAST, **case)
    go, orc = _ops(spec)
    x = mt_vector(spec.num_dofs)
    r0 = mt_vector(spec.num_dofs, seed=9)
    for _ in range(2):                                        # second call uses the cached R(0)
        r = go.residual(x, r0.copy())
        assert go.last_kernel() == "dg_fast_q2_3d+r0"
        assert rel_err(r, orc.residual(x, r0.copy())) < TOL
    # new source term: the cache must be invalidated
    f2 = np.random.default_rng(5).standard_normal(spec.arrays["f"].shape)
    go.update_coefficients(f=f2)
    from oracle import Oracle
    r = go.residual(x, r0.copy())
    assert rel_err(r, Oracle(spec.replace(f=f2)).residual(x, r0.copy())) < TOL


KRON_CASES = [
    dict(cells=(4, 4, 3), degree=4), dict(cells=(2, 1, 1), degree=4, a="identity"),
    dict(cells=(6, 5, 4), degree=4, a="diagonal", with_c=True, bc="mixed"),                   # partial tiles
    dict(cells=(4, 3, 2), degree=4, a="scalar", extent=(1.0, 0.7, 1.3), method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF, alpha=1.0),
    dict(cells=(8, 4, 2), degree=3), dict(cells=(10, 5, 3), degree=3, a="diagonal", with_c=True, bc="mixed", method=abi.DG_IIPG),
]


@pytest.mark.parametrize("case", KRON_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_kron_jacobian_apply_and_residual_match_oracle(cuda_lib, case):
    """Higher-degree Kronecker kernel (csrc/dg_kron.cu): cfg3 of BASELINE.json is k = 4."""
    spec = dg_problem(kernel=abi.KERNEL_FAST, **case)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    want = orc.jacobian_apply(z)
    y = go.apply(z, np.full(spec.num_dofs, np.nan))
    assert go.last_kernel() == "dg_kron_3d"
    assert rel_err(y, want) < TOL
    y0 = mt_vector(spec.num_dofs, seed=7)
    assert rel_err(go.jacobian_apply(z, y0.copy()), want + y0) < TOL
    # residual with source term and boundary data: J x + R(0)
    case = dict(case, with_f=True)
    if case.get("bc", "dirichlet") == "dirichlet":
        case["bc"] = "dirichlet_g"
    spec = dg_problem(kernel=abi.KERNEL_FAST, **case)
    go, orc = _ops(spec)
    r = go.residual(z, y0.copy())
    assert go.last_kernel() == "dg_kron_3d+r0"
    assert rel_err(r, orc.residual(z, y0.copy())) < TOL


def test_fast_and_generic_agree_on_device_tensors(cuda_lib):
    import torch
    spec = dg_problem((32, 16, 12), degree=2, a="scalar")
    from pdelab_b200.capi import GridOperator
    fast = GridOperator(spec.replace(kernel=abi.KERNEL_FAST))
    gen = GridOperator(spec.replace(kernel=abi.KERNEL_GENERIC))
    z = torch.from_numpy(mt_vector(spec.num_dofs)).cuda()
    yf, yg = torch.zeros_like(z), torch.zeros_like(z)
    fast.apply(z, yf)
    gen.apply(z, yg)
    fast.synchronize(), gen.synchronize()
    assert rel_err(yf.cpu().numpy(), yg.cpu().numpy()) < TOL


def test_linearity_and_affinity_full_size_property(cuda_lib):
    """Size-independent properties at a size the oracle cannot do in seconds:
    J(a u + v) = a J u + J v."""
    import torch
    spec = dg_problem((64, 64, 32), degree=2, a="scalar")
    from pdelab_b200.capi import GridOperator
    go = GridOperator(spec)
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.rand(spec.num_dofs, dtype=torch.float64, device="cuda", generator=g)
    v = torch.rand(spec.num_dofs, dtype=torch.float64, device="cuda", generator=g)
    Ju, Jv, Jw = (torch.empty_like(u) for _ in range(3))
    go.apply(u, Ju), go.apply(v, Jv), go.apply(2.5 * u + v, Jw)
    go.synchronize()
    err = (Jw - (2.5 * Ju + Jv)).abs().max() / Jw.abs().max()
    assert err.item() < TOL
    # SIPG with b = 0 is symmetric: v.Ju = u.Jv
    s1, s2 = torch.dot(v, Ju).item(), torch.dot(u, Jv).item()
    assert abs(s1 - s2) / abs(s1) < 1e-11


def test_nonlinear_variant_throws_like_reference(cuda_lib):
    from pdelab_b200.capi import GridOperator, PDELabError
    spec = dg_problem((2, 2), degree=1)
    go = GridOperator(spec)
    z = np.zeros(spec.num_dofs)
    with pytest.raises(PDELabError):
        go.jacobian_apply(z, z, z.copy())


def test_outflow_on_inflow_throws(cuda_lib):
    from pdelab_b200.capi import GridOperator, PDELabError
    spec = dg_problem((3, 3), degree=1, a="identity")
    nc = spec.ncells
    b = np.tile(np.array([1.0, 0.0]), (nc, 1))
    bct = np.full(spec.num_boundary_faces, abi.BC_OUTFLOW, dtype=np.int8)   # x=0 side is inflow
    go = GridOperator(spec.replace(b=b, bctype=bct))
    z = np.ones(spec.num_dofs)
    with pytest.raises(PDELabError, match="Outflow"):
        go.jacobian_apply(z, np.zeros_like(z))


# dg_small.cu: thread-per-cell Kronecker kernel for dim = 2 (k = 1, 2) and dim = 3 (k = 1) — the
# configurations of the reference's own DG tests
SMALL_CASES = [
    dict(cells=(16, 16), degree=1), dict(cells=(7, 5), degree=2, a="diagonal", with_c=True),
    dict(cells=(1, 1), degree=1, a="identity"), dict(cells=(9, 4), degree=2, a="scalar", extent=(1.0, 0.6)),
    dict(cells=(6, 5, 4), degree=1, a="scalar"), dict(cells=(3, 2, 1), degree=1, a="diagonal", with_c=True, bc="mixed"),
    dict(cells=(8, 7), degree=1, a="scalar", method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF, alpha=1.0),
    dict(cells=(8, 7), degree=2, a="diagonal", method=abi.DG_IIPG, bc="mixed"),
    dict(cells=(33, 17, 9), degree=1, a="scalar", extent=(1.0, 0.7, 1.3)),
]


@pytest.mark.parametrize("case", SMALL_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_small_cell_kernel_matches_oracle(cuda_lib, case):
    spec = dg_problem(kernel=abi.KERNEL_FAST, **case)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    want = orc.jacobian_apply(z)
    y = go.apply(z, np.full(spec.num_dofs, np.nan))            # overwrite form
    assert go.last_kernel() == "dg_small"
    assert rel_err(y, want) < TOL
    y0 = mt_vector(spec.num_dofs, seed=7)
    assert rel_err(go.jacobian_apply(z, y0.copy()), want + y0) < TOL   # accumulate form
    # residual = J x + cached R(0), with source term and inhomogeneous boundary data
    rcase = dict(case, with_f=True)
    if rcase.get("bc", "dirichlet") == "dirichlet":
        rcase["bc"] = "dirichlet_g"
    spec = dg_problem(kernel=abi.KERNEL_AUTO, **rcase)
    go, orc = _ops(spec)
    r = go.residual(z, y0.copy())
    assert go.last_kernel() == "dg_small+r0"
    assert rel_err(r, orc.residual(z, y0.copy())) < TOL


# ---- convection in the Kronecker kernel (cell-wise constant velocity): volume term -u b.grad psi and the upwind face
# flux with the velocity of the larger-index cell (convectiondiffusiondg.hh:178-187, 426-448, 797-822, 860) ----------
FAST_B_CASES = [
    dict(cells=(8, 4, 4), a="scalar", with_b=True), dict(cells=(16, 8, 8), a="diagonal", with_b=True, with_c=True),
    dict(cells=(6, 5, 3), a="scalar", with_b=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(10, 9, 7), a="diagonal", with_b=True, with_c=True, bc="mixed"),
    dict(cells=(12, 4, 5), a="scalar", with_b=True, method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF, alpha=1.0),
    dict(cells=(2, 1, 1), a="identity", with_b=True),
]


@pytest.mark.parametrize("case", FAST_B_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_fast_kernel_with_convection_matches_oracle(cuda_lib, case):
    spec = dg_problem(degree=2, kernel=abi.KERNEL_FAST, with_f=True, **case)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    want = orc.jacobian_apply(z)
    y = go.apply(z, np.full(spec.num_dofs, np.nan))
    assert go.last_kernel() == "dg_fast_q2_3d"
    assert rel_err(y, want) < TOL
    y0 = mt_vector(spec.num_dofs, seed=7)
    assert rel_err(go.jacobian_apply(z, y0.copy()), want + y0) < TOL
    assert rel_err(go.residual(z, y0.copy()), orc.residual(z, y0.copy())) < TOL       # J z + cached R(0)
    assert go.last_kernel() == "dg_fast_q2_3d+r0"


def _outflow_problem(cells, b, kernel):
    """constant velocity b > 0: inflow faces (lower sides) Dirichlet, outflow faces (upper sides) Outflow"""
    spec = dg_problem(cells, degree=2, a="scalar", with_c=True, with_f=True, kernel=kernel)
    nc = spec.ncells
    bct = np.full(spec.num_boundary_faces, abi.BC_DIRICHLET, dtype=np.int8)
    for d in range(3):
        o = spec.boundary_face_offset(d, 1)
        bct[o:o + nc // cells[d]] = abi.BC_OUTFLOW
    rng = np.random.default_rng(3)
    return spec.replace(b=np.tile(np.asarray(b, dtype=float), (nc, 1)), bctype=bct,
                        g=rng.standard_normal((spec.num_boundary_faces, spec.nfq)),
                        o=rng.standard_normal((spec.num_boundary_faces, spec.nfq)))


def test_fast_kernel_outflow_faces_and_the_inflow_exception(cuda_lib):
    from pdelab_b200.capi import PDELabError
    spec = _outflow_problem((8, 6, 4), (1.0, 0.5, 0.25), abi.KERNEL_FAST)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    assert rel_err(go.apply(z, np.zeros_like(z)), orc.jacobian_apply(z)) < TOL
    assert go.last_kernel() == "dg_fast_q2_3d"
    assert rel_err(go.residual(z, np.zeros_like(z)), orc.residual(z)) < TOL
    # the same faces with the velocity reversed: "Outflow boundary condition on inflow!" (:802-806)
    bad = _outflow_problem((8, 6, 4), (-1.0, 0.5, 0.25), abi.KERNEL_FAST)
    with pytest.raises(PDELabError, match="Outflow boundary condition on inflow"):
        _ops(bad)[0].apply(z, np.zeros_like(z))


SMALL_B_CASES = [
    dict(cells=(9, 7), degree=1, a="scalar", with_b=True), dict(cells=(6, 5), degree=2, a="diagonal", with_b=True, with_c=True),
    dict(cells=(8, 6), degree=2, a="scalar", with_b=True, bc="mixed"),
    dict(cells=(5, 4, 3), degree=1, a="diagonal", with_b=True, with_c=True, extent=(1.0, 0.7, 1.3)),
    dict(cells=(7, 5), degree=1, a="identity", with_b=True, method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF),
]


@pytest.mark.parametrize("case", SMALL_B_CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_small_kernel_with_convection_matches_oracle(cuda_lib, case):
    spec = dg_problem(kernel=abi.KERNEL_FAST, with_f=True, **case)
    go, orc = _ops(spec)
    z = mt_vector(spec.num_dofs)
    y0 = mt_vector(spec.num_dofs, seed=7)
    assert rel_err(go.apply(z, np.full(spec.num_dofs, np.nan)), orc.jacobian_apply(z)) < TOL
    assert go.last_kernel() == "dg_small"
    assert rel_err(go.jacobian_apply(z, y0.copy()), orc.jacobian_apply(z, y0.copy())) < TOL
    assert rel_err(go.residual(z, y0.copy()), orc.residual(z, y0.copy())) < TOL

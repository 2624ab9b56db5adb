"""Parity of the device OneStepGridOperator (csrc/onestep.cu, the fused stage operator) with the one-step
engines restated on the CPU oracle (tests/onestep_oracle.py: the two operators evaluated separately and combined
with the engine weights, the reference's order of operations).  -m gpu, through the C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from pdelab_b200 import abi, onestep as osm
from problems import dg_problem, fem_problem, mt_vector, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = {
    "dg_fast_scalar": lambda: dg_problem((8, 4, 4), degree=2, a="scalar", with_f=True, bc="dirichlet_g"),
    "dg_fast_diag_mixed": lambda: dg_problem((10, 9, 7), degree=2, a="diagonal", bc="mixed", with_c=True, with_f=True),
    "dg_generic_full_b": lambda: dg_problem((5, 4), degree=2, a="full", with_b=True, with_c=True, with_f=True, bc="mixed"),
    "dg_weights_off": lambda: dg_problem((6, 4, 3), degree=2, a="scalar", method=abi.DG_NIPG, weights=abi.DG_WEIGHTS_OFF,
                                         with_f=True, bc="dirichlet_g"),
    "dg_kron_k3": lambda: dg_problem((4, 2, 2), degree=3, a="scalar", with_f=True),
    "dg_small_2d_identity": lambda: dg_problem((8, 6), degree=1, a="identity", with_f=True, bc="dirichlet_g"),
    "fem_q1_2d": lambda: fem_problem((9, 7), degree=1, a="scalar"),
    "fem_q2_3d": lambda: fem_problem((4, 3, 3), degree=2, a="diagonal", with_c=True),
    "fem_q1_full_b_mixed": lambda: fem_problem((6, 5), degree=1, a="full", with_b=True, bc="mixed"),
}
METHODS = {"implicit_euler": osm.ImplicitEulerParameter, "theta_half": lambda: osm.OneStepThetaParameter(0.5),
           "alexander2": osm.Alexander2Parameter, "fractional_step": osm.FractionalStepParameter,
           "alexander3": osm.Alexander3Parameter}


def _pair(spec0, scaling=1.0, **kw):
    from onestep_oracle import OneStepOracle
    from pdelab_b200.capi import GridOperator
    spec1 = osm.l2_spec(spec0, scaling)
    go0, go1 = GridOperator(spec0), GridOperator(spec1)
    return osm.OneStepGridOperator(go0, go1, **kw), OneStepOracle(spec0, spec1), (go0, go1)


def _device_matrix(igo, x):
    rowptr, colidx = igo.fill_pattern()
    values = igo.jacobian(x, np.zeros(colidx.size))
    n = rowptr.size - 1
    return sp.csr_matrix((values, colidx.astype(np.int64), rowptr.astype(np.int64)), shape=(n, n))


@pytest.mark.parametrize("mname", list(METHODS))
@pytest.mark.parametrize("cname", list(CASES))
def test_stage_operator_matches_one_step_engines(cuda_lib, cname, mname):
    spec0 = CASES[cname]()
    method = METHODS[mname]()
    igo, orc, _keep = _pair(spec0, scaling=1.5)
    n = spec0.num_dofs
    time, dt = 0.25, 0.0625
    igo.preStep(method, time, dt)
    orc.preStep(method, time, dt)
    xs = [mt_vector(n, seed=11 + i) - 0.5 for i in range(method.s())]
    for r in range(1, method.s() + 1):
        igo.preStage(r, xs[:r])
        orc.preStage(r, xs[:r])
        assert abs(igo.timeAtStage(r) - (time + method.d(r) * dt)) < 1e-15
        const = igo.const_residual(np.zeros(n))
        scale = max(np.abs(orc.const).max(), 1e-300)
        assert np.abs(const - orc.const).max() / scale < TOL, ("const_residual", r)
        x = mt_vector(n, seed=31 + r)
        r0 = mt_vector(n, seed=41 + r)
        assert rel_err(igo.residual(x, r0.copy()), orc.residual(x, r0.copy())) < TOL, ("residual", r)
        z = mt_vector(n, seed=51 + r)
        want = orc.jacobian_apply(z)
        assert rel_err(igo.apply(z, np.full(n, np.nan)), want) < TOL, ("apply", r)
        y0 = mt_vector(n, seed=61 + r)
        con = orc.con
        wacc = want + y0
        wacc[con] = 0.0
        assert rel_err(igo.jacobian_apply(z, y0.copy()), wacc) < TOL, ("jacobian_apply", r)
        if n <= 1500:
            M, Mo = _device_matrix(igo, x), orc.matrix()
            assert abs(M - Mo).max() / abs(Mo).max() < TOL, ("jacobian", r)


@pytest.mark.parametrize("mode", [osm.OneStepGridOperator.DivideOperator1ByDT, osm.OneStepGridOperator.MultiplyOperator0ByDT])
def test_dt_assembling_modes(cuda_lib, mode):
    """divideMassTermByDeltaT / multiplySpatialTermByDeltaT (onestep.hh:78-91, localassembler.hh:101-130)."""
    spec0 = CASES["dg_fast_scalar"]()
    igo, orc, _keep = _pair(spec0)
    method = osm.Alexander2Parameter()
    if mode == osm.OneStepGridOperator.DivideOperator1ByDT:
        igo.setMethod(method)
        igo.divideMassTermByDeltaT()
    orc.dt_mode = mode
    n = spec0.num_dofs
    xs = [mt_vector(n, seed=3), mt_vector(n, seed=4)]
    igo.preStep(method, 0.0, 0.01)
    orc.preStep(method, 0.0, 0.01)
    igo.preStage(2, xs)
    orc.preStage(2, xs)
    x = mt_vector(n, seed=5)
    assert rel_err(igo.residual(x, np.zeros(n)), orc.residual(x)) < TOL


def test_time_dependent_coefficients_are_resampled_per_stage(cuda_lib):
    """prestageengine.hh:208-211 evaluates every earlier stage at its own time t + d_i dt."""
    base = dg_problem((6, 4, 4), degree=2, a="scalar", with_f=True, bc="dirichlet_g")
    f, g = base.arrays["f"], base.arrays["g"]

    def at(t):
        return base.replace(f=(1.0 + t) * f, g=np.cos(t) * g)

    from onestep_oracle import OneStepOracle
    from pdelab_b200.capi import GridOperator
    spec1 = osm.l2_spec(base)
    go0, go1 = GridOperator(at(0.0)), GridOperator(spec1)
    igo = osm.OneStepGridOperator(go0, go1, time_dependent=lambda t: dict(f=(1.0 + t) * f, g=np.cos(t) * g))
    orc = OneStepOracle(base, spec1, spec0_at=at)
    method = osm.FractionalStepParameter()
    n = base.num_dofs
    xs = [mt_vector(n, seed=7 + i) for i in range(3)]
    igo.preStep(method, 0.5, 0.125)
    orc.preStep(method, 0.5, 0.125)
    for r in (1, 2, 3):
        igo.preStage(r, xs[:r])
        orc.preStage(r, xs[:r])
        x = mt_vector(n, seed=20 + r)
        assert rel_err(igo.residual(x, np.zeros(n)), orc.residual(x)) < TOL, r


def test_device_vectors_and_stage_operator_handle(cuda_lib):
    """torch CUDA vectors are used in place; the fused stage operator is an ordinary handle (block Jacobi, solve)."""
    import torch
    spec0 = dg_problem((8, 6, 4), degree=2, a="scalar", with_f=True)
    igo, orc, _keep = _pair(spec0)
    method = osm.ImplicitEulerParameter()
    n = spec0.num_dofs
    x0 = mt_vector(n, seed=2)
    igo.preStep(method, 0.0, 0.02)
    orc.preStep(method, 0.0, 0.02)
    igo.preStage(1, [torch.from_numpy(x0).cuda()])
    orc.preStage(1, [x0])
    xd = torch.from_numpy(x0).cuda()
    rd = torch.zeros_like(xd)
    igo.residual(xd, rd)
    torch.cuda.synchronize()
    want = orc.residual(x0)
    assert rel_err(rd.cpu().numpy(), want) < TOL
    st = igo.stage_operator()
    assert st.last_kernel().startswith("dg_fast_q2_3d")
    z = torch.zeros_like(xd)
    res = st.solve(z, rd.clone(), 1e-10, solver=abi.SOLVER_CG, precond=abi.PRECOND_BLOCK_JACOBI)
    assert res["converged"] == 1
    jz = orc.jacobian_apply(z.cpu().numpy())
    assert np.abs(jz - want).max() / np.abs(want).max() < 1e-8


def test_errors_like_the_reference(cuda_lib):
    from pdelab_b200.capi import GridOperator, PDELabError
    spec0 = dg_problem((4, 4), degree=1, a="identity")
    igo, _orc, _keep = _pair(spec0)
    n = spec0.num_dofs
    with pytest.raises(PDELabError, match="no time-stepping method"):
        igo.preStage(1, [np.zeros(n)])
    igo.preStep(osm.ImplicitEulerParameter(), 0.0, 0.1)
    with pytest.raises(PDELabError, match="no stage selected"):
        igo.residual(np.zeros(n), np.zeros(n))
    igo.preStep(osm.HeunParameter(), 0.0, 0.1)           # explicit method on the implicit operator
    igo.preStage(1, [np.zeros(n)])
    with pytest.raises(PDELabError, match="explicit mode"):   # onestep.hh:143-144
        igo.residual(np.zeros(n), np.zeros(n))
    with pytest.raises(PDELabError, match="explicit mode"):   # onestep.hh:78-84
        igo.divideMassTermByDeltaT()
    other = GridOperator(osm.l2_spec(dg_problem((4, 5), degree=1, a="identity")))
    with pytest.raises(PDELabError, match="same grid"):
        osm.OneStepGridOperator(_keep[0], other)


@pytest.mark.parametrize("solver,precond,matrix_free", [(abi.SOLVER_BICGSTAB, abi.PRECOND_NONE, True),
                                                        (abi.SOLVER_CG, abi.PRECOND_BLOCK_JACOBI, True),
                                                        (abi.SOLVER_CG, abi.PRECOND_JACOBI, False)])
def test_reference_instationary_dg_test(cuda_lib, solver, precond, matrix_free):
    """test/testinstationaryfastdgassembler.cc on the device: QkDG k=1 on 8x8, SIPG alpha=2, L2, Alexander2, one step
    dt = 0.1 from the interpolated stationary solution; squared L2 error <= 5e-6 (:197), and the same end state as the
    oracle's time step with a direct stage solver."""
    from manufactured import l2_error_squared, node_coordinates
    from pdelab_b200.capi import GridOperator
    from test_onestep_oracle import heat_problem, oracle_time_step, u_exact
    spec0 = heat_problem()
    go0, go1 = GridOperator(spec0), GridOperator(osm.l2_spec(spec0))
    igo = osm.OneStepGridOperator(go0, go1)
    method = osm.Alexander2Parameter()
    stepper = osm.OneStepMethod(method, igo, reduction=1e-10, solver=solver, precond=precond, matrix_free=matrix_free)
    x = u_exact(node_coordinates(spec0))
    xnew = np.zeros_like(x)
    time, dt, T = 0.0, 0.1, 0.1
    while time < T - 1e-10:
        stepper.apply(time, dt, x, xnew)
        x, xnew = xnew, np.zeros_like(x)
        time += dt
    err2 = l2_error_squared(spec0, x, u_exact, npts=7)
    assert err2 <= 5e-6, err2
    want = oracle_time_step(spec0, method, u_exact(node_coordinates(spec0)), 0.0, 0.1)
    assert rel_err(x, want) < 1e-8
    assert stepper.linear_solver_iterations > 0


def test_fused_stage_equals_separate_operators_at_size(cuda_lib):
    """Size-independent property at 64^3 cells (7.1 M DOFs): the fused stage apply equals
    b_rr dt J0 z + J1 z evaluated with the two operators separately on the device."""
    import torch
    spec0 = dg_problem((64, 64, 64), degree=2, a="scalar")
    from pdelab_b200.capi import GridOperator
    go0, go1 = GridOperator(spec0), GridOperator(osm.l2_spec(spec0))
    igo = osm.OneStepGridOperator(go0, go1)
    method = osm.Alexander2Parameter()
    dt = 1e-3
    igo.preStep(method, 0.0, dt)
    n = spec0.num_dofs
    z = torch.rand(n, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    igo.preStage(1, [z])
    y = torch.empty_like(z)
    igo.apply(z, y)
    y0, y1 = torch.empty_like(z), torch.empty_like(z)
    go0.apply(z, y0)
    go1.apply(z, y1)
    torch.cuda.synchronize()
    want = method.b(1, 1) * dt * y0 + y1
    assert float((y - want).abs().max() / want.abs().max()) < TOL
    assert igo.stage_operator().last_kernel() == "dg_fast_q2_3d"


EXPLICIT = {"explicit_euler": osm.ExplicitEulerParameter, "heun": osm.HeunParameter, "shu3": osm.Shu3Parameter,
            "rk4": osm.RK4Parameter}


@pytest.mark.parametrize("mname", list(EXPLICIT))
@pytest.mark.parametrize("cname", ["dg_fast_scalar", "dg_small_2d_identity", "dg_kron_k3", "dg_generic_full_b"])
def test_explicit_stages_match_the_oracle(cuda_lib, cname, mname):
    """ExplicitOneStepMethod stage by stage (exact block mass inverse for k <= 2, CG on the mass operator for k = 3)
    against the engines restated on the oracle with a sparse direct mass solve."""
    from onestep_oracle import explicit_stage
    spec0 = CASES[cname]()
    spec1 = osm.l2_spec(spec0, 1.5)
    method = EXPLICIT[mname]()
    igo, _orc, _keep = _pair(spec0, scaling=1.5)
    n = spec0.num_dofs
    time, dt = 0.25, 1e-4
    igo.preStep(method, time, dt)
    xs = [mt_vector(n, seed=11 + i) - 0.5 for i in range(method.s())]
    for r in range(1, method.s() + 1):
        got = igo.explicit_stage(r, xs[:r], np.zeros(n))
        want = explicit_stage(spec0, spec1, method, r, time, dt, xs[:r])
        assert rel_err(got, want) < 1e-10, r
    with pytest.raises(Exception, match="explicit mode"):
        igo.residual(np.zeros(n), np.zeros(n))
    with pytest.raises(Exception, match="implicit scheme"):
        osm.ExplicitOneStepMethod(osm.Alexander2Parameter(), igo)


@pytest.mark.parametrize("mname", ["heun", "shu3", "rk4"])
def test_explicit_stages_resample_time_dependent_coefficients(cuda_lib, mname):
    """Every R0(x_i) of an explicit stage is evaluated with the coefficients at t + d_i dt (the explicit engine
    delegates to the pre-stage engine, prestageengine.hh:208-211) — f(t) and g(t) vary strongly over one step."""
    from onestep_oracle import explicit_stage
    from pdelab_b200.capi import GridOperator
    base = dg_problem((6, 4, 4), degree=2, a="scalar", with_f=True, bc="dirichlet_g")
    f, g = base.arrays["f"], base.arrays["g"]

    def at(t):
        return base.replace(f=(1.0 + 40.0 * t) * f, g=np.cos(30.0 * t) * g)

    spec1 = osm.l2_spec(base)
    go0, go1 = GridOperator(at(0.0)), GridOperator(spec1)
    igo = osm.OneStepGridOperator(go0, go1, time_dependent=lambda t: dict(f=(1.0 + 40.0 * t) * f, g=np.cos(30.0 * t) * g))
    method = EXPLICIT[mname]()
    n = base.num_dofs
    time, dt = 0.25, 0.05
    igo.preStep(method, time, dt)
    xs = [mt_vector(n, seed=11 + i) - 0.5 for i in range(method.s())]
    for r in range(1, method.s() + 1):
        got = igo.explicit_stage(r, xs[:r], np.zeros(n))
        want = explicit_stage(base, spec1, method, r, time, dt, xs[:r], spec0_at=at)
        frozen = explicit_stage(at(time), spec1, method, r, time, dt, xs[:r])
        assert rel_err(got, want) < 1e-10, r
        if r > 1 and any(method.d(i) != 0.0 and abs(method.b(r, i)) > 1e-6 for i in range(r)):
            assert rel_err(frozen, want) > 1e-3, r   # the test does distinguish the two


def test_explicit_then_implicit_method_keeps_the_dt_assembling_mode(cuda_lib):
    """An explicit method used once must not leave DoNotAssembleDT behind for a later implicit method (the reference fixes
    `implicit` per grid-operator type, onestep.hh:74-75; here one handle may see both)."""
    spec0 = CASES["dg_fast_scalar"]()
    igo, orc, _keep = _pair(spec0)
    n = spec0.num_dofs
    xs = [mt_vector(n, seed=3), mt_vector(n, seed=4)]
    igo.preStep(osm.HeunParameter(), 0.0, 0.05)
    igo.explicit_stage(1, xs[:1], np.zeros(n))
    method = osm.Alexander2Parameter()
    igo.preStep(method, 0.0, 0.05)
    orc.preStep(method, 0.0, 0.05)
    igo.preStage(1, xs[:1])
    orc.preStage(1, xs[:1])
    x = mt_vector(n, seed=5)
    assert rel_err(igo.residual(x, np.zeros(n)), orc.residual(x)) < TOL


def test_explicit_rk4_keeps_the_stationary_solution(cuda_lib):
    """The heat problem of testinstationaryfastdgassembler.cc stepped explicitly (RK4, dt well inside the diffusive
    stability limit of the 8x8 k=1 DG grid): 20 steps from the interpolated stationary solution stay within the
    reference test's bound."""
    from manufactured import l2_error_squared, node_coordinates
    from pdelab_b200.capi import GridOperator
    from test_onestep_oracle import heat_problem, u_exact
    spec0 = heat_problem()
    go0, go1 = GridOperator(spec0), GridOperator(osm.l2_spec(spec0))
    igo = osm.OneStepGridOperator(go0, go1)
    stepper = osm.ExplicitOneStepMethod(osm.RK4Parameter(), igo)
    x = u_exact(node_coordinates(spec0))
    time, dt = 0.0, 2e-5
    for _ in range(20):
        xnew = np.zeros_like(x)
        stepper.apply(time, dt, x, xnew)
        x, time = xnew, time + dt
    assert np.all(np.isfinite(x))
    assert l2_error_squared(spec0, x, u_exact, npts=7) <= 5e-6


def test_reference_testinstationary_q2(cuda_lib):
    """test/testinstationary.cc on the device: conforming Q2 32x32, L2, implicit Euler, dt = 0.1 to T = 1 from the
    interpolated Dirichlet extension, squared L2 error <= 1e-7 (:193-196); testGridOperatorInterface (:7-16) on the
    one-step operator; same end state as the oracle's time loop with a direct stage solver."""
    import scipy.sparse.linalg as spla
    from manufactured import l2_error_squared, node_coordinates
    from onestep_oracle import OneStepOracle
    from pdelab_b200.capi import GridOperator
    from test_onestep_oracle import fem_heat_problem, u_centre
    spec0 = fem_heat_problem()
    spec1 = osm.l2_spec(spec0)
    go0, go1 = GridOperator(spec0), GridOperator(spec1)
    igo = osm.OneStepGridOperator(go0, go1)
    method = osm.OneStepThetaParameter(1.0)
    stepper = osm.OneStepMethod(method, igo, reduction=1e-12, solver=abi.SOLVER_CG, precond=abi.PRECOND_JACOBI)
    x = u_centre(node_coordinates(spec0))
    xo = x.copy()
    orc = OneStepOracle(spec0, spec1)
    time, dt = 0.0, 0.1
    while time < 1.0 - 1e-8:
        xnew = x.copy()
        stepper.apply(time, dt, x, xnew)
        x = xnew
        orc.preStep(method, time, dt)
        orc.preStage(1, [xo])
        xo = xo - spla.spsolve(orc.matrix().tocsc(), orc.residual(xo.copy()))
        time += dt
    assert l2_error_squared(spec0, x, u_centre, npts=6) <= 1e-7
    assert rel_err(x, xo) < 1e-9
    # testGridOperatorInterface: residual, jacobian, jacobian_apply of the one-step operator on u = 0
    n = spec0.num_dofs
    u = np.zeros(n)
    r = igo.residual(u, np.zeros(n))
    rowptr, colidx = igo.fill_pattern()
    vals = igo.jacobian(u, np.zeros(colidx.size))
    y = igo.jacobian_apply(u, np.zeros(n))
    assert rel_err(r, orc.residual(u)) < TOL and np.all(y == 0.0)
    M = sp.csr_matrix((vals, colidx.astype(np.int64), rowptr.astype(np.int64)), shape=(n, n))
    Mo = orc.matrix()
    assert abs(M - Mo).max() / abs(Mo).max() < TOL


def test_reference_time_dependent_boundary(cuda_lib):
    """test/testtimedependentboundary_ovlpqk.cc (one rank) on the device: Q1 32 x 32, f = 1, g = t, implicit Euler with
    dt = 0.01 to T = 1, boundary values interpolated at every stage; sum (v - T)^2 <= 1e-18 (:206-216)."""
    from pdelab_b200.capi import GridOperator
    from test_onestep_oracle import time_boundary_problem
    spec0 = time_boundary_problem()
    go0, go1 = GridOperator(spec0), GridOperator(osm.l2_spec(spec0))
    igo = osm.OneStepGridOperator(go0, go1)
    stepper = osm.OneStepMethod(osm.OneStepThetaParameter(1.0), igo, reduction=1e-11, solver=abi.SOLVER_CG,
                                precond=abi.PRECOND_JACOBI)
    n = spec0.num_dofs
    x = np.zeros(n)
    time, dt = 0.0, 0.01
    while time < 1.0 - 1e-8:
        xnew = x.copy()
        stepper.apply(time, dt, x, xnew, f=lambda t: np.full(n, t))
        x, time = xnew, time + dt
    assert float(np.sum((x - time) ** 2)) <= 1e-18

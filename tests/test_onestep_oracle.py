"""CPU checks of the instationary layer (no GPU): the time-stepping parameter tables and the one-step oracle
restatement (tests/onestep_oracle.py), pinned against the reference's own instationary test
(test/testinstationaryfastdgassembler.cc: QkDG k=1 on 8x8, SIPG alpha=2, L2 mass operator, Alexander2, one step
dt = 0.1 from the interpolated stationary solution, squared L2 error <= 5e-6)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from manufactured import l2_error_squared, node_coordinates, sample_data
from onestep_oracle import OneStepOracle
from pdelab_b200 import abi, onestep as osm

METHODS = [osm.ImplicitEulerParameter, lambda: osm.OneStepThetaParameter(0.5), osm.Alexander2Parameter,
           osm.FractionalStepParameter, osm.Alexander3Parameter, osm.ExplicitEulerParameter, osm.HeunParameter,
           osm.Shu3Parameter, osm.RK4Parameter]


@pytest.mark.parametrize("make", METHODS)
def test_method_tables_are_consistent(make):
    """sum_i a_ri = 0 and sum_i a_ri d_i = sum_i b_ri (exact for u(t) = t) — holds for every table of
    instationary/onestepparameter.hh."""
    m = make()
    for r in range(1, m.s() + 1):
        a = [m.a(r, i) for i in range(r + 1)]
        b = [m.b(r, i) for i in range(r + 1)]
        d = [m.d(i) for i in range(r + 1)]
        assert abs(sum(a)) < 1e-14
        assert abs(sum(x * y for x, y in zip(a, d)) - sum(b)) < 1e-14
        assert m.a(r, r) == 1.0          # the stage weight of the temporal operator is dt_factor1 alone
    assert m.d(0) == 0.0 and m.d(m.s()) == 1.0
    assert m.implicit() == any(abs(m.b(r, r)) > 0 for r in range(1, m.s() + 1))


def u_exact(X):
    return np.exp(-np.sum(X * X, axis=1))


def heat_problem(cells=(8, 8), degree=1, space=abi.SPACE_QKDG):
    """ParameterA of test/testinstationaryfastdgassembler.cc:17-104: A = I, f = (2 d - 4 |x|^2) exp(-|x|^2),
    Dirichlet g = exp(-|x|^2)."""
    dim = len(cells)
    spec = abi.ProblemSpec(cells, space=space, degree=degree, method=abi.DG_SIPG, weights=abi.DG_WEIGHTS_ON, alpha=2.0)
    return sample_data(spec, u_exact, lambda X: (2.0 * dim - 4.0 * np.sum(X * X, axis=1)) * u_exact(X))


def oracle_time_step(spec0, method, x, time, dt):
    """OneStepMethod::apply (instationary/implicitonestep.hh:122-262) with a direct stage solver."""
    os_ = OneStepOracle(spec0, osm.l2_spec(spec0))
    xs = [x]
    os_.preStep(method, time, dt)
    for r in range(1, method.s() + 1):
        os_.preStage(r, xs)
        xr = xs[r - 1].copy()
        res = os_.residual(xr)
        xr -= spla.spsolve(os_.matrix().tocsc(), res)
        xs.append(xr)
    return xs[-1]


def test_reference_instationary_dg_test_on_the_oracle():
    spec0 = heat_problem()
    x = u_exact(node_coordinates(spec0))                    # interpolate(g, gfs, x)
    x = oracle_time_step(spec0, osm.Alexander2Parameter(), x, 0.0, 0.1)
    err2 = l2_error_squared(spec0, x, u_exact, npts=7)       # integrateGridFunction(..., 12)
    assert err2 <= 5e-6, err2                                # testinstationaryfastdgassembler.cc:197


def test_one_step_residual_is_affine_and_consistent():
    """residual(x) = J x + residual(0) and J from matrix() == jacobian_apply, for a stage with a constant part."""
    spec0 = heat_problem((4, 3))
    os_ = OneStepOracle(spec0, osm.l2_spec(spec0, 2.0))
    rng = np.random.default_rng(1)
    xs = [rng.random(spec0.num_dofs) for _ in range(3)]
    os_.preStep(osm.Alexander3Parameter(), 0.3, 0.05)
    os_.preStage(3, xs)
    x = rng.random(spec0.num_dofs)
    M = os_.matrix()
    assert np.allclose(os_.residual(x), M @ x + os_.residual(np.zeros_like(x)), rtol=0, atol=1e-11)
    assert np.allclose(os_.jacobian_apply(x), M @ x, rtol=0, atol=1e-11)
    assert np.abs(os_.const).max() > 0


def u_centre(X):
    return np.exp(-np.sum((X - 0.5) ** 2, axis=1))


def fem_heat_problem(cells=(32, 32), degree=2):
    """PoissonProblem of test/testinstationary.cc:19-52: f = 4 (1 - c) exp(-c), c = |x - 1/2|^2, Dirichlet g = exp(-c);
    conforming Q2 on the 4x4 grid refined three times (:62-74)."""
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=degree)
    return sample_data(spec, u_centre, lambda X: 4.0 * (1.0 - np.sum((X - 0.5) ** 2, axis=1)) * u_centre(X))


def test_reference_testinstationary_on_the_oracle():
    """test/testinstationary.cc: conforming Q2 32x32, L2 temporal operator, implicit Euler (OneStepThetaParameter(1.0)),
    dt = 0.1 up to T = 1 from the interpolated Dirichlet extension; squared L2 error <= 1e-7 (:193-196); and the
    grid-operator interface calls of testGridOperatorInterface (:7-16) on the one-step operator."""
    spec0 = fem_heat_problem()
    x = u_centre(node_coordinates(spec0))
    method = osm.OneStepThetaParameter(1.0)
    time, dt = 0.0, 0.1
    os_ = OneStepOracle(spec0, osm.l2_spec(spec0))
    while time < 1.0 - 1e-8:
        os_.preStep(method, time, dt)
        os_.preStage(1, [x])
        xr = x.copy()                       # interpolate(...) leaves the (time-independent) Dirichlet values in place
        xr -= spla.spsolve(os_.matrix().tocsc(), os_.residual(xr))
        x, time = xr, time + dt
    assert l2_error_squared(spec0, x, u_centre, npts=6) <= 1e-7
    u = np.zeros(spec0.num_dofs)
    assert np.all(np.isfinite(os_.residual(u))) and np.all(np.isfinite(os_.jacobian_apply(u)))
    assert os_.matrix().shape == (spec0.num_dofs,) * 2


def time_boundary_problem(cells=(32, 32)):
    """ConvectionDiffusionModelProblem of test/testtimedependentboundary_ovlpqk.cc:25-115: A = I, f = 1, Dirichlet
    g(x, t) = t everywhere, conforming Q1 on 32 x 32 (:228-247); the solution is u = t."""
    spec = abi.ProblemSpec(cells, space=abi.SPACE_QK, degree=1)
    return spec.replace(f=np.ones((spec.ncells, spec.nq)))


def test_reference_time_dependent_boundary_on_the_oracle():
    """test/testtimedependentboundary_ovlpqk.cc (sequential): implicit Euler, dt = 0.01 to T = 1, the Dirichlet values
    g = t interpolated at every stage (OneStepMethod::apply with f, implicitonestep.hh:264-400); sum (v - T)^2 <= 1e-18
    (:206-216)."""
    spec0 = time_boundary_problem()
    os_ = OneStepOracle(spec0, osm.l2_spec(spec0))
    method = osm.OneStepThetaParameter(1.0)
    n = spec0.num_dofs
    x = np.zeros(n)                                   # interpolate(gf, gfs, zs) at t = 0
    time, dt = 0.0, 0.01
    lu = None
    while time < 1.0 - 1e-8:
        os_.preStep(method, time, dt)
        os_.preStage(1, [x])
        xr = x.copy()
        xr[os_.con] = time + dt                       # igos.interpolate(r, init_guess, gf, x[r])
        if lu is None:
            lu = spla.splu(os_.matrix().tocsc())      # the stage matrix does not change from step to step
        xr -= lu.solve(os_.residual(xr))
        x, time = xr, time + dt
    assert float(np.sum((x - time) ** 2)) <= 1e-18


class _OracleBackedIGO(osm.OneStepGridOperator):
    """The host-side mirror (OneStepGridOperator / OneStepMethod of pdelab_b200.onestep) with the device calls replaced
    by the oracle restatement, so that the HOST logic (stage loop, initial guesses, boundary interpolation) runs in
    the CPU suite.  Only the methods that cross the C ABI are overridden."""

    def __init__(self, spec0, spec1):  # noqa: D401 — no library, no device
        self.orc = OneStepOracle(spec0, spec1)
        self._time_dependent = None
        self._sampled_time = None

        class _GO0:
            def constrained_dofs(_self):
                return self.orc.con.astype(np.uint64)
        self.go0 = _GO0()

    def __del__(self):
        pass

    def preStep(self, method, time, dt):
        self._m, self._t, self._dt = method, time, dt
        self.orc.preStep(method, time, dt)

    def timeAtStage(self, stage):
        return self._t + self._m.d(stage) * self._dt

    def preStage(self, stage, x):
        self.orc.preStage(stage, [np.asarray(v) for v in x[:stage]])

    def solve_stationary(self, x, **kw):
        x -= spla.spsolve(self.orc.matrix().tocsc(), self.orc.residual(x.copy()))
        return dict(converged=1, iterations=1)


def test_host_mirror_stage_loop_with_boundary_interpolation():
    """OneStepMethod.apply(time, dt, xold, xnew, f) of the Python mirror (implicitonestep.hh:264-400) on the
    time-dependent boundary problem, two-stage method: the interpolated Dirichlet values follow g = t at the stage times
    and the exact solution u = t is reproduced."""
    spec0 = time_boundary_problem((8, 8))
    igo = _OracleBackedIGO(spec0, osm.l2_spec(spec0))
    stepper = osm.OneStepMethod(osm.Alexander2Parameter(), igo)
    n = spec0.num_dofs
    x = np.zeros(n)
    time, dt = 0.0, 0.05
    seen = []
    for _ in range(4):
        xnew = x.copy()
        stepper.apply(time, dt, x, xnew, f=lambda t: (seen.append(t), np.full(n, t))[1])
        x, time = xnew, time + dt
    assert float(np.sum((x - time) ** 2)) <= 1e-20
    m = osm.Alexander2Parameter()
    assert np.allclose(seen[:2], [m.d(1) * dt, dt]) and len(seen) == 8
    assert stepper.linear_solver_iterations == 8 and stepper.step == 5

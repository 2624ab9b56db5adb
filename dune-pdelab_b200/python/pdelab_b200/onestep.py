"""Host mirror of PDELab's instationary layer over the C ABI (pdb200_onestep_*).

  * time-stepping parameter classes  (dune/pdelab/instationary/onestepparameter.hh:43-698)
  * OneStepGridOperator              (dune/pdelab/gridoperator/onestep.hh:30-308)
  * OneStepMethod                    (dune/pdelab/instationary/implicitonestep.hh:37-439, linear problems:
                                      the stage solver is StationaryLinearProblemSolver::apply on the device)

Same method names and argument meaning as the reference; errors are raised as PDELabError.  On the device a
stage is one fused operator (csrc/onestep.cu), so `stage_operator()` returns an ordinary GridOperator view on
which solve / fill_pattern / jacobian / block_jacobi_apply work unchanged.
"""
import ctypes as C
import math

import numpy as np

from . import abi
from .capi import GridOperator, PDELabError, SolveResult, _ptr, load_library


# ---- onestepparameter.hh ------------------------------------------------------------------------
class TimeSteppingParameterInterface:
    """a(r, i), b(r, i) for r in 1..s, i in 0..r; d(i) for i in 0..s (onestepparameter.hh:43-84)."""

    A = B = D = None
    _name = ""

    def implicit(self):
        raise NotImplementedError

    def s(self):
        return len(self.A)

    def a(self, r, i):
        return self.A[r - 1][i]

    def b(self, r, i):
        return self.B[r - 1][i]

    def d(self, i):
        return self.D[i]

    def name(self):
        return self._name


class OneStepThetaParameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:88-151"""
    _name = "one step theta"

    def __init__(self, theta):
        self.theta = float(theta)
        self.D = [0.0, 1.0]
        self.A = [[-1.0, 1.0]]
        self.B = [[1.0 - self.theta, self.theta]]

    def implicit(self):
        return self.theta > 0.0


class ExplicitEulerParameter(OneStepThetaParameter):
    _name = "explicit Euler"

    def __init__(self):
        super().__init__(0.0)


class ImplicitEulerParameter(OneStepThetaParameter):
    _name = "implicit Euler"

    def __init__(self):
        super().__init__(1.0)


class HeunParameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:213-280"""
    _name = "Heun"
    D = [0.0, 1.0, 1.0]
    A = [[-1.0, 1.0, 0.0], [-0.5, -0.5, 1.0]]
    B = [[1.0, 0.0, 0.0], [0.0, 0.5, 0.0]]

    def implicit(self):
        return False


class Shu3Parameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:286-357"""
    _name = "Shu's third order method"
    D = [0.0, 1.0, 0.5, 1.0]
    A = [[-1.0, 1.0, 0.0, 0.0], [-0.75, -0.25, 1.0, 0.0], [-1.0 / 3.0, 0.0, -2.0 / 3.0, 1.0]]
    B = [[1.0, 0.0, 0.0, 0.0], [0.0, 0.25, 0.0, 0.0], [0.0, 0.0, 2.0 / 3.0, 0.0]]

    def implicit(self):
        return False


class RK4Parameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:363-436"""
    _name = "RK4"
    D = [0.0, 0.5, 0.5, 1.0, 1.0]
    A = [[-1.0, 1.0, 0.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 0.0, 1.0]]
    B = [[0.5, 0.0, 0.0, 0.0, 0.0], [0.0, 0.5, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0, 0.0],
         [1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0, 0.0]]

    def implicit(self):
        return False


class Alexander2Parameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:444-510"""
    _name = "Alexander (order 2)"

    def __init__(self):
        al = 1.0 - 0.5 * math.sqrt(2.0)
        self.D = [0.0, al, 1.0]
        self.A = [[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]]
        self.B = [[0.0, al, 0.0], [0.0, 1.0 - al, al]]

    def implicit(self):
        return True


class FractionalStepParameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:521-598"""
    _name = "Fractional step theta"

    def __init__(self):
        theta = 1.0 - 0.5 * math.sqrt(2.0)
        thetap = 1.0 - 2.0 * theta
        alpha = 2.0 - math.sqrt(2.0)
        beta = 1.0 - alpha
        self.D = [0.0, theta, 1.0 - theta, 1.0]
        self.A = [[-1.0, 1.0, 0.0, 0.0], [0.0, -1.0, 1.0, 0.0], [0.0, 0.0, -1.0, 1.0]]
        self.B = [[beta * theta, alpha * theta, 0.0, 0.0], [0.0, alpha * thetap, alpha * theta, 0.0],
                  [0.0, 0.0, beta * theta, alpha * theta]]

    def implicit(self):
        return True


class Alexander3Parameter(TimeSteppingParameterInterface):
    """onestepparameter.hh:604-698"""
    _name = "Alexander (claims order 3)"

    def __init__(self):
        al = 0.4358665215
        for _ in range(10):  # the reference's Newton iteration for alpha (:613-621)
            al = al - (al * (al * al - 3.0 * (al - 0.5)) - 1.0 / 6.0) / (3.0 * al * (al - 2.0) + 1.5)
        tau2 = (1.0 + al) * 0.5
        b1 = -(6.0 * al * al - 16.0 * al + 1.0) * 0.25
        b2 = (6 * al * al - 20.0 * al + 5.0) * 0.25
        self.D = [0.0, al, tau2, 1.0]
        self.A = [[-1.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 1.0]]
        self.B = [[0.0, al, 0.0, 0.0], [0.0, tau2 - al, al, 0.0], [0.0, b1, b2, al]]

    def implicit(self):
        return True


# ---- gridoperator/onestep.hh ---------------------------------------------------------------------
class _StageOperatorView(GridOperator):
    """The fused operator of the current stage as a GridOperator (handle owned by the one-step operator)."""

    def __init__(self, lib, handle, spec):  # noqa: D401 — no pdb200_create here
        self.lib, self.spec, self._h = lib, spec, handle

    def __del__(self):
        self._h = None

    close = __del__


class OneStepGridOperator:
    """OneStepGridOperator<GO0, GO1, implicit> (gridoperator/onestep.hh:30-308).

    go0: spatial GridOperator, go1: temporal GridOperator (the L2 mass operator).  Both are referenced, not
    owned.  `time_dependent(t)` (optional call-back) re-samples go0's coefficient arrays at time t and returns
    them as a dict for GridOperator.update_coefficients — the mirror of lop.setTime(t) -> param.setTime(t).
    """

    DivideOperator1ByDT, MultiplyOperator0ByDT, DoNotAssembleDT = 0, 1, 2

    def __init__(self, go0: GridOperator, go1: GridOperator, time_dependent=None):
        self.lib = load_library()
        self._bind()
        self.go0, self.go1 = go0, go1
        self._time_dependent = time_dependent
        self._sampled_time = None
        self._method = None
        self._stage = 0
        h = C.c_void_p()
        self._h = None
        self._chk(self.lib.pdb200_onestep_create(go0._h, go1._h, C.byref(h)))
        self._h = h

    _bound = False

    def _bind(self):
        if OneStepGridOperator._bound:
            return
        lib, vp = self.lib, C.c_void_p
        lib.pdb200_onestep_create.argtypes = [vp, vp, C.POINTER(vp)]
        lib.pdb200_onestep_destroy.argtypes = [vp]
        lib.pdb200_onestep_set_method.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int]
        lib.pdb200_onestep_set_dt_mode.argtypes = [vp, C.c_int]
        lib.pdb200_onestep_pre_step.argtypes = [vp, C.c_double, C.c_double]
        lib.pdb200_onestep_time_at_stage.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        lib.pdb200_onestep_pre_stage.argtypes = [vp, C.c_int, C.POINTER(vp)]
        lib.pdb200_onestep_pre_stage_begin.argtypes = [vp, C.c_int]
        lib.pdb200_onestep_pre_stage_add.argtypes = [vp, C.c_int, vp]
        lib.pdb200_onestep_const_residual.argtypes = [vp, vp]
        lib.pdb200_onestep_explicit_stage.argtypes = [vp, C.c_int, C.POINTER(vp), vp, C.c_double]
        lib.pdb200_onestep_explicit_stage_begin.argtypes = [vp, C.c_int]
        lib.pdb200_onestep_explicit_stage_add.argtypes = [vp, C.c_int, vp]
        lib.pdb200_onestep_explicit_stage_finish.argtypes = [vp, vp, C.c_double]
        for name in ("pdb200_onestep_residual", "pdb200_onestep_jacobian_apply", "pdb200_onestep_onthefly_apply"):
            getattr(lib, name).argtypes = [vp, vp, vp]
        lib.pdb200_onestep_jacobian.argtypes = [vp, vp, vp, C.c_int]
        lib.pdb200_onestep_stage_operator.argtypes = [vp, C.POINTER(vp)]
        lib.pdb200_onestep_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
        lib.pdb200_onestep_solve_stationary.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_double, C.c_double,
                                                        C.c_uint32, C.POINTER(SolveResult)]
        OneStepGridOperator._bound = True

    def _chk(self, rc):
        if rc != 0:
            raise PDELabError(self.lib.pdb200_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.pdb200_onestep_destroy(self._h)
                self._h = None
        except Exception:
            pass

    close = __del__

    # onestep.hh:78-91
    def divideMassTermByDeltaT(self):
        self._chk(self.lib.pdb200_onestep_set_dt_mode(self._h, self.DivideOperator1ByDT))

    def multiplySpatialTermByDeltaT(self):
        self._chk(self.lib.pdb200_onestep_set_dt_mode(self._h, self.MultiplyOperator0ByDT))

    def trialGridFunctionSpace(self):
        return self.go0.spec

    testGridFunctionSpace = trialGridFunctionSpace

    def globalSizeU(self):
        return self.go0.globalSizeU()

    globalSizeV = globalSizeU

    # onestep.hh:245-248
    def setMethod(self, method: TimeSteppingParameterInterface):
        s = method.s()
        a = np.zeros((s, s + 1))
        b = np.zeros((s, s + 1))
        for r in range(1, s + 1):
            for i in range(r + 1):
                a[r - 1, i], b[r - 1, i] = method.a(r, i), method.b(r, i)
        d = np.array([method.d(i) for i in range(s + 1)], dtype=np.float64)
        self._chk(self.lib.pdb200_onestep_set_method(self._h, s, a.ctypes.data, b.ctypes.data, d.ctypes.data,
                                                     1 if method.implicit() else 0))
        self._method = method

    # onestep.hh:250-254
    def preStep(self, method, time, dt):
        self.setMethod(method)
        self._chk(self.lib.pdb200_onestep_pre_step(self._h, float(time), float(dt)))

    def timeAtStage(self, stage):
        t = C.c_double()
        self._chk(self.lib.pdb200_onestep_time_at_stage(self._h, int(stage), C.byref(t)))
        return t.value

    def _set_time(self, t):
        """la0.setTime(t) (prestageengine.hh:210, residualengine.hh:155): re-sample time-dependent coefficients."""
        if self._time_dependent is None or self._sampled_time == t:
            return
        self.go0.update_coefficients(**self._time_dependent(t))
        self._sampled_time = t

    # onestep.hh:130-139
    def preStage(self, stage, x):
        assert len(x) >= stage
        self._chk(self.lib.pdb200_onestep_pre_stage_begin(self._h, int(stage)))
        for i in range(stage):
            self._set_time(self.timeAtStage(i))
            self._chk(self.lib.pdb200_onestep_pre_stage_add(self._h, i, _ptr(x[i])))
        self._stage = stage
        self._set_time(self.timeAtStage(stage))  # the stage operator itself lives at t + d_r dt

    def explicit_stage(self, stage, x, xr, reduction=1e-12):
        """One stage of an explicit method (explicit_jacobian_residual + the mass solve of ExplicitOneStepMethod::apply,
        onestep.hh:161-178, instationary/explicitonestep.hh:365-407): xr = -M^-1 sum_i (a_ri M x_i + b_ri dt R0(x_i))."""
        assert len(x) >= stage
        # split per earlier stage: every R0(x_i) is evaluated with the coefficients at t + d_i dt (the explicit engine
        # delegates to the pre-stage engine, prestageengine.hh:208-211)
        self._chk(self.lib.pdb200_onestep_explicit_stage_begin(self._h, int(stage)))
        for i in range(stage):
            self._set_time(self.timeAtStage(i))
            self._chk(self.lib.pdb200_onestep_explicit_stage_add(self._h, i, _ptr(x[i])))
        self._chk(self.lib.pdb200_onestep_explicit_stage_finish(self._h, _ptr(xr), float(reduction)))
        self._stage = stage
        return xr

    def const_residual(self, out):
        self._chk(self.lib.pdb200_onestep_const_residual(self._h, _ptr(out)))
        return out

    # onestep.hh:141-149
    def residual(self, x, r):
        self._chk(self.lib.pdb200_onestep_residual(self._h, _ptr(x), _ptr(r)))
        return r

    # onestep.hh:180-192
    def jacobian_apply(self, *args):
        if len(args) == 3:
            raise PDELabError("Your trying to use a non linear jacobian apply for a linear problem.")
        z, y = args
        self._chk(self.lib.pdb200_onestep_jacobian_apply(self._h, _ptr(z), _ptr(y)))
        return y

    def apply(self, x, y):
        self._chk(self.lib.pdb200_onestep_onthefly_apply(self._h, _ptr(x), _ptr(y)))
        return y

    # onestep.hh:151-159
    def jacobian(self, x, values, layout=abi.LAYOUT_CSR):
        self._chk(self.lib.pdb200_onestep_jacobian(self._h, _ptr(x), _ptr(values), layout))
        return values

    # onestep.hh:113-128: the pattern of go0 (the temporal operator's couplings are a subset)
    def fill_pattern(self, **kw):
        return self.go0.fill_pattern(**kw)

    # onestep.hh:194-211
    def interpolate(self, stage, xold, f, x):
        """x = interpolation of f(t_stage) on the constrained DOFs, xold elsewhere.  f: call-back t -> vector of nodal
        values in container order (what Dune::PDELab::interpolate(f, gfs, x) produces for Lagrange spaces)."""
        t = self.timeAtStage(stage)
        self._set_time(t)
        if x is not xold:
            x[:] = xold
        con = self.go0.constrained_dofs().astype(np.int64)
        if con.size:
            vals = np.asarray(f(t), dtype=np.float64)
            if hasattr(x, "data_ptr"):
                import torch
                idx = torch.from_numpy(con).to(x.device)
                x[idx] = torch.from_numpy(vals[con]).to(x.device)
            else:
                x[con] = vals[con]
        return x

    def stage_operator(self):
        h = C.c_void_p()
        self._chk(self.lib.pdb200_onestep_stage_operator(self._h, C.byref(h)))
        return _StageOperatorView(self.lib, h, self.go0.spec)

    def solve_stationary(self, x, reduction=1e-10, min_defect=1e-99, solver=abi.SOLVER_BICGSTAB,
                         precond=abi.PRECOND_NONE, matrix_free=True, maxiter=5000):
        """StationaryLinearProblemSolver::apply on the one-step operator (the stage solver of OneStepMethod),
        one device-resident call: x -= J^-1 residual(x)."""
        res = SolveResult()
        self._chk(self.lib.pdb200_onestep_solve_stationary(self._h, solver, precond, 1 if matrix_free else 0, _ptr(x),
                                                           float(reduction), float(min_defect), int(maxiter), C.byref(res)))
        return res.as_dict()

    def postStage(self):
        pass

    def postStep(self):
        pass

    def launch_count(self):
        n = C.c_uint64()
        self._chk(self.lib.pdb200_onestep_launch_count(self._h, C.byref(n)))
        return n.value


# ---- instationary/implicitonestep.hh -------------------------------------------------------------
class OneStepMethod:
    """OneStepMethod<T, IGOS, PDESOLVER, TrlV, TstV>::apply (implicitonestep.hh:122-262) for linear problems.

    The stage solver is StationaryLinearProblemSolver::apply (stationary/linearproblem.hh:188-302) run on the
    device against the fused stage operator: r = residual(x); solve J z = r; x -= z.
    """

    def __init__(self, method, igos: OneStepGridOperator, reduction=1e-10, solver=abi.SOLVER_BICGSTAB,
                 precond=abi.PRECOND_NONE, maxiter=5000, min_defect=1e-99, matrix_free=True):
        self.method, self.igos, self.matrix_free = method, igos, matrix_free
        self.reduction, self.solver, self.precond, self.maxiter, self.min_defect = reduction, solver, precond, maxiter, min_defect
        self.step = 1
        self.linear_solver_iterations = 0
        self.last_results = []

    def setMethod(self, method):
        self.method = method

    def _solve_stage(self, x):
        """pdesolver.apply(x): x -= J^-1 (R(x) + const_residual), one device-resident call."""
        res = self.igos.solve_stationary(x, reduction=self.reduction, min_defect=self.min_defect, solver=self.solver,
                                         precond=self.precond, matrix_free=self.matrix_free, maxiter=self.maxiter)
        if not res["converged"]:
            raise PDELabError("OneStepMethod: linear solver did not converge")
        return res

    def apply(self, time, dt, xold, xnew, f=None):
        """One step from xold (time) to xnew (time + dt); returns dt.  xnew holds the initial guess.
        With f (a call-back t -> vector of nodal values) the constrained DOFs of every stage are interpolated from f at
        the stage time before the solve — the second overload of the reference (implicitonestep.hh:264-400)."""
        m, igos = self.method, self.igos
        x = [xold]
        igos.preStep(m, time, dt)                                       # :159
        self.last_results = []
        for r in range(1, m.s() + 1):
            igos.preStage(r, x)                                         # :174
            xr = xnew if r == m.s() else (xnew.clone() if hasattr(xnew, "clone") else xnew.copy())
            init_guess = xnew if r == 1 else x[r - 1]
            if xr is not init_guess:
                xr[:] = init_guess
            if f is not None:
                igos.interpolate(r, init_guess, f, xr)                  # :351-352
            x.append(xr)
            res = self._solve_stage(xr)                                 # :191 pdesolver.apply(*x[r])
            self.linear_solver_iterations += res["iterations"]
            self.last_results.append(res)
            igos.postStage()
        igos.postStep()
        self.step += 1
        return dt


class ExplicitOneStepMethod:
    """ExplicitOneStepMethod<T, IGOS, LS, TrlV, TstV, TC>::apply (instationary/explicitonestep.hh:282-414) for QkDG
    spaces: every stage is one pass over the earlier stages plus the exact inverse of the block-diagonal mass matrix.
    The reference's time-step controller (CFL limit from the local operator) is not part of this path: dt is used as
    given (SimpleTimeController)."""

    def __init__(self, method, igos: OneStepGridOperator, reduction=0.99):
        if method.implicit():
            raise PDELabError("explicit one step method called with implicit scheme")   # explicitonestep.hh:226-228
        self.method, self.igos, self.reduction = method, igos, reduction
        self.step = 1

    def setMethod(self, method):
        if method.implicit():
            raise PDELabError("explicit one step method called with implicit scheme")
        self.method = method

    def setReduction(self, reduction):
        self.reduction = reduction

    def apply(self, time, dt, xold, xnew):
        m, igos = self.method, self.igos
        x = [xold]
        igos.preStep(m, time, dt)
        for r in range(1, m.s() + 1):
            xr = xnew if r == m.s() else (xnew.clone() if hasattr(xnew, "clone") else xnew.copy())
            igos.explicit_stage(r, x, xr, 1e-12 if self.reduction >= 0.99 else self.reduction)
            x.append(xr)
            igos.postStage()
        igos.postStep()
        self.step += 1
        return dt


# ---- localoperator/l2.hh ---------------------------------------------------------------------------
def l2_spec(spec, scaling=1.0):
    """Problem description of the L2 mass operator (localoperator/l2.hh:149-250: alpha_volume = scaling * (u, v),
    no skeleton or boundary terms) on the grid and function space of `spec`: the convection-diffusion form with
    A = 0, b = 0, c = scaling and boundary type None — the temporal operator go1 of OneStepGridOperator.
    The reference's `intorderadd` of L2 is not carried over: k+1 Gauss points integrate the Qk mass matrix exactly."""
    nc, nbf = spec.ncells, spec.num_boundary_faces
    # alpha = 0: with weightsOff the SIPG penalty does not contain A (convectiondiffusiondg.hh:334-338) and would
    # survive A = 0; the mass operator has no face terms at all
    return spec.replace(a_mode=abi.A_SCALAR, A=np.zeros(nc), b=None, c=np.full(nc, float(scaling)), f=None,
                        bctype=np.full(nbf, abi.BC_NONE, dtype=np.int8), g=None, j=None, o=None, alpha=0.0)

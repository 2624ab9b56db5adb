"""CPU restatement of PDELab's one-step engines on top of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Follows gridoperator/onestep/localassembler.hh:101-130 (dt factors), prestageengine.hh:166-230 (constant part,
|a|,|b| > 1e-6 switches, per-stage times), residualengine.hh:135-176, jacobianengine.hh:92-96 and
jacobianapplyengine.hh (stage weights b_rr * dt_factor0 and dt_factor1; constraints applied after the sum).
The two operators are evaluated SEPARATELY by the oracle and combined with the engine weights — the order of
operations of the reference — so the comparison checks the device's fused stage operator (csrc/onestep.cu)
against an independent formulation.
"""
import numpy as np
import scipy.sparse as sp

from oracle import Oracle

DivideOperator1ByDT, MultiplyOperator0ByDT, DoNotAssembleDT = 0, 1, 2


class OneStepOracle:
    def __init__(self, spec0, spec1, spec0_at=None):
        self.spec0_at = spec0_at or (lambda t: spec0)
        self.o1 = Oracle(spec1)
        self.con = Oracle(spec0).constrained_dofs().astype(np.int64)
        self.n = spec0.num_dofs
        self.dt_mode = MultiplyOperator0ByDT
        self.const = np.zeros(self.n)

    def preStep(self, method, time, dt):
        self.m, self.time, self.dt = method, time, dt
        if not method.implicit():
            self.dt_mode = DoNotAssembleDT
        self.f0, self.f1 = {DivideOperator1ByDT: (1.0, 1.0 / dt), MultiplyOperator0ByDT: (dt, 1.0),
                            DoNotAssembleDT: (1.0, 1.0)}[self.dt_mode]

    def _o0(self, i):
        return Oracle(self.spec0_at(self.time + self.m.d(i) * self.dt))

    def preStage(self, stage, xs):
        self.stage = stage
        c = np.zeros(self.n)
        for i in range(stage):
            a, b = self.m.a(stage, i), self.m.b(stage, i)
            if abs(b) > 1e-6:
                c += b * self.f0 * self._o0(i).residual(xs[i])
            if abs(a) > 1e-6:
                c += a * self.f1 * self.o1.residual(xs[i])
        c[self.con] = 0.0
        self.const = c

    def _weights(self):
        b_rr = self.m.b(self.stage, self.stage)
        return (b_rr * self.f0 if abs(b_rr) > 1e-6 else 0.0), self.f1

    def residual(self, x, r=None):
        r = np.zeros(self.n) if r is None else r
        w0, w1 = self._weights()
        if w0 != 0.0:
            r += w0 * self._o0(self.stage).residual(x)
        r += w1 * self.o1.residual(x)
        r += self.const
        r[self.con] = 0.0
        return r

    def jacobian_apply(self, z, y=None):
        y = np.zeros(self.n) if y is None else y
        w0, w1 = self._weights()
        if w0 != 0.0:
            y += w0 * self._o0(self.stage).jacobian_apply(z)
        y += w1 * self.o1.jacobian_apply(z)
        y[self.con] = 0.0
        return y

    def matrix(self):
        """Stage Jacobian as scipy CSR; constrained rows are unit rows (set_trivial_rows after the weighted sum)."""
        w0, w1 = self._weights()

        def csr(o):
            rp, ci, v = o.jacobian()
            return sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=(self.n, self.n))
        M = (w0 * csr(self._o0(self.stage)) + w1 * csr(self.o1)).tolil()
        for i in self.con:
            M.rows[i], M.data[i] = [int(i)], [1.0]
        return M.tocsr()


def explicit_stage(spec0, spec1, method, stage, time, dt, xs, spec0_at=None):
    """x_r = -M^-1 sum_{i<r} (a_ri M x_i + b_ri dt R0(x_i; t + d_i dt)): ExplicitOneStepMethod::apply, one stage
    (instationary/explicitonestep.hh:365-407; D = -M by the weight -1 of onestep/jacobianresidualengine.hh:303; the
    engine delegates to the pre-stage engine, which evaluates stage i at its own time, prestageengine.hh:208-211).
    spec0_at(t): the spatial problem with its coefficients sampled at time t."""
    import scipy.sparse.linalg as spla
    o0, o1 = Oracle(spec0), Oracle(spec1)
    n = spec0.num_dofs
    alpha, beta = np.zeros(n), np.zeros(n)
    for i in range(stage):
        a, b = method.a(stage, i), method.b(stage, i)
        if spec0_at is not None:
            o0 = Oracle(spec0_at(time + method.d(i) * dt))
        if abs(b) > 1e-6:
            beta += b * o0.residual(xs[i])
        if abs(a) > 1e-6:
            alpha += a * o1.residual(xs[i])
    alpha += dt * beta
    rp, ci, v = o1.jacobian()
    M = sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=(n, n))
    return spla.spsolve((-M).tocsc(), alpha)

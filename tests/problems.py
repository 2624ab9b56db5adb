"""Seeded synthetic problems shared by the parity tests, smoke() and bench.py.

Vectors: default-seeded std::mt19937_64 + uniform_real_distribution<double>(0,1) in container
order (test/test-blocked-istl-ordering.cc:45-48).  Coefficients: kappa_e = 10^(2u-1) with u from
mt19937_64(42) in cell order.
"""
import numpy as np

from pdelab_b200 import abi
from pdelab_b200.abi import ProblemSpec


class MT19937_64:
    """std::mt19937_64 (so that vectors equal the reference test's, seed 5489 by default)."""

    def __init__(self, seed=5489):
        self.bitgen = None
        n, f = 312, 6364136223846793005
        mt = np.zeros(n, dtype=np.uint64)
        mt[0] = np.uint64(seed)
        x = int(seed)
        for i in range(1, n):
            x = (f * (x ^ (x >> 62)) + i) & 0xFFFFFFFFFFFFFFFF
            mt[i] = x
        self.mt = [int(v) for v in mt]
        self.idx = n

    def _twist(self):
        n, m = 312, 156
        mt = self.mt
        UM, LM, A = 0xFFFFFFFF80000000, 0x7FFFFFFF, 0xB5026F5AA96619E9
        for i in range(n):
            x = (mt[i] & UM) | (mt[(i + 1) % n] & LM)
            xa = x >> 1
            if x & 1:
                xa ^= A
            mt[i] = mt[(i + m) % n] ^ xa
        self.idx = 0

    def next_u64(self):
        if self.idx >= 312:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        return y

    def uniform01(self, count):
        """libstdc++ generate_canonical<double,53> with a 64-bit engine: one draw per value."""
        out = np.empty(count)
        for i in range(count):
            v = self.next_u64() * (1.0 / 18446744073709551616.0)
            out[i] = v if v < 1.0 else np.nextafter(1.0, 0.0)
        return out


def mt_vector(n, seed=5489):
    """Reference-style random vector.  Exact mt19937_64 stream for small n; for large n a numpy
    generator with the same distribution (the exact stream is a pure-Python loop)."""
    if n <= 200_000:
        return MT19937_64(seed).uniform01(n)
    return np.random.Generator(np.random.MT19937(seed)).random(n)


def kappa_field(ncells, seed=42):
    u = mt_vector(ncells, seed)
    return 10.0 ** (2.0 * u - 1.0)


def dg_problem(cells, degree=2, extent=None, a="scalar", with_b=False, with_c=False, with_f=False,
               bc="dirichlet", method=abi.DG_SIPG, weights=abi.DG_WEIGHTS_ON, alpha=3.0, seed=42,
               kernel=abi.KERNEL_AUTO, intorderadd=0, basis=abi.BASIS_LAGRANGE):
    """QkDG ConvectionDiffusionDG problem (SIPG, weightsOn, alpha=3: test/matrixfree/
    matrix_free_linear.cc:105-108) with synthetic coefficient fields."""
    cells = tuple(cells)
    dim = len(cells)
    nc = int(np.prod(cells))
    rng = np.random.default_rng(seed)
    kw = {}
    if a == "identity":
        kw.update(a_mode=abi.A_IDENTITY)
    elif a == "scalar":
        kw.update(a_mode=abi.A_SCALAR, A=kappa_field(nc, seed))
    elif a == "diagonal":
        kw.update(a_mode=abi.A_DIAGONAL, A=10.0 ** (2.0 * rng.random((nc, dim)) - 1.0))
    elif a == "full":
        # symmetric positive definite tensors
        Q = rng.standard_normal((nc, dim, dim))
        A = np.einsum("nij,nkj->nik", Q, Q) + 0.5 * np.eye(dim)
        kw.update(a_mode=abi.A_FULL, A=A)
    else:
        raise ValueError(a)
    spec = ProblemSpec(cells, space=abi.SPACE_QKDG, degree=degree, upper=extent, method=method,
                       weights=weights, alpha=alpha, intorderadd=intorderadd, kernel=kernel, basis=basis, **kw)
    extra = {}
    if with_b:
        extra["b"] = rng.standard_normal((nc, dim))
    if with_c:
        extra["c"] = rng.random(nc)
    if with_f:
        extra["f"] = rng.standard_normal((nc, spec.nq))
    nbf = spec.num_boundary_faces
    if bc == "dirichlet":
        pass
    elif bc == "dirichlet_g":
        extra["g"] = rng.standard_normal((nbf, spec.nfq))
    elif bc == "mixed":
        # no outflow faces here: with random b they would raise "Outflow on inflow"
        extra["bctype"] = rng.choice(np.array([abi.BC_DIRICHLET, abi.BC_NEUMANN, abi.BC_NONE], dtype=np.int8), nbf)
        extra["g"] = rng.standard_normal((nbf, spec.nfq))
        extra["j"] = rng.standard_normal((nbf, spec.nfq))
    else:
        raise ValueError(bc)
    return spec.replace(**extra)


def fem_problem(cells, degree=1, extent=None, a="scalar", with_b=False, with_c=False, with_f=True,
                bc="dirichlet", seed=42, intorderadd=0):
    """Conforming Qk ConvectionDiffusionFEM problem."""
    cells = tuple(cells)
    dim = len(cells)
    nc = int(np.prod(cells))
    rng = np.random.default_rng(seed)
    kw = {}
    if a == "identity":
        kw.update(a_mode=abi.A_IDENTITY)
    elif a == "scalar":
        kw.update(a_mode=abi.A_SCALAR, A=kappa_field(nc, seed))
    elif a == "diagonal":
        kw.update(a_mode=abi.A_DIAGONAL, A=10.0 ** (2.0 * rng.random((nc, dim)) - 1.0))
    elif a == "full":
        Q = rng.standard_normal((nc, dim, dim))
        kw.update(a_mode=abi.A_FULL, A=np.einsum("nij,nkj->nik", Q, Q) + 0.5 * np.eye(dim))
    spec = ProblemSpec(cells, space=abi.SPACE_QK, degree=degree, upper=extent, intorderadd=intorderadd, **kw)
    extra = {}
    if with_b:
        extra["b"] = rng.standard_normal((nc, dim))
    if with_c:
        extra["c"] = rng.random(nc)
    if with_f:
        extra["f"] = rng.standard_normal((nc, spec.nq))
    nbf = spec.num_boundary_faces
    if bc == "mixed":
        bct = rng.choice(np.array([abi.BC_DIRICHLET, abi.BC_NEUMANN, abi.BC_OUTFLOW], dtype=np.int8), nbf)
        extra["bctype"] = bct
        extra["j"] = rng.standard_normal((nbf, spec.nfq))
        extra["o"] = rng.standard_normal((nbf, spec.nfq))
    return spec.replace(**extra)


def pointwise_problem(spec, which=("A", "b", "c", "bctype"), seed=7):
    """Turns `spec` into a problem with spatially varying coefficient call-backs, sampled into the point-wise layouts
    of include/pdelab_b200.h (layout (2)): A(x) non-constant per cell (permeabilityIsConstantPerCell() == false), a
    rotating velocity b(x) = (-y, x, ...), c(x) varying inside the cells, and for QkDG a boundary type that changes
    INSIDE boundary faces.  The fields also depend on the cell number (discontinuous across cells), so that a kernel
    reading the wrong cell's face trace is caught.  The call-backs stay on the spec (spec.fns) for
    tests/numpy_assembly.py, which evaluates them at its own quadrature points."""
    dim, N = spec.dim, spec.cells
    h = [(spec.upper[d] - spec.lower[d]) / N[d] for d in range(dim)]
    x1, _ = np.polynomial.legendre.leggauss(spec.m)
    xq = 0.5 * (x1 + 1.0)
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((dim, dim)) * 0.2

    def A_fn(x, e):
        s = (1.0 + 0.5 * np.sin(3.0 * x[0] + 2.0 * x[1])) * (1.0 + 0.1 * (e % 3))
        if spec.a_mode == abi.A_SCALAR:
            return s * np.eye(dim)
        if spec.a_mode == abi.A_DIAGONAL:
            return np.diag([s * (1.0 + 0.3 * d + 0.2 * x[d]) for d in range(dim)])
        M = np.eye(dim) * s + (R @ R.T) * (1.0 + x[0])
        return M

    def b_fn(x, e):
        v = np.zeros(dim)
        v[0], v[1] = -(x[1] - 0.5), x[0] - 0.5
        if dim == 3:
            v[2] = 0.3 + 0.2 * x[2]
        return v * (1.0 + 0.05 * (e % 2))

    def c_fn(x, e):
        return 1.0 + x[0] * x[1] + 0.1 * (e % 4)

    def bc_fn(x):
        # changes type in the middle of boundary faces (never Outflow: the rotating b has inflow parts)
        s = np.sin(7.0 * x[0] + 5.0 * x[1] + (3.0 * x[2] if dim == 3 else 0.0))
        return abi.BC_DIRICHLET if s > 0.2 else (abi.BC_NEUMANN if s > -0.4 else abi.BC_NONE)

    nc, nq, nfq, NP = spec.ncells, spec.nq, spec.nfq, spec.points_per_cell
    # physical coordinates of the NP sample points of every cell
    X = np.zeros((nc, NP, dim))
    tang = [[t for t in range(dim) if t != d] for d in range(dim)]
    for e in range(nc):
        c, r = [], e
        for d in range(dim):
            c.append(r % N[d])
            r //= N[d]
        x0 = np.array([spec.lower[d] + c[d] * h[d] for d in range(dim)])
        for q in range(nq):
            qq = q
            for d in range(dim):
                X[e, q, d] = x0[d] + xq[qq % spec.m] * h[d]
                qq //= spec.m
        for d in range(dim):
            for side in range(2):
                for q in range(nfq):
                    pt, qq = spec.face_point(d, side, q), q
                    X[e, pt, d] = x0[d] + side * h[d]
                    for t in tang[d]:
                        X[e, pt, t] = x0[t] + xq[qq % spec.m] * h[t]
                        qq //= spec.m
    kw, fns, mask = {}, {}, 0
    if "A" in which and spec.a_mode != abi.A_IDENTITY:
        full = np.array([[A_fn(X[e, p], e) for p in range(NP)] for e in range(nc)])
        if spec.a_mode == abi.A_SCALAR:
            kw["A"] = np.ascontiguousarray(full[:, :, 0, 0])
        elif spec.a_mode == abi.A_DIAGONAL:
            kw["A"] = np.ascontiguousarray(np.einsum("epii->epi", full))
        else:
            kw["A"] = np.ascontiguousarray(full)
        fns["A"] = A_fn
        mask |= abi.POINTWISE_A
    if "b" in which:
        kw["b"] = np.array([[b_fn(X[e, p], e) for p in range(NP)] for e in range(nc)])
        fns["b"] = b_fn
        mask |= abi.POINTWISE_B
    if "c" in which:
        kw["c"] = np.array([[c_fn(X[e, q], e) for q in range(nq)] for e in range(nc)])
        fns["c"] = c_fn
        mask |= abi.POINTWISE_C
    if "bctype" in which and spec.space == abi.SPACE_QKDG:
        bct = np.zeros((spec.num_boundary_faces, nfq), dtype=np.int8)
        for d in range(dim):
            for side in range(2):
                off = spec.boundary_face_offset(d, side)
                for e in range(nc):
                    c, r = [], e
                    for dd in range(dim):
                        c.append(r % N[dd])
                        r //= N[dd]
                    if c[d] != (N[d] - 1 if side else 0):
                        continue
                    idx, stride = 0, 1
                    for dd in range(dim):
                        if dd != d:
                            idx += stride * c[dd]
                            stride *= N[dd]
                    for q in range(nfq):
                        bct[off + idx, q] = bc_fn(X[e, spec.face_point(d, side, q)])
        kw["bctype"] = bct
        rng2 = np.random.default_rng(seed + 1)
        for nm in ("g", "j"):
            if spec.arrays.get(nm) is None:
                kw[nm] = rng2.standard_normal((spec.num_boundary_faces, nfq))
        fns["bctype"] = bc_fn
        mask |= abi.POINTWISE_BCTYPE
    out = spec.replace(pointwise=mask, **kw)
    out.fns = fns
    return out


def rel_err(a, b):
    """Norm-relative error  ||a-b||_inf / ||b||_inf  (entry-wise relative error is
    meaningless where cancellation gives ~0)."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / (scale if scale > 0 else 1.0))

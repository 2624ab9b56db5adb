"""Manufactured-solution helpers for the restated reference integration tests (test infrastructure).

Builds the coefficient arrays of the C ABI from analytic u and f = -Laplace(u) the way a PDELab
parameter class is sampled (f at volume quadrature points, g at face quadrature points), solves
the linear problem like StationaryLinearProblemSolver::apply (stationary/linearproblem.hh:188-302):
    r = residual(x0);  J z = r;  x = x0 - z
and integrates the squared L2 error like integrateGridFunction(DifferenceSquaredAdapter, ., 10).
Works with any operator object exposing residual / jacobian_apply / pattern+jacobian (the CPU
oracle or the CUDA GridOperator)."""
import itertools

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from numpy_assembly import Grid, dof_map, gauss01
from pdelab_b200 import abi


def sample_data(spec, u_exact, f_func):
    """f [cells, m^dim] and g [bfaces, m^(dim-1)] at the Gauss points (ascending, x fastest)."""
    G = Grid(spec)
    dim, m = G.dim, G.m
    lo = spec.lower
    f = np.zeros((G.ncells, m ** dim))
    g = np.zeros((spec.num_boundary_faces, m ** (dim - 1)))
    qidx = np.array(list(itertools.product(*[range(m)] * dim)))[:, ::-1]
    fidx = np.array(list(itertools.product(*[range(m)] * (dim - 1))))[:, ::-1]
    for c in G.cells():
        e = G.cell_index(c)
        X = np.stack([lo[d] + G.h[d] * (c[d] + G.xq[qidx[:, d]]) for d in range(dim)], axis=1)
        f[e] = f_func(X)
        for d in range(dim):
            for side in range(2):
                if c[d] != (G.N[d] - 1 if side else 0):
                    continue
                tang = [x for x in range(dim) if x != d]
                X = np.zeros((m ** (dim - 1), dim))
                X[:, d] = lo[d] + G.h[d] * (c[d] + side)
                for t, dd in enumerate(tang):
                    X[:, dd] = lo[dd] + G.h[dd] * (c[dd] + G.xq[fidx[:, t]])
                g[G.bface_index(c, d, side)] = u_exact(X)
    return spec.replace(f=f, g=g)


def node_coordinates(spec):
    """Physical coordinates of every DOF's Lagrange node, [ndofs, dim]."""
    G = Grid(spec)
    dmap, nd = dof_map(G)
    X = np.zeros((nd, G.dim))
    loc = np.array(list(itertools.product(*[range(G.n1)] * G.dim)))[:, ::-1]
    for c in G.cells():
        e = G.cell_index(c)
        for d in range(G.dim):
            X[dmap[e], d] = spec.lower[d] + G.h[d] * (c[d] + loc[:, d] / G.k)
    return X


def l2_error_squared(spec, x, u_exact, npts=6):
    G = Grid(spec)
    dmap, _ = dof_map(G)
    xq, wq = gauss01(npts)
    phi, _ = G.basis([xq] * G.dim)
    qidx = np.array(list(itertools.product(*[range(npts)] * G.dim)))[:, ::-1]
    w = np.prod(wq[qidx], axis=1) * float(np.prod(G.h))
    err = 0.0
    for c in G.cells():
        e = G.cell_index(c)
        X = np.stack([spec.lower[d] + G.h[d] * (c[d] + xq[qidx[:, d]]) for d in range(G.dim)], axis=1)
        err += float(np.sum(w * (phi @ x[dmap[e]] - u_exact(X)) ** 2))
    return err


class OracleOps:
    """Adapter: CPU oracle with the call surface solve_stationary needs."""

    def __init__(self, spec):
        from oracle import Oracle
        self.o = Oracle(spec)
        self.n = self.o.num_dofs

    def residual(self, x):
        return self.o.residual(x)

    def jacobian_apply(self, z):
        return self.o.jacobian_apply(z)

    def matrix(self):
        rowptr, colidx, values = self.o.jacobian()
        return sp.csr_matrix((values, colidx.astype(np.int64), rowptr.astype(np.int64)), shape=(self.n, self.n))


class GpuOps:
    """Adapter: CUDA GridOperator through the C ABI (host buffers)."""

    def __init__(self, spec):
        from pdelab_b200.capi import GridOperator
        self.go = GridOperator(spec)
        self.n = self.go.globalSizeU()

    def residual(self, x):
        return self.go.residual(np.ascontiguousarray(x), np.zeros(self.n))

    def jacobian_apply(self, z):
        return self.go.jacobian_apply(np.ascontiguousarray(z), np.zeros(self.n))

    def matrix(self):
        rowptr, colidx = self.go.fill_pattern()
        values = self.go.jacobian(np.zeros(self.n), np.zeros(colidx.size))
        return sp.csr_matrix((values, colidx.astype(np.int64), rowptr.astype(np.int64)), shape=(self.n, self.n))


def bicgstab(apply, b, reduction, maxit=5000):
    """Unpreconditioned BiCGSTAB (the reference tests use dune-istl's BiCGSTABSolver); returns
    (x, iterations).  Deterministic, so iteration counts can be compared between operators."""
    x = np.zeros_like(b)
    r = b.copy()
    rt = r.copy()
    rho = alpha = omega = 1.0
    v = np.zeros_like(b)
    p = np.zeros_like(b)
    norm0 = np.linalg.norm(r)
    if norm0 == 0.0:
        return x, 0
    for it in range(1, maxit + 1):
        rho_new = rt @ r
        beta = (rho_new / rho) * (alpha / omega)
        rho = rho_new
        p = r + beta * (p - omega * v)
        v = apply(p)
        alpha = rho / (rt @ v)
        s = r - alpha * v
        if np.linalg.norm(s) < reduction * norm0:
            return x + alpha * p, it
        t = apply(s)
        omega = (t @ s) / (t @ t)
        x = x + alpha * p + omega * s
        r = s - omega * t
        if np.linalg.norm(r) < reduction * norm0:
            return x, it
    raise RuntimeError("BiCGSTAB did not converge")


def solve_stationary(ops, x0, matrix_free=False, reduction=1e-10):
    """StationaryLinearProblemSolver::apply.  Returns (x, krylov iterations or None)."""
    r = ops.residual(x0)
    if matrix_free:
        z, its = bicgstab(ops.jacobian_apply, r, reduction)
        return x0 - z, its
    J = ops.matrix()
    return x0 - spla.spsolve(J.tocsc(), r), None

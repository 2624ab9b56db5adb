"""Independent numpy derivation of the DG operator on an axis-aligned grid.

Not the oracle: it does not follow the reference's loops.  It assembles y = J z from the weak form
in Kronecker (tensor-product) form with exactly integrated 1-D matrices, for cell-wise constant
diagonal diffusion, reaction c and b = 0.  Used (a) to cross-check the oracle restatement from a
second derivation and (b) as the mathematical specification of the fast CUDA kernel.
"""
import numpy as np
from numpy.polynomial import polynomial as P


def lagrange_1d(k):
    """Coefficient arrays of the Lagrange polynomials on nodes j/k (ascending powers)."""
    nodes = np.arange(k + 1) / k
    polys = []
    for i in range(k + 1):
        c = np.array([1.0])
        for j in range(k + 1):
            if j != i:
                c = P.polymul(c, np.array([-nodes[j], 1.0]) / (nodes[i] - nodes[j]))
        polys.append(c)
    return polys


def matrices_1d(k):
    """Exact mass, stiffness, advection matrices and end-point derivative vectors on [0,1]."""
    ps = lagrange_1d(k)
    dps = [P.polyder(p) for p in ps]
    n = k + 1

    def integ(c):
        ci = P.polyint(c)
        return P.polyval(1.0, ci) - P.polyval(0.0, ci)

    M = np.array([[integ(P.polymul(ps[i], ps[j])) for j in range(n)] for i in range(n)])
    K = np.array([[integ(P.polymul(dps[i], dps[j])) for j in range(n)] for i in range(n)])
    d0 = np.array([P.polyval(0.0, dp) for dp in dps])
    d1 = np.array([P.polyval(1.0, dp) for dp in dps])
    return M, K, d0, d1


def dg_apply_kron(cells, extent, k, Adiag, z, alpha, theta=-1.0, weights_on=True, c=None,
                  dirichlet=None):
    """y = J z for SIPG/NIPG/IIPG QkDG, diagonal cell-wise A, b = 0.

    cells: (Nx,Ny,Nz) or (Nx,Ny); Adiag: [ncells, dim]; z: flat DG vector (cell-major, x fastest);
    dirichlet: optional bool array [dim][2] -> whether that outer side is Dirichlet (else no
    u-dependent boundary term: None/Neumann, or Outflow with b = 0)."""
    dim = len(cells)
    N = list(cells)
    h = [extent[d] / N[d] for d in range(dim)]
    n1 = k + 1
    M, K, d0, d1 = matrices_1d(k)
    Minv = np.linalg.inv(M)
    # z as array [Nz,Ny,Nx, kz,ky,kx]
    shp = tuple(reversed(N)) + (n1,) * dim
    Z = np.asarray(z).reshape(shp)
    Acell = np.asarray(Adiag).reshape(tuple(reversed(N)) + (dim,))
    T = np.zeros_like(Z)
    pen = k * (k + dim - 1)
    for d in range(dim):
        cax = dim - 1 - d          # cell axis of direction d
        nax = 2 * dim - 1 - d      # node axis of direction d
        Zd = np.moveaxis(np.moveaxis(Z, cax, 0), nax, -1)  # [N_d, ..., n1 along d]  (views)
        a = np.moveaxis(Acell[..., d], cax, 0)             # [N_d, other cells]
        nd = N[d]
        out = np.zeros(Zd.shape)
        # broadcasting helper: a has cell dims only; node dims follow
        def ex(v):
            return v.reshape(v.shape + (1,) * (dim - 1))
        out += (ex(a)[..., None] / h[d]) * np.einsum("ij,...j->...i", K, Zd)
        for side in (0, 1):
            # per cell: neighbour coefficient (or boundary)
            if side == 0:
                a_o = np.concatenate([a[:1], a[:-1]], axis=0)
                Zo = np.concatenate([Zd[:1], Zd[:-1]], axis=0)
                tr_s, tr_o, ds, do, nsign, row = 0, k, d0, d1, -1.0, 0
            else:
                a_o = np.concatenate([a[1:], a[-1:]], axis=0)
                Zo = np.concatenate([Zd[1:], Zd[-1:]], axis=0)
                tr_s, tr_o, ds, do, nsign, row = k, 0, d1, d0, 1.0, k
            interior = np.ones(nd, dtype=bool)
            interior[0 if side == 0 else nd - 1] = False
            if weights_on:
                ws = a_o / (a + a_o + 1e-20)
                wo = a / (a + a_o + 1e-20)
                harm = 2 * a * a_o / (a + a_o + 1e-20)
            else:
                ws = np.full_like(a, 0.5)
                wo = np.full_like(a, 0.5)
                harm = np.ones_like(a)
            isd = True if dirichlet is None else bool(dirichlet[d][side])
            bmask = (~interior).reshape((nd,) + (1,) * (a.ndim - 1))
            # boundary cells: ws = 1, wo = 0, harm = a (weights on) or 1; or no term at all
            ws = np.where(bmask, 1.0 if isd else 0.0, ws)
            wo = np.where(bmask, 0.0, wo)
            harm = np.where(bmask, (a if weights_on else 1.0) if isd else 0.0, harm)
            jmask = np.where(bmask, 1.0 if isd else 0.0, 1.0)  # jump active?
            gamma = alpha / h[d] * harm * pen
            us = Zd[..., tr_s]
            uo = np.where(ex(bmask), 0.0, Zo[..., tr_o])
            jump = (us - uo) * ex(jmask)
            dus = np.einsum("j,...j->...", ds, Zd) / h[d]
            duo = np.einsum("j,...j->...", do, Zo) / h[d]
            flux = ex(ws * a) * dus + ex(wo * a_o) * duo
            out[..., row] += -nsign * flux + ex(gamma) * jump
            out += (theta * jump * ex(ws * a) * nsign / h[d])[..., None] * ds
        Td = np.einsum("ij,...j->...i", Minv, out) / h[d]
        T += np.moveaxis(np.moveaxis(Td, -1, nax), 0, cax)
    if c is not None:
        T += np.asarray(c).reshape(tuple(reversed(N)) + (1,) * dim) * Z
    vol = np.prod(h)
    Y = T * vol
    for d in range(dim):
        nax = 2 * dim - 1 - d
        Y = np.moveaxis(np.einsum("ij,...j->...i", M, np.moveaxis(Y, nax, -1)), -1, nax)
    # memory layout: cell-major, local index x fastest  == reshape of [cells(z,y,x), k(z,y,x)]
    return Y.reshape(-1)

"""Independent dense numpy assembly of the two weak forms (test infrastructure).

NOT the oracle and NOT a restatement of the reference's loops: the operators are assembled from
their textbook bilinear forms in jump/average notation,

  DG   a(u,v) = sum_K int_K (A grad u - b u).grad v + c u v
              + sum_F int_F  beta u^up [v] - {A grad u . n}_w [v] + theta [u] {A grad v . n}_w
                             + gamma [u][v]
  FEM  a(u,v) = sum_K int_K (A grad u - b u).grad v + c u v + sum_{F outflow} int_F (b.n) u v

with [v] = v_s - v_n, n pointing from s to n, {q}_w = w_s q_s + w_n q_n, using
  * Lagrange polynomials in monomial-coefficient form (numpy.polynomial), not the product formula,
  * numpy.polynomial.legendre.leggauss for the Gauss rule,
  * an enumerative (dictionary-based) DOF numbering for conforming Q1/Q2 that lists the grid
    entities one by one instead of using the closed-form index formula.
It yields a dense matrix J and a vector r0 with  residual(x) = J x + r0  (before constraints),
against which tests/test_oracle_vs_numpy.py checks the C++ oracle.

Spatially varying coefficients: a spec made by problems.pointwise_problem carries the analytic fields as Python
call-backs (spec.fns: A(x, cell), b(x, cell), c(x, cell), bctype(x)); this assembly EVALUATES THEM at its own physical quadrature
points and never reads the point-wise arrays, so it also checks the sample-point layout of pdelab_b200.h.
"""
import itertools

import numpy as np
from numpy.polynomial import polynomial as Pl

from pdelab_b200 import abi


def gauss01(m):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_tables(k, pts):
    """values[p, i], derivs[p, i] of the Lagrange polynomials on nodes j/k at the points pts."""
    nodes = np.arange(k + 1) / k
    V = np.zeros((len(pts), k + 1))
    D = np.zeros((len(pts), k + 1))
    for i in range(k + 1):
        c = np.array([1.0])
        for j in range(k + 1):
            if j != i:
                c = Pl.polymul(c, np.array([-nodes[j], 1.0]) / (nodes[i] - nodes[j]))
        V[:, i] = Pl.polyval(np.asarray(pts), c)
        D[:, i] = Pl.polyval(np.asarray(pts), Pl.polyder(c))
    return V, D


def basis_tables(k, pts, basis=0):
    """1-D basis of the QkDG space as power-basis polynomials (independent of the oracle's recurrences / closed forms):
    0 Lagrange on j/k, 1 shifted Legendre P_n(2x-1) via numpy.polynomial.legendre, 2 Lagrange on the Gauss-Lobatto points
    (end points + roots of P_k' found numerically), ascending."""
    if basis == 0:
        return lagrange_tables(k, pts)
    from numpy.polynomial import legendre as Lg
    pts = np.asarray(pts)
    V = np.zeros((len(pts), k + 1))
    D = np.zeros((len(pts), k + 1))
    if basis == 1:
        for n in range(k + 1):
            c = np.zeros(n + 1)
            c[n] = 1.0
            V[:, n] = Lg.legval(2.0 * pts - 1.0, c)
            D[:, n] = 2.0 * Lg.legval(2.0 * pts - 1.0, Lg.legder(c)) if n > 0 else 0.0
        return V, D
    c = np.zeros(k + 1)
    c[k] = 1.0
    inner = np.sort(Lg.legroots(Lg.legder(c))) if k > 1 else np.array([])
    nodes = (1.0 + np.concatenate([[-1.0], inner, [1.0]])) / 2.0
    for i in range(k + 1):
        c = np.array([1.0])
        for j in range(k + 1):
            if j != i:
                c = Pl.polymul(c, np.array([-nodes[j], 1.0]) / (nodes[i] - nodes[j]))
        V[:, i] = Pl.polyval(pts, c)
        D[:, i] = Pl.polyval(pts, Pl.polyder(c))
    return V, D


class Grid:
    def __init__(self, spec):
        self.spec = spec
        self.dim = spec.dim
        self.N = list(spec.cells)
        self.h = [(spec.upper[d] - spec.lower[d]) / self.N[d] for d in range(self.dim)]
        self.k = spec.degree
        self.n1 = self.k + 1
        self.n = self.n1 ** self.dim
        self.m = spec.m
        self.ncells = int(np.prod(self.N))
        self.xq, self.wq = gauss01(self.m)

    def cells(self):
        # lexicographic, x fastest
        for c in itertools.product(*[range(n) for n in reversed(self.N)]):
            yield tuple(reversed(c))

    def cell_index(self, c):
        idx, stride = 0, 1
        for d in range(self.dim):
            idx += stride * c[d]
            stride *= self.N[d]
        return idx

    def bface_index(self, c, dirn, side):
        off = 0
        for d in range(self.dim):
            for s in range(2):
                if d == dirn and s == side:
                    idx, stride = 0, 1
                    for dd in range(self.dim):
                        if dd != dirn:
                            idx += stride * c[dd]
                            stride *= self.N[dd]
                    return off + idx
                off += self.ncells // self.N[d]
        raise ValueError

    def A(self, cell):
        s, dim = self.spec, self.dim
        A = s.arrays["A"]
        if s.a_mode == abi.A_IDENTITY:
            return np.eye(dim)
        if s.a_mode == abi.A_SCALAR:
            return np.eye(dim) * np.asarray(A).reshape(-1)[cell]
        if s.a_mode == abi.A_DIAGONAL:
            return np.diag(np.asarray(A).reshape(-1, dim)[cell])
        return np.asarray(A).reshape(-1, dim, dim)[cell]

    def b(self, cell):
        b = self.spec.arrays["b"]
        return np.zeros(self.dim) if b is None else np.asarray(b).reshape(-1, self.dim)[cell]

    def c(self, cell):
        c = self.spec.arrays["c"]
        return 0.0 if c is None else float(np.asarray(c).reshape(-1)[cell])

    # ---- coefficient fields at physical points X [P, dim] of cell e (call-backs if the spec has them) ----
    def fn(self, name):
        return (getattr(self.spec, "fns", None) or {}).get(name)

    def A_at(self, e, X):
        f = self.fn("A")
        if f is None:
            return np.broadcast_to(self.A(e), (len(X), self.dim, self.dim))
        return np.stack([np.asarray(f(x, e), dtype=float).reshape(self.dim, self.dim) for x in X])

    def b_at(self, e, X):
        f = self.fn("b")
        if f is None:
            return np.broadcast_to(self.b(e), (len(X), self.dim))
        return np.stack([np.asarray(f(x, e), dtype=float) for x in X])

    def c_at(self, e, X):
        f = self.fn("c")
        if f is None:
            return np.full(len(X), self.c(e))
        return np.array([float(f(x, e)) for x in X])

    def vol_points(self, c):
        """physical volume quadrature points of cell c, x fastest"""
        idx = np.array(list(itertools.product(*[range(self.m)] * self.dim)))[:, ::-1]
        return np.stack([self.spec.lower[d] + (c[d] + self.xq[idx[:, d]]) * self.h[d] for d in range(self.dim)], axis=1)

    def face_points(self, c, dirn, side):
        """physical quadrature points of face (dirn, side) of cell c, tangential directions increasing, first fastest"""
        tang = [d for d in range(self.dim) if d != dirn]
        idx = np.array(list(itertools.product(*[range(self.m)] * len(tang))))[:, ::-1].reshape(-1, len(tang))
        X = np.zeros((len(idx), self.dim))
        X[:, dirn] = self.spec.lower[dirn] + (c[dirn] + side) * self.h[dirn]
        for t, d in enumerate(tang):
            X[:, d] = self.spec.lower[d] + (c[d] + self.xq[idx[:, t]]) * self.h[d]
        return X

    def arr(self, name, shape):
        a = self.spec.arrays[name]
        return None if a is None else np.asarray(a).reshape(shape)

    # basis on a point set given per direction: returns phi[P, n], grad[P, n, dim] (physical)
    def basis(self, pts_per_dir):
        tabs = [basis_tables(self.k, p, getattr(self.spec, "basis", 0)) for p in pts_per_dir]
        # tensor over points: point index p = p0 + m0*(p1 + m1*p2) (x fastest), same for basis index
        dim = self.dim
        shapes = [len(p) for p in pts_per_dir]
        P = int(np.prod(shapes))
        phi = np.ones((P, self.n))
        grad = np.ones((P, self.n, dim))
        pidx = np.array(list(itertools.product(*[range(s) for s in reversed(shapes)])))[:, ::-1]
        bidx = np.array(list(itertools.product(*[range(self.n1)] * dim)))[:, ::-1]
        for d in range(dim):
            V, D = tabs[d]
            vd = V[pidx[:, d]][:, bidx[:, d]]
            dd = D[pidx[:, d]][:, bidx[:, d]] / self.h[d]
            phi *= vd
            for e in range(dim):
                grad[:, :, e] *= dd if e == d else vd
        return phi, grad

    def weights(self, dirs):
        """tensor-product weights over the listed directions, first listed direction fastest"""
        nd = len(dirs)
        if nd == 0:
            return np.ones(1)
        idx = np.array(list(itertools.product(*[range(self.m)] * nd)))[:, ::-1]
        return np.prod(self.wq[idx], axis=1)


def enumerate_qk_dofs(G):
    """Dictionary lattice point -> container index for conforming Q1/Q2, by listing entities."""
    dim, k, N = G.dim, G.k, G.N
    table = {}
    nxt = 0
    if k == 1:
        for v in itertools.product(*[range(n + 1) for n in reversed(N)]):
            table[tuple(reversed(v))] = nxt
            nxt += 1
        return table, nxt
    for edim in range(dim + 1):                      # vertices | edges | faces | cells
        for s in range(1 << dim):                    # extension bitset, ascending
            if bin(s).count("1") != edim:
                continue
            box = [N[d] if (s >> d) & 1 else N[d] + 1 for d in range(dim)]
            for a in itertools.product(*[range(b) for b in reversed(box)]):
                a = tuple(reversed(a))               # anchor, x fastest
                lat = tuple(2 * a[d] + ((s >> d) & 1) for d in range(dim))
                table[lat] = nxt
                nxt += 1
    return table, nxt


def dof_map(G):
    """global indices per cell [ncells, n] and number of DOFs."""
    spec = G.spec
    if spec.space == abi.SPACE_QKDG:
        return np.arange(G.ncells * G.n).reshape(G.ncells, G.n), G.ncells * G.n
    table, nd = enumerate_qk_dofs(G)
    out = np.zeros((G.ncells, G.n), dtype=np.int64)
    loc = np.array(list(itertools.product(*[range(G.n1)] * G.dim)))[:, ::-1]
    for c in G.cells():
        e = G.cell_index(c)
        for i, l in enumerate(loc):
            out[e, i] = table[tuple(G.k * c[d] + l[d] for d in range(G.dim))]
    return out, nd


def assemble(spec):
    """Dense J [ndofs, ndofs], r0 [ndofs], constrained flags [ndofs]."""
    G = Grid(spec)
    dim, m, n = G.dim, G.m, G.n
    dg = spec.space == abi.SPACE_QKDG
    dmap, ndofs = dof_map(G)
    J = np.zeros((ndofs, ndofs))
    r0 = np.zeros(ndofs)
    con = np.zeros(ndofs, dtype=bool)
    theta = {abi.DG_SIPG: -1.0, abi.DG_NIPG: 1.0, abi.DG_IIPG: 0.0}[spec.method]
    vol = float(np.prod(G.h))
    phi, grad = G.basis([G.xq] * dim)
    wv = G.weights(range(dim)) * vol
    f = G.arr("f", (G.ncells, m ** dim))
    bct = G.arr("bctype", (-1,))
    gq, jq, oq = (G.arr(nm, (-1, m ** (dim - 1))) for nm in ("g", "j", "o"))
    pen_k = G.k * (G.k + dim - 1)
    # trace tables: basis at face points of side 0 / 1 in direction d
    trace = {}
    for d in range(dim):
        for side in range(2):
            pts = [G.xq] * dim
            pts[d] = np.array([float(side)])
            trace[(d, side)] = G.basis(pts)
    for c in G.cells():
        e = G.cell_index(c)
        XV = G.vol_points(c)
        Aq, bq, cq = G.A_at(e, XV), G.b_at(e, XV), G.c_at(e, XV)
        ids = dmap[e]
        Agrad = np.einsum("pab,pjb->pja", Aq, grad)
        Kloc = np.einsum("p,pja,pia->ij", wv, Agrad, grad)
        Kloc -= np.einsum("p,pj,pa,pia->ij", wv, phi, bq, grad)
        Kloc += np.einsum("p,p,pj,pi->ij", wv, cq, phi, phi)
        J[np.ix_(ids, ids)] += Kloc
        if f is not None:
            r0[ids] -= np.einsum("p,p,pi->i", wv, f[e], phi)
        for d in range(dim):
            area = vol / G.h[d]
            wf = G.weights([x for x in range(dim) if x != d]) * area
            for side in range(2):
                onb = c[d] == (G.N[d] - 1 if side else 0)
                nrm = np.zeros(dim)
                nrm[d] = 1.0 if side else -1.0
                ps, gs = trace[(d, side)]
                if not onb:
                    if not dg or side == 1:
                        continue      # every interior face once, from its upper cell (s = larger index)
                    cn = list(c)
                    cn[d] -= 1
                    en = G.cell_index(cn)
                    idn = dmap[en]
                    XF = G.face_points(c, d, side)
                    As, An = G.A_at(e, XF), G.A_at(en, XF)            # both traces of A at the face points
                    pn, gn = trace[(d, 1)]
                    ds, dn = np.einsum("a,pab,b->p", nrm, As, nrm), np.einsum("a,pab,b->p", nrm, An, nrm)
                    if spec.weights == abi.DG_WEIGHTS_ON:
                        ws, wn = dn / (ds + dn + 1e-20), ds / (ds + dn + 1e-20)
                        harm = 2 * ds * dn / (ds + dn + 1e-20)
                    else:
                        ws = wn = np.full(len(XF), 0.5)
                        harm = np.ones(len(XF))
                    gamma = spec.alpha / G.h[d] * harm * pen_k
                    beta = G.b_at(e, XF) @ nrm     # velocity of the inside (larger-index) cell
                    # jump and average operators as row vectors over the 2n local DOFs [s | n]
                    jump = np.concatenate([ps, -pn], axis=1)                       # [P, 2n]
                    flux = np.concatenate([ws[:, None] * np.einsum("pja,pba,b->pj", gs, As, nrm),
                                           wn[:, None] * np.einsum("pja,pba,b->pj", gn, An, nrm)], axis=1)
                    ups = (beta >= 0)[:, None]
                    up = np.concatenate([np.where(ups, ps, 0.0), np.where(ups, 0.0, pn)], axis=1)
                    B = (np.einsum("p,p,pj,pi->ij", wf, beta, up, jump)
                         - np.einsum("p,pj,pi->ij", wf, flux, jump)
                         + theta * np.einsum("p,pj,pi->ij", wf, jump, flux)
                         + np.einsum("p,p,pj,pi->ij", wf, gamma, jump, jump))
                    both = np.concatenate([ids, idn])
                    J[np.ix_(both, both)] += B
                    continue
                if spec.side_kind[d][side] == abi.SIDE_PROCESSOR:
                    if dg:
                        con[ids] = True
                    else:
                        con[ids[np.abs(ps).sum(axis=0) > 1e-14]] = True
                    continue
                bf = G.bface_index(c, d, side)
                XF = G.face_points(c, d, side)
                beta = G.b_at(e, XF) @ nrm
                if not dg:
                    bt = abi.BC_DIRICHLET if bct is None else int(bct[bf])   # the type at the face centre
                    if bt == abi.BC_DIRICHLET:
                        con[ids[np.abs(ps).sum(axis=0) > 1e-14]] = True
                    elif bt == abi.BC_NEUMANN:
                        if jq is not None:
                            r0[ids] += np.einsum("p,p,pi->i", wf, jq[bf], ps)
                    elif bt == abi.BC_OUTFLOW:
                        J[np.ix_(ids, ids)] += np.einsum("p,p,pj,pi->ij", wf, beta, ps, ps)
                        if oq is not None:
                            r0[ids] += np.einsum("p,p,pi->i", wf, oq[bf], ps)
                    continue
                # DG: the boundary type is a function of the face point
                if G.fn("bctype") is not None:
                    btp = np.array([int(G.fn("bctype")(x)) for x in XF])
                else:
                    btp = np.full(len(XF), abi.BC_DIRICHLET if bct is None else int(bct[bf]))
                As = G.A_at(e, XF)
                ds = np.einsum("a,pab,b->p", nrm, As, nrm)
                harm = ds if spec.weights == abi.DG_WEIGHTS_ON else np.ones(len(XF))
                gamma = spec.alpha / G.h[d] * harm * pen_k
                fl = np.einsum("pja,pba,b->pj", gs, As, nrm)
                for kind in (abi.BC_NEUMANN, abi.BC_OUTFLOW, abi.BC_DIRICHLET):
                    w = np.where(btp == kind, wf, 0.0)   # quadrature restricted to the points of this type
                    if not w.any():
                        continue
                    if kind == abi.BC_NEUMANN:
                        if jq is not None:
                            r0[ids] += np.einsum("p,p,pi->i", w, jq[bf], ps)
                    elif kind == abi.BC_OUTFLOW:
                        J[np.ix_(ids, ids)] += np.einsum("p,p,pj,pi->ij", w, beta, ps, ps)
                        if oq is not None:
                            r0[ids] += np.einsum("p,p,pi->i", w, oq[bf], ps)
                    else:
                        B = (-np.einsum("p,pj,pi->ij", w, fl, ps) + theta * np.einsum("p,pj,pi->ij", w, ps, fl)
                             + np.einsum("p,p,pj,pi->ij", w, gamma, ps, ps))
                        B += np.einsum("p,p,pj,pi->ij", w, np.where(beta >= 0, beta, 0.0), ps, ps)
                        J[np.ix_(ids, ids)] += B
                        if gq is not None:
                            g = gq[bf]
                            r0[ids] -= theta * np.einsum("p,p,pi->i", w, g, fl) + np.einsum("p,p,p,pi->i", w, gamma, g, ps)
                            r0[ids] += np.einsum("p,p,p,pi->i", w, np.where(beta < 0, beta, 0.0), g, ps)
    return J, r0, con


def apply_constraints(J, r0, con):
    """Rows of constrained DOFs: identity in the matrix, zero in vectors."""
    Jc = J.copy()
    Jc[con, :] = 0.0
    Jc[con, con] = 1.0
    return Jc

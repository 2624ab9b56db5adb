#!/usr/bin/env python
"""Shared-memory wavefront count of dg_kron_3d_kernel<4> (csrc/dg_kron.cu) from its address arithmetic alone.

A 64-bit shared access of a warp is served per half-warp; a half-warp needs as many wavefronts as the largest number of
distinct 8-byte words that fall on one of the 16 bank pairs (same word = broadcast).  The script replays the addresses
of every LDS/STS of the kernel's sweeps and transposes for all warps of a CTA and prints ideal vs actual wavefronts -
ncu measured 63.7 M wavefronts of which 23.6 M are conflict replays (profiles/r02_dg_kron4_ncu_summary.json) - and
then searches for a lane -> (cell, plane) assignment that would make the z-sweep conflict-free (there is none with
cells stored densely, 125 doubles apart, which is what a TMA box delivers)."""
import itertools

K = 4
N1, NPL, NLOC = K + 1, (K + 1) ** 2, (K + 1) ** 3
import sys
TX, TY, TZ = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 2, 3)  # tile of KronCfg<4>
ROWX, CPW = TX + 4, 32 // N1
CELLS = TX * TY * TZ
WARPS = CELLS // CPW


def al(d):
    return (d + 15) // 16 * 16


R0 = 0
R1 = al(R0 + TZ * TY * ROWX * NLOC)
R2 = al(R1 + TZ * TX * NLOC)
R3 = al(R2 + TZ * TX * NLOC)
R4 = al(R3 + TY * TX * NLOC)


def wavefronts(addrs):
    """addrs: word address (doubles) per lane or None; returns (ideal, actual) wavefronts of one warp instruction."""
    ideal = actual = 0
    for h in (addrs[:16], addrs[16:]):
        words = {a for a in h if a is not None}
        if not words:
            continue
        per_bank = {}
        for w in words:
            per_bank.setdefault(w % 16, set()).add(w)
        ideal += 1
        actual += max(len(v) for v in per_bank.values())
    return ideal, actual


def cell_bases(warp):
    out = []
    for lane in range(32):
        cw, s = divmod(lane, N1)
        if cw >= CPW:
            out.append(None)
            continue
        ci = warp * CPW + cw
        cx, cy, cz = ci % TX, (ci // TX) % TY, ci // (TX * TY)
        so = R0 + ((cz * TY + cy) * ROWX + cx + 2) * NLOC
        nb = dict(o=so, xl=so - NLOC, xr=so + NLOC,
                  yl=so - ROWX * NLOC if cy > 0 else R1 + (cz * TX + cx) * NLOC,
                  yr=so + ROWX * NLOC if cy < TY - 1 else R2 + (cz * TX + cx) * NLOC,
                  zl=so - TY * ROWX * NLOC if cz > 0 else R3 + (cy * TX + cx) * NLOC,
                  zr=so + TY * ROWX * NLOC if cz < TZ - 1 else R4 + (cy * TX + cx) * NLOC,
                  scratch=CELLS * NLOC + ci * NLOC, stage=((cz * TY + cy) * TX + cx) * NLOC)
        out.append((s, nb))
    return out


def count(name, fn):
    ideal = actual = 0
    for warp in range(WARPS):
        lanes = cell_bases(warp)
        for ln in range(N1):
            for j in range(N1):
                for a in fn(ln, j):
                    addrs = [None if l is None else l[1][a[0]] + a[1](l[0]) for l in lanes]
                    i, w = wavefronts(addrs)
                    ideal += i
                    actual += w
    print(f"{name:34s} ideal {ideal:6d}  actual {actual:6d}  replays {100.0 * (actual - ideal) / actual:5.1f} %")
    return ideal, actual


tot = [0, 0]
for name, fn in [
    ("x-sweep (z-plane layout)", lambda ln, j: [(k, lambda s, ln=ln, j=j: s * NPL + ln * N1 + j) for k in ("o", "xl", "xr")]),
    ("y-sweep (z-plane layout)", lambda ln, j: [(k, lambda s, ln=ln, j=j: s * NPL + ln + j * N1) for k in ("o", "yl", "yr")]),
    ("z-sweep (y-plane layout)", lambda ln, j: [(k, lambda s, ln=ln, j=j: s * N1 + ln + j * NPL) for k in ("o", "zl", "zr")]),
    ("tz -> scratch (y-plane layout)", lambda ix, iz: [("scratch", lambda s, ix=ix, iz=iz: iz * NPL + s * N1 + ix)]),
    ("t += scratch, scratch = t (x2)", lambda a, b: [("scratch", lambda s, a=a, b=b: s * NPL + a * N1 + b)] * 2),
    ("u = scratch (y-plane layout)", lambda ix, iz: [("scratch", lambda s, ix=ix, iz=iz: iz * NPL + s * N1 + ix)]),
    ("stage = u (y-plane layout)", lambda ix, iz: [("stage", lambda s, ix=ix, iz=iz: iz * NPL + s * N1 + ix)]),
]:
    i, a = count(name, fn)
    tot[0] += i
    tot[1] += a
print(f"{'all of the above':34s} ideal {tot[0]:6d}  actual {tot[1]:6d}  replays {100.0 * (tot[1] - tot[0]) / tot[1]:5.1f} %   (per CTA)")

# Is there a conflict-free z-sweep?  Three cells of a half-warp at word offsets a_i = 125 c_i (mod 16 = 13 c_i), five
# planes each at 5 s (y-plane layout): the sets a_i + {0, 5, 10, 15, 20} must be pairwise disjoint mod 16.
D = {(5 * s) % 16 for s in range(N1)}
ok_z = {d for d in range(16) if not (D & {(x + d) % 16 for x in D})}
Dxy = {(NPL * s) % 16 for s in range(N1)}
ok_xy = {d for d in range(16) if not (Dxy & {(x + d) % 16 for x in Dxy})}
print("cell-offset differences (mod 16) that keep two cells apart: z-sweep", sorted(ok_z), " x/y-sweeps", sorted(ok_xy))
both = ok_z & ok_xy
triples = [t for t in itertools.combinations(range(16), 3)
           if all(((b - a) % 16) in both for a, b in itertools.combinations(t, 2))]
print("triples of cell offsets that are conflict-free in all three sweeps:", triples or "none")

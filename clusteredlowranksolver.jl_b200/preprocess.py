"""Linearly dependent constraints: the part of `preprocess!` (src/pre_postprocessing.jl:4-137, 183-213,
278-310) that needs heavy arithmetic, on top of the library's column-pivoted QR (`clrs_mp_qr_pivot`, on the device).

`find_linear_dependencies` of the reference vectorises the PSD part of every constraint (columns = constraints,
`vectorize_constraint`, :138-180), takes `qr(mpsd, ColumnNorm())` in BigFloat (:36) and reads the dependent constraints off
the diagonal of R (|R_ii| < tol, :39-44).  A dependent constraint is a combination of the kept ones in its PSD part
(`Rp = R11 \\ R12`, :52); the same combination of the free-variable parts and right-hand sides must vanish, otherwise it
is a linear constraint between the free variables (:64-66).  This module covers the case the path needs before the solver
runs: the combination vanishes (the constraint is redundant: it is removed) or it is `0 = b` with `b != 0` (the SDP is
infeasible: the reference raises, :89-99).  Eliminating free variables (`remove_lindep_freevars!`, :215-276) rewrites the
modelling layer's objects and stays in Julia; such an SDP raises `NotImplementedError` here.
"""
from __future__ import annotations

from typing import List, Tuple

import mpmath
import numpy as np

from . import wire
from .sdp import ClusteredSDP, Cluster, PSDBlock, LowRankTerm


def _block_matrices(blk: PSDBlock, p: int, prec: int):
    """{(r, s): delta x delta mpf matrix} of constraint row p in this block (missing subblocks are zero)."""
    d = blk.delta
    out = {}
    if blk.high_rank:
        A = None
        if p in blk.dense:
            A = wire.from_wire(np.asarray(blk.dense[p]), prec)
        elif p in blk.sparse:
            rows, cols, vals, mirror = blk.sparse[p]
            A = np.full((d, d), mpmath.mpf(0), dtype=object)
            for r, c, v in zip(rows, cols, wire.from_wire(vals, prec)):
                A[r, c] = v
                if mirror and r != c:
                    A[c, r] = v
        if A is not None:
            out[(0, 0)] = A
        return out
    for t in blk.lowrank:
        if t.p != p:
            continue
        lam, vs, ws = (wire.from_wire(a, prec) for a in (t.lam, t.vs, t.ws))
        M = out.setdefault((t.r, t.s), np.full((d, d), mpmath.mpf(0), dtype=object))
        for k in range(len(lam)):
            for a in range(d):
                for b in range(d):
                    M[a, b] += lam[k] * vs[k][a] * ws[k][b]
    return out


def vectorize_constraints(sdp: ClusteredSDP):
    """(mpsd as a wire matrix, rows = packed PSD entries of all blocks, columns = constraints; [(j, p)] per column),
    `vectorize_constraint` (src/pre_postprocessing.jl:138-180): per block and subblock pair s <= r the sum of the (r,s) and (s,r)
    subblocks (all delta^2 entries) off the diagonal, the packed lower triangle with doubled off-diagonal entries on it."""
    prec = sdp.prec
    cols: List[Tuple[int, int]] = [(j, p) for j, c in enumerate(sdp.clusters) for p in range(c.P)]
    with mpmath.workprec(prec + 64):
        offs, total = {}, 0
        for j, c in enumerate(sdp.clusters):
            for l, blk in enumerate(c.blocks):
                offs[(j, l)] = total
                d, m = blk.delta, blk.m
                total += m * (d * (d + 1) // 2) + (m * (m - 1) // 2) * d * d
        M = np.full((max(total, 1), max(len(cols), 1)), mpmath.mpf(0), dtype=object)
        for col, (j, p) in enumerate(cols):
            for l, blk in enumerate(sdp.clusters[j].blocks):
                mats = _block_matrices(blk, p, prec)
                if not mats:
                    continue
                d, k = blk.delta, offs[(j, l)]
                zero = np.full((d, d), mpmath.mpf(0), dtype=object)
                for r in range(blk.m):
                    for s in range(r + 1):
                        if r != s:
                            tot = mats.get((r, s), zero) + mats.get((s, r), zero)
                            for i, v in enumerate(tot.reshape(-1)):
                                M[k + i, col] = v
                            k += d * d
                        else:
                            A = mats.get((r, r), zero)
                            for i1 in range(d):
                                for i2 in range(i1 + 1):
                                    M[k, col] = A[i1, i2] if i1 == i2 else A[i1, i2] + A[i2, i1]
                                    k += 1
        return wire.to_wire(M.tolist(), prec), cols


def find_dependent_constraints(sdp: ClusteredSDP, solver, tol=None):
    """[(j, p)] of the constraints that are linear combinations of the others (and consistent), through `solver.mp_qr_pivot`
    (any handle of the right precision: the QR does not touch the handle's SDP)."""
    prec = sdp.prec
    mpsd, cols = vectorize_constraints(sdp)
    if not cols:
        return []
    R, perm = solver.mp_qr_pivot(mpsd)
    with mpmath.workprec(prec + 64):
        tol = mpmath.sqrt(mpmath.mpf(2) ** (1 - prec)) if tol is None else mpmath.mpf(tol)          # sqrt(eps(BigFloat)), :4
        Rm = wire.from_wire(R, prec)
        kmax, n = Rm.shape
        istart = next((i for i in range(kmax) if abs(Rm[i, i]) < tol), None)
        if istart is None:
            if kmax == n:
                return []
            istart = kmax                                   # more constraints than PSD entries (:41-44)
        dep = list(range(istart, n))
        # Rp = R11 \ R12: constraint perm[c] (c in dep) = sum_i Rp[i][c] * constraint perm[i] in the PSD part (:52)
        R11 = mpmath.matrix([[Rm[i, k] for k in range(istart)] for i in range(istart)]) if istart else None
        N = sdp.N
        rows = []                                           # [B | c] of every constraint, in the pivoted order
        for col in perm:
            j, p = cols[int(col)]
            cl = sdp.clusters[j]
            rows.append(list(wire.from_wire(cl.B[p], prec)) + [wire.from_wire(cl.c[p], prec)] if N else [wire.from_wire(cl.c[p], prec)])
        for c in dep:
            comb = mpmath.lu_solve(R11, mpmath.matrix([Rm[i, c] for i in range(istart)])) if istart else []
            resid = [sum(comb[i] * rows[i][k] for i in range(istart)) - rows[c][k] for k in range(N + 1)]
            scale = max([mpmath.mpf(1)] + [abs(v) for v in rows[c]])
            if any(abs(v) > tol * scale for v in resid[:N]):
                raise NotImplementedError("linearly dependent constraints that relate free variables: remove_lindep_freevars! stays in Julia "
                                          "(src/pre_postprocessing.jl:215-276)")
            if abs(resid[N]) > tol * scale:
                raise ValueError("Linear dependent constraint(s) resulting in a constraint 0 = b_i with b_i nonzero.")      # :91-99
        return [cols[int(perm[c])] for c in dep]


def remove_constraints(sdp: ClusteredSDP, cs) -> ClusteredSDP:
    """`remove_lindep_constraints!` (src/pre_postprocessing.jl:183-213) on the compact-row container: rows of B_j and c_j are dropped and
    the remaining constraints are renumbered (the container's indices ARE the compact rows that cs_map produces, src/solver.jl:156-167)."""
    drop = {}
    for j, p in cs:
        drop.setdefault(j, set()).add(p)
    clusters = []
    for j, c in enumerate(sdp.clusters):
        if j not in drop:
            clusters.append(c)
            continue
        keep = [p for p in range(c.P) if p not in drop[j]]
        new = {p: i for i, p in enumerate(keep)}
        blocks = []
        for blk in c.blocks:
            nb = PSDBlock(m=blk.m, delta=blk.delta, high_rank=blk.high_rank, C=blk.C, name=blk.name)
            nb.dense = {new[p]: A for p, A in blk.dense.items() if p in new}
            nb.sparse = {new[p]: A for p, A in blk.sparse.items() if p in new}
            nb.lowrank = [LowRankTerm(t.r, t.s, new[t.p], t.lam, t.vs, t.ws) for t in blk.lowrank if t.p in new]
            blocks.append(nb)
        clusters.append(Cluster(B=np.ascontiguousarray(c.B[keep]), c=np.ascontiguousarray(c.c[keep]), blocks=blocks))
    return ClusteredSDP(prec=sdp.prec, maximize=sdp.maximize, constant=sdp.constant, b=sdp.b, clusters=clusters, name=sdp.name)


def preprocess(sdp: ClusteredSDP, solver, tol=None):
    """(sdp without its redundant constraints, [(j, p)] removed) — `preprocess!` (src/pre_postprocessing.jl:278-310) for the
    constraint part; the dual variables of removed constraints are zero in the solution of the original SDP (`postprocess`, :312-325)."""
    cs = find_dependent_constraints(sdp, solver, tol)
    return (remove_constraints(sdp, cs) if cs else sdp), cs

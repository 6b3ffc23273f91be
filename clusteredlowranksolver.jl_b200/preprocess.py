"""Linearly dependent constraints: the part of `preprocess!` (src/pre_postprocessing.jl:4-137, 183-213,
278-310) that needs heavy arithmetic, on top of the library's column-pivoted QR (`clrs_mp_qr_pivot`, on the device).

`find_linear_dependencies` of the reference vectorises the PSD part of every constraint (columns = constraints,
`vectorize_constraint`, :138-180), takes `qr(mpsd, ColumnNorm())` in BigFloat (:36) and reads the dependent constraints off
the diagonal of R (|R_ii| < tol, :39-44).  A dependent constraint is a combination of the kept ones in its PSD part
(`Rp = R11 \\ R12`, :52); the same combination of the free-variable parts and right-hand sides must vanish, otherwise it
is a linear equation between the free variables (:64-66): the
equations are solved for as many variables as their rank (`nf_vars`, expressed through the remaining `ff_vars`, :67-119), the
substitution is applied to B_j, c_j, b and the constant (`remove_lindep_freevars!`, :215-236), and free variables whose columns
became linearly dependent are set to zero (:121-136).  `0 = b` with `b != 0` raises like the reference (:89-99).
The big factorisation — the (sum n(n+1)/2) x P matrix of the PSD parts — runs on the device; the two small ones (as many rows as
dependent constraints, N columns) are done here in mpmath.  `postprocess` puts the removed duals / variables back (:238-276, 312-325).
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


def _qr_pivot_host(M, tol):
    """Column-pivoted modified Gram-Schmidt of a small mpmath matrix: (Q (m x rank), R (rank x n, pivoted order), perm, rank)."""
    m, n = M.rows, M.cols
    A = M.copy()
    perm = list(range(n))
    norms = [sum(A[i, j] ** 2 for i in range(m)) for j in range(n)]
    Q, Rrows = [], []
    for k in range(min(m, n)):
        p = max(range(k, n), key=lambda j: (norms[j], -j))
        if mpmath.sqrt(norms[p]) < tol:
            break
        if p != k:
            for i in range(m):
                A[i, k], A[i, p] = A[i, p], A[i, k]
            for row in Rrows:
                row[k], row[p] = row[p], row[k]
            norms[k], norms[p] = norms[p], norms[k]
            perm[k], perm[p] = perm[p], perm[k]
        r = mpmath.sqrt(norms[k])
        q = [A[i, k] / r for i in range(m)]
        row = [mpmath.mpf(0)] * n
        row[k] = r
        for j in range(k + 1, n):
            d = sum(q[i] * A[i, j] for i in range(m))
            row[j] = d
            for i in range(m):
                A[i, j] -= d * q[i]
            norms[j] = sum(A[i, j] ** 2 for i in range(m))
        Q.append(q)
        Rrows.append(row)
    return Q, Rrows, perm, len(Rrows)


def _upper_solve(Rrows, rank, rhs_cols):
    """inv(R11) rhs for the leading rank x rank triangle; rhs_cols = list of columns (each of length rank)."""
    out = []
    for col in rhs_cols:
        x = [mpmath.mpf(0)] * rank
        for i in range(rank - 1, -1, -1):
            x[i] = (col[i] - sum(Rrows[i][k] * x[k] for k in range(i + 1, rank))) / Rrows[i][i]
        out.append(x)
    return out


def find_linear_dependencies(sdp: ClusteredSDP, solver, tol=None):
    """`find_linear_dependencies` (src/pre_postprocessing.jl:4-137).  Returns (cs, var_rels) with cs = [(j, p)] of the constraints to
    remove and var_rels = dict(fv_zeros, fv_nonzeros, Rref, rhs_changed, nf_vars, ff_vars) describing the substitution
    y[nf_vars] = rhs_changed - Rref y[ff_vars], after which the columns fv_zeros (indices into ff_vars) are set to zero."""
    prec, N = sdp.prec, sdp.N
    mpsd, cols = vectorize_constraints(sdp)
    none = dict(fv_zeros=[], fv_nonzeros=list(range(N)), Rref=[], rhs_changed=[], nf_vars=[], ff_vars=list(range(N)))
    if not cols:
        return [], none
    R, perm = solver.mp_qr_pivot(mpsd)                                  # the heavy factorisation: on the device
    with mpmath.workprec(prec + 64):
        tol = mpmath.sqrt(mpmath.mpf(2) ** (1 - prec)) if tol is None else mpmath.mpf(tol)          # sqrt(eps(BigFloat)), :4
        Rm = wire.from_wire(R, prec)
        kmax, n = Rm.shape
        istart = next((i for i in range(kmax) if abs(Rm[i, i]) < tol), None)
        if istart is None:
            istart = kmax if kmax < n else n                 # more constraints than PSD entries (:41-44), or nothing dependent
        dep = list(range(istart, n))
        perm = [int(v) for v in perm]
        Bc = []                                              # [B | c] of every constraint, in the pivoted order
        for col in perm:
            j, p = cols[col]
            cl = sdp.clusters[j]
            Bc.append((list(wire.from_wire(cl.B[p], prec)) if N else []) + [wire.from_wire(cl.c[p], prec)])
        # Rp = R11 \ R12: in its PSD part, constraint perm[c] = sum_i Rp[i][c] constraint perm[i]  (:52)
        Rrows = [[Rm[i, k] for k in range(n)] for i in range(istart)]
        Rp = _upper_solve(Rrows, istart, [[Rm[i, c] for i in range(istart)] for c in dep])
        # the same combination of the free parts: F y = h  (:64-66)
        FH = mpmath.matrix(max(len(dep), 1), N + 1)
        for a, c in enumerate(dep):
            for k in range(N + 1):
                FH[a, k] = sum(Rp[a][i] * Bc[i][k] for i in range(istart)) - Bc[c][k]
        cs = [cols[perm[c]] for c in dep]
        scale = max([mpmath.mpf(1)] + [abs(v) for row in Bc for v in row])
        ftol = tol * scale
        rel = dict(none)
        if dep and N:
            F = FH[:len(dep), :N]
            Q, Rr, p2, rank = _qr_pivot_host(F, ftol)
            qh = [sum(Q[i][a] * FH[a, N] for a in range(len(dep))) for i in range(rank)]           # Q^T h
            resid = [FH[a, N] - sum(Q[i][a] * qh[i] for i in range(rank)) for a in range(len(dep))]
            if any(abs(v) > ftol for v in resid):
                raise ValueError("Linear dependent constraint(s) resulting in a constraint 0 = b_i with b_i nonzero.")      # :89-99
            if rank:
                nf, ff = p2[:rank], p2[rank:]
                Rref = _upper_solve(Rr, rank, [[Rr[i][c] for i in range(rank)] for c in range(rank, N)])               # columns: ff variables
                rel.update(nf_vars=nf, ff_vars=ff, rhs_changed=_upper_solve(Rr, rank, [qh])[0],
                           Rref=[[Rref[c][i] for c in range(len(ff))] for i in range(rank)])                        # rank x len(ff)
        elif dep:
            if any(abs(FH[a, N]) > ftol for a in range(len(dep))):
                raise ValueError("Linear dependent constraint(s) resulting in a constraint 0 = b_i with b_i nonzero.")
        # free variables whose columns are linearly dependent after the substitution can be set to zero  (:121-136)
        if N:
            keep = [i for i in range(n) if i < istart]
            ff, nf = rel["ff_vars"], rel["nf_vars"]
            cm = _changemat(rel, N)
            Bnew = mpmath.matrix(max(len(keep), 1), max(len(ff), 1))
            for a, i in enumerate(keep):
                for c in range(len(ff)):
                    Bnew[a, c] = sum(Bc[i][v] * cm[v][c] for v in range(N))
            if keep and ff:
                _, _, p3, rank3 = _qr_pivot_host(Bnew[:len(keep), :len(ff)], ftol)
                rel["fv_zeros"] = sorted(p3[rank3:])
                rel["fv_nonzeros"] = [i for i in range(len(ff)) if i not in rel["fv_zeros"]]
            else:
                rel["fv_zeros"], rel["fv_nonzeros"] = ([], list(range(len(ff)))) if keep else (list(range(len(ff))), [])
        return cs, rel


def _changemat(rel, N):
    """N x len(ff_vars): y = changemat y_ff + shift, with y[nf] = rhs_changed - Rref y_ff and y[ff] = y_ff  (:123, 226)."""
    nf, ff = rel["nf_vars"], rel["ff_vars"]
    cm = [[mpmath.mpf(0)] * len(ff) for _ in range(N)]
    for a, v in enumerate(nf):
        for c in range(len(ff)):
            cm[v][c] = -rel["Rref"][a][c]
    for c, v in enumerate(ff):
        cm[v][c] = mpmath.mpf(1)
    return cm


def find_dependent_constraints(sdp: ClusteredSDP, solver, tol=None):
    """[(j, p)] of the linearly dependent constraints (the first return value of `find_linear_dependencies`)."""
    return find_linear_dependencies(sdp, solver, tol)[0]


def remove_free_variables(sdp: ClusteredSDP, rel) -> ClusteredSDP:
    """`remove_lindep_freevars!` (src/pre_postprocessing.jl:215-236): substitute y[nf] = rhs_changed - Rref y[ff], drop the columns fv_zeros."""
    N, prec = sdp.N, sdp.prec
    if not rel["nf_vars"] and not rel["fv_zeros"]:
        return sdp
    with mpmath.workprec(prec + 64):
        cm = _changemat(rel, N)
        nf, nz = rel["nf_vars"], rel["fv_nonzeros"]
        shift = rel["rhs_changed"]
        b = list(wire.from_wire(sdp.b, prec))
        clusters = []
        for c in sdp.clusters:
            B = wire.from_wire(c.B, prec).reshape(c.P, N)
            cv = list(wire.from_wire(c.c, prec))
            newc = [cv[p] - sum(B[p, v] * shift[a] for a, v in enumerate(nf)) for p in range(c.P)]
            newB = [[sum(B[p, v] * cm[v][k] for v in range(N)) for k in nz] for p in range(c.P)]
            clusters.append(Cluster(B=wire.to_wire(newB, prec) if nz else wire.wire_zeros((c.P, 0), prec), c=wire.to_wire(newc, prec), blocks=c.blocks))
        const = wire.from_wire(sdp.constant, prec) + sum(b[v] * shift[a] for a, v in enumerate(nf))
        newb = [sum(cm[v][k] * b[v] for v in range(N)) for k in nz]
        return ClusteredSDP(prec=prec, maximize=sdp.maximize, constant=wire.to_wire(const, prec),
                            b=wire.to_wire(newb, prec) if nz else wire.wire_zeros((0,), prec), clusters=clusters, name=sdp.name)


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
    """`preprocess!` (src/pre_postprocessing.jl:278-310): (reduced sdp, cs, var_rels).  `solver` is any handle of the SDP's precision: its
    `mp_qr_pivot` does the large factorisation (on the device for lib="device")."""
    cs, rel = find_linear_dependencies(sdp, solver, tol)
    new = remove_constraints(sdp, cs) if cs else sdp
    new = remove_free_variables(new, rel)
    return new, cs, rel


def postprocess(sdp: ClusteredSDP, x, y, cs, rel):
    """`postprocess` (src/pre_postprocessing.jl:238-276, 312-325) for the ORIGINAL sdp: x gets a zero for every removed constraint (dual
    variables, in (j, p) order), y the removed free variables.  x, y: lists of mpf of the reduced SDP."""
    removed = set(cs)
    it = iter(x)
    xfull = [mpmath.mpf(0) if (j, p) in removed else next(it) for j, c in enumerate(sdp.clusters) for p in range(c.P)]
    ff, nf = rel["ff_vars"], rel["nf_vars"]
    ity = iter(y)
    yff = [mpmath.mpf(0) if k in rel["fv_zeros"] else next(ity) for k in range(len(ff))]
    yfull = [mpmath.mpf(0)] * sdp.N
    for k, v in enumerate(ff):
        yfull[v] = yff[k]
    for a, v in enumerate(nf):
        yfull[v] = rel["rhs_changed"][a] - sum(rel["Rref"][a][k] * yff[k] for k in range(len(ff)))
    return xfull, yfull

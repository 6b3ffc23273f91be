"""SDPA-sparse input -> `ClusteredSDP` with triplet-form constraint matrices.

Mirror of `read_sdpa_sparse_file` / `sdpa_sparse_to_problem` (src/SDPAtoCLRS.jl:3-83) followed by the clustering of
`ClusteredLowRankSDP(problem)` (src/interface.jl:850-905):

    maximise  <F_0, Y> + obj_shift   subject to   <F_i, Y> = c_i  (i = 1..m),   Y PSD block diagonal

* a block of positive size b is one dense PSD variable; a "diagonal" block of size -b is b separate 1 x 1 variables
  (src/SDPAtoCLRS.jl:12,20-27,35-37);
* constraints without any matrix are dropped (src/SDPAtoCLRS.jl:66-80);
* constraints that share a PSD variable end up in the same cluster (connected components).

The constraint matrices never become dense on the host: every F_i block is kept as (rows, cols, values) and reaches the
device through `clrs_add_sparse_term` (only the triplets cross PCIe).  Together with `Solver(sparse_schur=True)` a sparse
SDPA problem is solved without any dense n x n x P intermediate.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import mpmath
import numpy as np

from . import wire
from .sdp import ClusteredSDP, Cluster, PSDBlock


def parse_sdpa_sparse(text: str):
    """-> (m, blocksizes, c (strings), entries {(cidx, var) -> {(i, j) -> str}}) with var = (bidx,) or (bidx, sub), 1-based
    like the file.  Leading comment lines (not starting with a digit) are skipped as in src/SDPAtoCLRS.jl:6-8."""
    lines = [ln.split() for ln in text.splitlines() if ln.strip()]
    i = 0
    while not lines[i][0][0].isdigit():
        i += 1
    m = int(lines[i][0]); i += 1
    nblocks = int(lines[i][0]); i += 1
    blocksizes = [int(t) for t in re.sub(r"[{}(),]", " ", " ".join(lines[i])).split()]; i += 1
    if len(blocksizes) != nblocks:
        raise ValueError(f"SDPA file: {nblocks} blocks announced, {len(blocksizes)} sizes given")
    c = re.sub(r"[{}(),]", " ", " ".join(lines[i])).split(); i += 1
    if len(c) != m:
        raise ValueError(f"SDPA file: {m} constraints announced, {len(c)} right-hand sides given")
    entries: Dict[Tuple[int, tuple], Dict[Tuple[int, int], str]] = {}
    for ln in lines[i:]:
        cidx, bidx, r, s = (int(t) for t in ln[:4])
        if not (0 <= cidx <= m and 1 <= bidx <= nblocks):
            raise ValueError(f"SDPA file: entry {ln} out of range")
        b = blocksizes[bidx - 1]
        if b < 0:
            if r != s:
                raise ValueError("SDPA file: off-diagonal entry in a diagonal block")
            var, pos = (bidx, r), (1, 1)
        else:
            var, pos = (bidx,), (min(r, s), max(r, s))
        if not (1 <= r <= abs(b) and 1 <= s <= abs(b)):
            raise ValueError(f"SDPA file: entry {ln} outside its block")
        entries.setdefault((cidx, var), {})[pos] = ln[4]          # a repeated position overwrites, as the reference's assignment does
    return m, blocksizes, c, entries


def read_sdpa_sparse(path: str, prec: int = 256, obj_shift=0, float64: bool = False) -> ClusteredSDP:
    with open(path) as f:
        return sdpa_sparse_to_sdp(f.read(), prec=prec, obj_shift=obj_shift, float64=float64, name=path)


def sdpa_sparse_to_sdp(text: str, prec: int = 256, obj_shift=0, float64: bool = False, name: str = "sdpa") -> ClusteredSDP:
    """float64 = True parses the numbers as Float64 first (the reference's default T, src/SDPAtoCLRS.jl:3)."""
    m, blocksizes, cstr, entries = parse_sdpa_sparse(text)
    with mpmath.workprec(prec + 64):
        num = (lambda t: mpmath.mpf(float(t))) if float64 else (lambda t: mpmath.mpf(t))
        # drop explicit zeros: the reference keeps a matrix only if it is not all zero (src/SDPAtoCLRS.jl:56-63)
        mats: Dict[Tuple[int, tuple], Dict[Tuple[int, int], object]] = {}
        for key, d in entries.items():
            nz = {pos: num(v) for pos, v in d.items()}
            nz = {pos: v for pos, v in nz.items() if v != 0}
            if nz:
                mats[key] = nz
        cons = [i for i in range(1, m + 1) if any(k[0] == i for k in mats)]          # empty constraints are removed
        cvars = {i: sorted(k[1] for k in mats if k[0] == i) for i in cons}
        used = sorted({v for i in cons for v in cvars[i]})
        for (ci, v) in mats:
            if ci == 0 and v not in used:
                raise ValueError(f"SDPA file: block {v} appears in the objective only (unbounded)")
        # connected components: constraints sharing a PSD variable belong to one cluster
        parent = {v: v for v in used}

        def find(v):
            while parent[v] != v:
                parent[v] = parent[parent[v]]
                v = parent[v]
            return v
        for i in cons:
            for v in cvars[i][1:]:
                parent[find(v)] = find(cvars[i][0])
        roots: List[tuple] = []
        for i in cons:
            r = find(cvars[i][0])
            if r not in roots:
                roots.append(r)
        size = lambda v: 1 if len(v) == 2 else blocksizes[v[0] - 1]

        def triplets(d):
            rows = np.array([p[0] - 1 for p in d], dtype=np.int32)
            cols = np.array([p[1] - 1 for p in d], dtype=np.int32)
            return rows, cols, wire.to_wire(list(d.values()), prec), True
        clusters = []
        for r in roots:
            ci = [i for i in cons if find(cvars[i][0]) == r]
            vs = [v for v in used if find(v) == r]
            blocks = []
            for v in vs:
                n = size(v)
                Cw = wire.wire_zeros((n, n), prec)
                for (a, b), val in mats.get((0, v), {}).items():
                    w = wire.to_wire(val, prec)[()]
                    Cw[a - 1, b - 1] = w
                    Cw[b - 1, a - 1] = w
                blk = PSDBlock(m=1, delta=n, high_rank=True, C=Cw, name=v)
                for row, i in enumerate(ci):
                    if (i, v) in mats:
                        blk.sparse[row] = triplets(mats[(i, v)])
                blocks.append(blk)
            cw = wire.to_wire([num(cstr[i - 1]) for i in ci], prec)
            clusters.append(Cluster(B=wire.wire_zeros((len(ci), 0), prec), c=cw, blocks=blocks))
        return ClusteredSDP(prec=prec, maximize=True, constant=wire.to_wire(mpmath.mpf(obj_shift), prec),
                            b=wire.wire_zeros((0,), prec), clusters=clusters, name=name)


def write_sdpa_sparse(blocksizes, c, entries, comment: str = "") -> str:
    """entries: iterable of (cidx, bidx, i, j, value) with 1-based indices, upper triangle."""
    out = []
    if comment:
        out.append('"' + comment + '"')
    out.append(str(len(c)))
    out.append(str(len(blocksizes)))
    out.append(" ".join(str(b) for b in blocksizes))
    out.append(" ".join(str(v) for v in c))
    for e in entries:
        out.append(" ".join(str(t) for t in e))
    return "\n".join(out) + "\n"


def maxcut_sdpa_text(L) -> str:
    """The Goemans-Williamson relaxation of README.md:39-65 as an SDPA-sparse file: max <L/4, Y>, Y_ii = 1."""
    L = np.asarray(L)
    n = L.shape[0]
    ent = []
    for i in range(n):
        for j in range(i, n):
            v = int(L[i, j])
            if v:
                ent.append((0, 1, i + 1, j + 1, f"{v / 4}"))
    for i in range(n):
        ent.append((i + 1, 1, i + 1, i + 1, "1"))
    return write_sdpa_sparse([n], ["1"] * n, ent, comment=f"maxcut n={n}")

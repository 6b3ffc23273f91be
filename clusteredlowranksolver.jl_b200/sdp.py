"""Host-side mirror of the reference's canonical SDP container.

`ClusteredSDP` holds exactly what `ClusteredLowRankSDP` holds
(src/interface.jl:807-819): maximize, constant, A[j][l][r,s][p], B[j], c[j],
C[j][l], b — with every number already in wire format at `prec` bits
(what `convert_to_prec`, src/interface.jl:1078-1112, produces).  Constraint
indices p are 0-based compact rows inside their cluster (cs_map applied).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from . import wire


@dataclass
class LowRankTerm:
    r: int
    s: int
    p: int
    lam: np.ndarray   # (rank,) wire
    vs: np.ndarray    # (rank, delta) wire
    ws: np.ndarray    # (rank, delta) wire


@dataclass
class PSDBlock:
    m: int
    delta: int
    high_rank: bool
    C: np.ndarray                                   # (n, n) wire
    dense: Dict[int, np.ndarray] = field(default_factory=dict)      # p -> (n, n) wire
    sparse: Dict[int, tuple] = field(default_factory=dict)          # p -> (rows, cols, vals (nnz,) wire, mirror): triplet form of a dense term
    lowrank: List[LowRankTerm] = field(default_factory=list)
    name: object = None

    @property
    def n(self) -> int:
        return self.m * self.delta


@dataclass
class Cluster:
    B: np.ndarray        # (P, N) wire
    c: np.ndarray        # (P,) wire
    blocks: List[PSDBlock] = field(default_factory=list)

    @property
    def P(self) -> int:
        return int(self.c.shape[0])


@dataclass
class ClusteredSDP:
    prec: int
    maximize: bool
    constant: np.ndarray   # 0-d wire
    b: np.ndarray          # (N,) wire
    clusters: List[Cluster] = field(default_factory=list)
    name: str = ""

    @property
    def N(self) -> int:
        return int(self.b.shape[0])

    @property
    def num_constraints(self) -> int:
        return sum(c.P for c in self.clusters)

    def block_shapes(self) -> List[Tuple[int, int, int]]:
        return [(j, l, blk.n) for j, c in enumerate(self.clusters) for l, blk in enumerate(c.blocks)]

    def describe(self) -> str:
        J = len(self.clusters)
        blocks = [blk.n for c in self.clusters for blk in c.blocks]
        return (f"{self.name}: J={J} P={self.num_constraints} N={self.N} blocks={len(blocks)} "
                f"K={sum(blocks)} max_n={max(blocks) if blocks else 0} prec={self.prec}")

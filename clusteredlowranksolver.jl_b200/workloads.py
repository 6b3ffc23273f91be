"""Generators for the BASELINE.json workloads as `ClusteredSDP`s.

They restate the problem definitions of the reference's examples so that the
oracle and the device library can be fed identical inputs without Julia:

  maxcut          README.md:39-65           (config 2, dense constraint path)
  polyopt         examples/PolyOpt.jl:7-30  (config 1)
  min_f           examples/PolyOpt.jl:32-87 (the reference's documented solver log)
  delsarte        examples/Delsarte.jl:7-49 (config 3)
  sphere_packing  examples/SpherePacking.jl:13-115 (config 5)

and the pieces of src/basesandsamples.jl:28-99,146-169 and
src/approximate_fekete.jl:25-50 they use.  Everything is evaluated with mpmath
at prec+64 bits and rounded to `prec` bits on conversion to wire numbers
(what convert_to_prec does, src/interface.jl:1078-1112).  Cluster / block /
free-variable ORDER is this module's own (the reference orders them by `hash`,
src/interface.jl:885,1032, which only relabels the problem).
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np
import mpmath
from mpmath import mpf

from . import wire
from .sdp import ClusteredSDP, Cluster, PSDBlock, LowRankTerm


# ---------------------------------------------------------------------------
# bases and sample points (src/basesandsamples.jl)
# ---------------------------------------------------------------------------
def chebyshev_values(d, x):
    """T_0..T_d at x  (basis_chebyshev, src/basesandsamples.jl:66-76)."""
    v = [mpf(1)]
    if d >= 1:
        v.append(x)
    for _ in range(2, d + 1):
        v.append(2 * x * v[-1] - v[-2])
    return v


def gegenbauer_values(d, n, x):
    """Gegenbauer polynomials in dimension n, normalised to 1 at 1 (src/basesandsamples.jl:86-96)."""
    v = [mpf(1)]
    if d >= 1:
        v.append(x)
    for l in range(2, d + 1):
        v.append(mpf(2 * l + n - 4) / (l + n - 3) * x * v[-1] - mpf(l - 1) / (l + n - 3) * v[-2])
    return v


def laguerre_values(d, alpha, x):
    """Generalised Laguerre L_0..L_d (src/basesandsamples.jl:28-38)."""
    v = [mpf(1)]
    if d >= 1:
        v.append(1 + alpha - x)
    for l in range(2, d + 1):
        v.append(((2 * l - 1 + alpha - x) * v[-1] - (l + alpha - 1) * v[-2]) / l)
    return v


def laguerre_coefficients(d, alpha, scale):
    """Coefficient lists (low degree first) of L_k^alpha(scale*x), k=0..d."""
    polys = [[mpf(1)]]
    if d >= 1:
        polys.append([1 + alpha, -scale])
    for l in range(2, d + 1):
        a, b = polys[-1], polys[-2]
        new = [mpf(0)] * (l + 1)
        for i, c in enumerate(a):
            new[i] += (2 * l - 1 + alpha) * c
            new[i + 1] -= scale * c
        for i, c in enumerate(b):
            new[i] -= (l + alpha - 1) * c
        polys.append([c / l for c in new])
    return polys


def sample_points_chebyshev(d, a=-1, b=1):
    """d+1 Chebyshev points in [a,b] (src/basesandsamples.jl:162-169)."""
    a, b = mpf(a), mpf(b)
    return [(a + b) / 2 + (b - a) / 2 * mpmath.cospi(mpf(2 * k - 1) / (2 * (d + 1))) for k in range(1, d + 2)]


def sample_points_rescaled_laguerre(d):
    """(src/basesandsamples.jl:146-155)"""
    const = -mpmath.sqrt(mpmath.pi) / (64 * (d + 1) * mpmath.log(3 - 2 * mpmath.sqrt(2)))
    return [const * (-1 + 4 * k) ** 2 for k in range(d + 1)]


def approximatefekete(V, samples, s=3):
    """Orthogonalise the basis w.r.t. the samples and keep a unisolvent subset of them
    (src/approximate_fekete.jl:6-22,51-80, the default :Arb variant).

    V[i][k] = basis_k(sample_i), with at least as many samples as basis functions.  s rounds of
    "QR in Float64, basis change in high precision" on all samples, a column-pivoted QR of V^T to pick
    as many samples as there are basis functions, one more basis change on the chosen rows; rows are
    returned sorted by sample.  Returns (V', samples').
    """
    V = mpmath.matrix(V)
    n = V.cols

    def basis_change(V):
        F = np.array([[float(V[i, j]) for j in range(n)] for i in range(V.rows)])
        R = np.linalg.qr(F, mode="r")
        U = np.linalg.solve(R, np.eye(n))
        Um = mpmath.matrix(n, n)
        for i in range(n):
            for j in range(i, n):
                Um[i, j] = mpf(float(U[i, j]))
        return V * Um

    for _ in range(s):
        V = basis_change(V)
    if V.rows > n:                                  # F2 = qr(Float64.(V'), ColumnNorm()); point_indices = F2.p[1:n]
        import scipy.linalg
        F = np.array([[float(V[i, j]) for j in range(n)] for i in range(V.rows)])
        piv = scipy.linalg.qr(F.T, mode="r", pivoting=True)[1][:n]
        Vs = mpmath.matrix(n, n)
        for ii, i in enumerate(piv):
            for j in range(n):
                Vs[ii, j] = V[int(i), j]
        V, samples = Vs, [samples[int(i)] for i in piv]
    V = basis_change(V)
    order = sorted(range(len(samples)), key=lambda i: samples[i])
    Vs = mpmath.matrix(V.rows, n)
    for ii, i in enumerate(order):
        for j in range(n):
            Vs[ii, j] = V[i, j]
    return Vs, [samples[i] for i in order]


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def _w(values, prec):
    return wire.to_wire(values, prec)


def _rank1(r, s, p, lam, vec, prec):
    return LowRankTerm(r, s, p, _w([lam], prec), _w([list(vec)], prec), _w([list(vec)], prec))


class _Lazy:
    """Dense constraint matrix produced on demand (keeps host memory bounded for n=300)."""

    def __init__(self, fn):
        self.fn = fn

    def __array__(self, dtype=None, copy=None):
        return self.fn()


# ---------------------------------------------------------------------------
# config 2: Goemans-Williamson MAX-CUT relaxation (README.md:39-65)
# ---------------------------------------------------------------------------
def laplacian_random(n, p=0.5, seed=0):
    rng = np.random.default_rng(seed)
    A = np.triu((rng.random((n, n)) < p).astype(np.int64), 1)
    A = A + A.T
    return np.diag(A.sum(axis=1)) - A


def laplacian_complete(n):
    return n * np.eye(n, dtype=np.int64) - np.ones((n, n), dtype=np.int64)


def laplacian_cycle(n):
    L = 2 * np.eye(n, dtype=np.int64)
    for i in range(n):
        L[i, (i + 1) % n] -= 1
        L[(i + 1) % n, i] -= 1
    return L


def maxcut(L, prec=256):
    """maximize <L/4, X> s.t. X_ii = 1, X PSD; every E_ii is passed as a dense matrix (dense path)."""
    L = np.asarray(L)
    n = L.shape[0]
    with mpmath.workprec(prec + 64):
        Cw = wire.wire_zeros((n, n), prec)
        cache = {}
        for i in range(n):
            for j in range(n):
                v = int(L[i, j])
                if v == 0:
                    continue
                if v not in cache:
                    cache[v] = _w(mpf(v) / 4, prec)[()]
                Cw[i, j] = cache[v]
        one = _w(1, prec)[()]
        blk = PSDBlock(m=1, delta=n, high_rank=True, C=Cw, name="X")

        def make(i):
            def fn():
                A = wire.wire_zeros((n, n), prec)
                A[i, i] = one
                return A
            return fn

        for i in range(n):
            blk.dense[i] = _Lazy(make(i))
        c = wire.wire_zeros((n,), prec)
        c[:] = one
        cl = Cluster(B=wire.wire_zeros((n, 0), prec), c=c, blocks=[blk])
        return ClusteredSDP(prec=prec, maximize=True, constant=_w(0, prec), b=wire.wire_zeros((0,), prec),
                            clusters=[cl], name=f"maxcut(n={n})")


def lovasz_theta_cycle(n, prec=256):
    """Lovasz theta of the n-cycle: maximize <J, X> s.t. tr X = 1, X_ij = 0 on the edges, X PSD (the problem behind
    example_theta_problem, test/moi_tests.jl:7-8; theta(C_5) = sqrt 5).  General dense constraint matrices (identity and
    off-diagonal pairs), one cluster, no free variables."""
    with mpmath.workprec(prec + 64):
        one = _w(1, prec)[()]
        half = _w(mpf(1) / 2, prec)[()]
        Cw = wire.wire_zeros((n, n), prec)
        Cw[:, :] = one
        blk = PSDBlock(m=1, delta=n, high_rank=True, C=Cw, name="X")
        A0 = wire.wire_zeros((n, n), prec)
        for i in range(n):
            A0[i, i] = one
        blk.dense[0] = A0
        for e in range(n):
            i, j = e, (e + 1) % n
            A = wire.wire_zeros((n, n), prec)
            A[i, j] = half
            A[j, i] = half
            blk.dense[1 + e] = A
        c = wire.wire_zeros((n + 1,), prec)
        c[0] = one
        cl = Cluster(B=wire.wire_zeros((n + 1, 0), prec), c=c, blocks=[blk])
        return ClusteredSDP(prec=prec, maximize=True, constant=_w(0, prec), b=wire.wire_zeros((0,), prec),
                            clusters=[cl], name=f"theta(C_{n})")


def povm_two_states(prec=256):
    """Optimal discrimination of the two pure qubit states of example_POVM (examples/jump.jl:41-58, test/moi_tests.jl:9-10):
    maximize (1/2) sum_i Re <rho_i, E_i> s.t. E_1 + E_2 = I, E_i Hermitian PSD; optimum 1/2 + sqrt(2)/4.
    JuMP's HermitianPSDCone reaches the solver as real symmetric PSD blocks Z_i = [[A_i, -B_i], [B_i, A_i]] (E_i = A_i + i B_i)
    with equality constraints for the block structure, i.e. a dense-path SDP with TWO blocks in ONE cluster that share the
    coupling constraints: 16 constraints, two 4 x 4 dense blocks, no free variables."""
    with mpmath.workprec(prec + 64):
        def sym(entries):                       # symmetric 4 x 4 matrix with the given (a, b, value) in both triangles
            M = [[mpf(0)] * 4 for _ in range(4)]
            for a, b, v in entries:
                M[a][b] += mpf(v) / (1 if a == b else 2)
                if a != b:
                    M[b][a] += mpf(v) / 2
            return M
        # <sym([(a,b,v)]), Z> = v Z[a][b] for a symmetric Z
        structure = [[(0, 0, 1), (2, 2, -1)], [(1, 1, 1), (3, 3, -1)], [(0, 1, 1), (2, 3, -1)],      # upper-left block = lower-right block
                     [(2, 0, 1)], [(3, 1, 1)], [(3, 0, 1), (2, 1, 1)]]                             # lower-left block antisymmetric
        coupling = [([(0, 0, 1)], 1), ([(1, 1, 1)], 1), ([(0, 1, 1)], 0), ([(3, 0, 1)], 0)]          # sum_i A_i = I, sum_i B_i = 0
        # rho_1 = 1/2 [1,-1][1,-1]^T ; rho_2 = 1/2 [1,-i][1,-i]^* = 1/2 [[1, i], [-i, 1]]: R + iS with S = 1/2 [[0, 1], [-1, 0]]
        R = [[[mpf(1) / 2, -mpf(1) / 2], [-mpf(1) / 2, mpf(1) / 2]], [[mpf(1) / 2, 0], [0, mpf(1) / 2]]]
        S = [[[0, 0], [0, 0]], [[0, mpf(1) / 2], [-mpf(1) / 2, 0]]]
        blocks = []
        P = 2 * len(structure) + len(coupling)
        for i in range(2):
            Zrho = [[mpf(0)] * 4 for _ in range(4)]
            for a in range(2):
                for b in range(2):
                    Zrho[a][b] = Zrho[2 + a][2 + b] = mpf(R[i][a][b])
                    Zrho[2 + a][b] = mpf(S[i][a][b])
                    Zrho[a][2 + b] = -mpf(S[i][a][b])
            C = [[v / 4 for v in row] for row in Zrho]             # (1/N) (1/2) <Z(rho_i), Z_i>, N = 2
            blk = PSDBlock(m=1, delta=4, high_rank=True, C=_w(C, prec), name=("E", i + 1))
            for k, ent in enumerate(structure):
                blk.dense[len(structure) * i + k] = _w(sym(ent), prec)
            for k, (ent, _) in enumerate(coupling):
                blk.dense[2 * len(structure) + k] = _w(sym(ent), prec)
            blocks.append(blk)
        c = [mpf(0)] * (2 * len(structure)) + [mpf(rhs) for _, rhs in coupling]
        cl = Cluster(B=wire.wire_zeros((P, 0), prec), c=_w(c, prec), blocks=blocks)
        return ClusteredSDP(prec=prec, maximize=True, constant=_w(0, prec), b=wire.wire_zeros((0,), prec), clusters=[cl], name="povm_two_states")


# ---------------------------------------------------------------------------
# config 1: univariate polynomial minimisation (examples/PolyOpt.jl:7-30)
# ---------------------------------------------------------------------------
def polyopt(f, d, prec=256, name=None):
    """maximize lambda s.t. f - lambda is a sum of squares of degree 2d.

    `f` is a callable evaluating the polynomial (degree <= 2d) at an mpf.
    Chebyshev basis T_0..T_d, 2d+1 Chebyshev sample points on [-1,1].
    """
    with mpmath.workprec(prec + 64):
        samples = sample_points_chebyshev(2 * d, -1, 1)
        P = len(samples)
        blk = PSDBlock(m=1, delta=d + 1, high_rank=False, C=wire.wire_zeros((d + 1, d + 1), prec), name=("sos", 1))
        for p, x in enumerate(samples):
            blk.lowrank.append(_rank1(0, 0, p, 1, chebyshev_values(d, x), prec))
        B = _w([[1]] * P, prec)
        c = _w([f(x) for x in samples], prec)
        cl = Cluster(B=B, c=c, blocks=[blk])
        return ClusteredSDP(prec=prec, maximize=True, constant=_w(0, prec), b=_w([1], prec), clusters=[cl],
                            name=name or f"polyopt(d={d})")


def polyopt_random(d=20, seed=0, prec=256):
    """Config 1: degree-2d polynomial with N(0,1) Chebyshev coefficients (seed) plus 2*T_2d."""
    rng = np.random.default_rng(seed)
    coef = [float(v) for v in rng.standard_normal(2 * d + 1)]
    coef[2 * d] += 2.0

    def f(x):
        T = chebyshev_values(2 * d, x)
        return mpmath.fsum(mpf(cf) * t for cf, t in zip(coef, T))

    return polyopt(f, d, prec, name=f"polyopt_random(d={d},seed={seed})")


def min_f(d=2, prec=256):
    """The S_3-invariant trivariate example of the reference, `min_f(d)` (examples/PolyOpt.jl:32-87):
    maximize M s.t. x^4 + y^4 + z^4 - 4xyz + x + y + z - M is a sum of squares written in the invariant ring
    (one block per irreducible representation with low degree enough equivariants).  The reference prints its
    solver log for d = 2 in docs/src/solving.md:38-52 (56 iterations, objective -2.1129138814236...), which
    tests/test_oracle_pins.py uses to pin the oracle's TRAJECTORY, not only its optimum.
    d = 2: one cluster, P = 11 samples, blocks of size 4 (rank 1) and 3 (rank 2), one free variable."""
    with mpmath.workprec(prec + 64):
        # invariant basis e1^(deg-2i-3j) e2^i e3^j, deg = 0..2d  (examples/PolyOpt.jl:33-36)
        expo = [(deg - 2 * i - 3 * j, i, j, deg) for deg in range(2 * d + 1) for j in range(deg // 3 + 1) for i in range((deg - 3 * j) // 2 + 1)]
        degrees = [e[3] for e in expo]
        cheb = [sample_points_chebyshev(2 * d + k) for k in range(3)]
        grid = [(cheb[0][i], cheb[1][j], cheb[2][k]) for i in range(2 * d + 1) for j in range(2 * d + 2) for k in range(2 * d + 3)]

        def inv_basis(pt):
            x, y, z = pt
            e1, e2, e3 = x + y + z, x * y + y * z + z * x, x * y * z
            return [e1 ** a * e2 ** b * e3 ** c for (a, b, c, _) in expo]
        V, samples = approximatefekete([inv_basis(pt) for pt in grid], grid)
        P = len(samples)
        # equivariants per irreducible representation of S_3, with their degrees (examples/PolyOpt.jl:61-63)
        equivariants = [[[(lambda x, y, z: mpf(1), 0)]],
                        [[(lambda x, y, z: (x - y) * (y - z) * (z - x), 3)]],
                        [[(lambda x, y, z: 2 * x - y - z, 1), (lambda x, y, z: 2 * y * z - x * z - x * y, 2)],
                         [(lambda x, y, z: y - z, 1), (lambda x, y, z: x * z - x * y, 2)]]]
        factors = [[1], [1], [Fraction(1, 2), Fraction(3, 2)]]
        blocks = []
        for eqi, rows in enumerate(equivariants):
            sel = [[(eq, k) for (eq, edeg) in row for k, qdeg in enumerate(degrees) if 2 * edeg + 2 * qdeg <= 2 * d] for row in rows]
            sel = [v for v in sel if v]
            if not sel:
                continue
            n = len(sel[0])
            blk = PSDBlock(m=1, delta=n, high_rank=False, C=wire.wire_zeros((n, n), prec), name=("trivariatesos", eqi + 1))
            lam = [mpf(f.numerator) / f.denominator if isinstance(f, Fraction) else mpf(f) for f in factors[eqi]]
            for p, pt in enumerate(samples):
                vecs = [[eq(*pt) * V[p, k] for (eq, k) in row] for row in sel]
                blk.lowrank.append(LowRankTerm(0, 0, p, _w(lam[:len(vecs)], prec), _w(vecs, prec), _w(vecs, prec)))
            blocks.append(blk)
        f = lambda x, y, z: x ** 4 + y ** 4 + z ** 4 - 4 * x * y * z + x + y + z
        cl = Cluster(B=_w([[1]] * P, prec), c=_w([f(*pt) for pt in samples], prec), blocks=blocks)
        return ClusteredSDP(prec=prec, maximize=True, constant=_w(0, prec), b=_w([1], prec), clusters=[cl], name=f"min_f(d={d})")


# ---------------------------------------------------------------------------
# config 3: Delsarte LP bound (examples/Delsarte.jl:7-49)
# ---------------------------------------------------------------------------
def delsarte(n, d, costheta, prec=256):
    with mpmath.workprec(prec + 64):
        ct = mpf(costheta.numerator) / costheta.denominator if isinstance(costheta, Fraction) else mpf(costheta)
        samples = sample_points_chebyshev(2 * d, -1, ct)
        V = [chebyshev_values(2 * d, x) for x in samples]
        V, samples = approximatefekete(V, samples)
        ns = len(samples)                 # 2d+1
        P = ns + 1
        blocks = []
        # (:a, k), k = 1..2d : dense 1x1 blocks, coefficient gp[k](x) in constraint 1 and 1 in constraint 2
        gp = [gegenbauer_values(2 * d, n, x) for x in samples]
        for k in range(1, 2 * d + 1):
            blk = PSDBlock(m=1, delta=1, high_rank=True, C=wire.wire_zeros((1, 1), prec), name=("a", k))
            for p in range(ns):
                blk.dense[p] = _w([[gp[p][k]]], prec)
            blk.dense[ns] = _w([[1]], prec)
            blocks.append(blk)
        sos1 = PSDBlock(m=1, delta=d + 1, high_rank=False, C=wire.wire_zeros((d + 1, d + 1), prec), name=("SOS", 1))
        sos2 = PSDBlock(m=1, delta=d, high_rank=False, C=wire.wire_zeros((d, d), prec), name=("SOS", 2))
        for p, x in enumerate(samples):
            sos1.lowrank.append(_rank1(0, 0, p, 1, [V[p, k] for k in range(d + 1)], prec))
            sos2.lowrank.append(_rank1(0, 0, p, (1 + x) * (ct - x), [V[p, k] for k in range(d)], prec))
        blocks += [sos1, sos2]
        slack = PSDBlock(m=1, delta=1, high_rank=True, C=wire.wire_zeros((1, 1), prec), name="slack")
        slack.dense[ns] = _w([[1]], prec)
        blocks.append(slack)
        B = _w([[0]] * ns + [[-1]], prec)
        c = _w([-1] * P, prec)
        cl = Cluster(B=B, c=c, blocks=blocks)
        return ClusteredSDP(prec=prec, maximize=False, constant=_w(0, prec), b=_w([1], prec), clusters=[cl],
                            name=f"delsarte(n={n},d={d})")


# ---------------------------------------------------------------------------
# config 5: N-radii sphere packing (examples/SpherePacking.jl:13-115)
# ---------------------------------------------------------------------------
def sphere_packing(n, d, r, prec=256):
    with mpmath.workprec(prec + 64):
        Nr = len(r)
        r = [mpf(v.numerator) / v.denominator if isinstance(v, Fraction) else mpf(v) for v in r]
        pairs = [(i, j) for i in range(Nr) for j in range(i + 1)]        # i >= j
        T = len(pairs)
        deg = 2 * d + 1
        ns = deg + 1                                                     # samples per polynomial constraint
        alpha = mpf(n) / 2 - 1
        # free variables: (k,i,j) for k=0..2d+1, then M
        fidx = {}
        for (i, j) in pairs:
            for k in range(deg + 1):
                fidx[(k, i, j)] = len(fidx)
        fidx["M"] = len(fidx)
        N = len(fidx)

        def vol(rad):
            return mpmath.sqrt(mpmath.pi) ** n / mpmath.gamma(mpf(n) / 2 + 1) * rad ** n

        # orthogonalised Laguerre basis on the rescaled Laguerre points (:50-54)
        samples = sample_points_rescaled_laguerre(deg)
        polys = laguerre_coefficients(deg, alpha, 2 * mpmath.pi)
        maxc = [max(pl) for pl in polys]
        V = [[mpmath.polyval(list(reversed(polys[k])), x) / maxc[k] for k in range(deg + 1)] for x in samples]
        V, samples = approximatefekete(V, samples)
        basis = lambda p, cnt: [V[p, k] for k in range(cnt)]
        zeroB = lambda rows: [[mpf(0)] * N for _ in range(rows)]
        clusters = []

        # constraint 1 (:33-48): cluster PSD1, one block with Nr x Nr subblocks of size 1
        blk = PSDBlock(m=Nr, delta=1, high_rank=False, C=wire.wire_zeros((Nr, Nr), prec), name="PSD1")
        Bm, cv = zeroB(T), []
        for p, (i, j) in enumerate(pairs):
            cv.append(-mpmath.sqrt(vol(r[i]) * vol(r[j])))
            Bm[p][fidx[(0, i, j)]] = mpf(-1)
            if i != j:
                blk.lowrank.append(_rank1(i, j, p, mpf(1) / 2, [1], prec))
                blk.lowrank.append(_rank1(j, i, p, mpf(1) / 2, [1], prec))
            else:
                blk.lowrank.append(_rank1(i, i, p, 1, [1], prec))
        clusters.append(Cluster(B=_w(Bm, prec), c=_w(cv, prec), blocks=[blk]))

        # constraint 2 (:56-81): cluster SOS2, blocks SOS21 and SOS22 with Nr x Nr subblocks of size d+1
        b21 = PSDBlock(m=Nr, delta=d + 1, high_rank=False, C=wire.wire_zeros((Nr * (d + 1),) * 2, prec), name="SOS21")
        b22 = PSDBlock(m=Nr, delta=d + 1, high_rank=False, C=wire.wire_zeros((Nr * (d + 1),) * 2, prec), name="SOS22")
        Bm = zeroB(T * ns)
        row = 0
        for (i, j) in pairs:
            for p, x in enumerate(samples):
                vec = basis(p, d + 1)
                for k in range(deg + 1):
                    Bm[row][fidx[(k, i, j)]] = (-2 if i != j else -1) * x ** k
                for (a, b_) in ([(i, j), (j, i)] if i != j else [(i, i)]):
                    b21.lowrank.append(_rank1(a, b_, row, 1, vec, prec))
                    b22.lowrank.append(_rank1(a, b_, row, x, vec, prec))
                row += 1
        clusters.append(Cluster(B=_w(Bm, prec), c=_w([0] * (T * ns), prec), blocks=[b21, b22]))

        # constraint 3 (:83-95): one cluster per pair, blocks SOS31 (1x1) and SOS32 (d+1)
        fact = [mpmath.factorial(k) / mpmath.pi ** k for k in range(deg + 1)]
        for (i, j) in pairs:
            b31 = PSDBlock(m=1, delta=1, high_rank=False, C=wire.wire_zeros((1, 1), prec), name=("SOS31", i, j))
            b32 = PSDBlock(m=1, delta=d + 1, high_rank=False, C=wire.wire_zeros((d + 1, d + 1), prec), name=("SOS32", i, j))
            Bm = zeroB(ns)
            for p, x in enumerate(samples):
                Lv = laguerre_values(deg, alpha, mpmath.pi * x)
                for k in range(deg + 1):
                    Bm[p][fidx[(k, i, j)]] = fact[k] * Lv[k]
                b31.lowrank.append(_rank1(0, 0, p, 1, basis(p, 1), prec))
                b32.lowrank.append(_rank1(0, 0, p, x - (r[i] + r[j]) ** 2, basis(p, d + 1), prec))
            clusters.append(Cluster(B=_w(Bm, prec), c=_w([0] * ns, prec), blocks=[b31, b32]))

        # constraint 4 (:97-107): one cluster per radius, a dense 1x1 slack block
        L0 = laguerre_values(deg, alpha, mpf(0))
        for i in range(Nr):
            sl = PSDBlock(m=1, delta=1, high_rank=True, C=wire.wire_zeros((1, 1), prec), name=("slack4", i))
            sl.dense[0] = _w([[1]], prec)
            Bm = zeroB(1)
            for k in range(deg + 1):
                Bm[0][fidx[(k, i, i)]] = fact[k] * L0[k]
            Bm[0][fidx["M"]] = mpf(-1)
            clusters.append(Cluster(B=_w(Bm, prec), c=_w([0], prec), blocks=[sl]))

        bvec = [mpf(0)] * N
        bvec[fidx["M"]] = mpf(1)
        return ClusteredSDP(prec=prec, maximize=False, constant=_w(0, prec), b=_w(bvec, prec), clusters=clusters,
                            name=f"sphere_packing(n={n},d={d},Nr={Nr})")


def cohnelkies(n, d, r=1, prec=256):
    """The Cohn-Elkies linear-programming bound for sphere packings in R^n as an SDP (examples/SpherePacking.jl:117-185;
    test/runtests_solver.jl:19-20: cohnelkies(8, 15, prec=256) = pi^4/384 to 1e-4).

    Free variables a_1..a_{2d+1} (a_0 = 1 is the constant of the first constraint):
      con1 at the rescaled Laguerre points x:   <SOS21, b b^T> + x <SOS22, b b^T> - sum_k a_k x^k = 1
      con2 at the points x + r^2:               b_0(x)^2 SOS31 + (x - r^2) <SOS32, b b^T> + sum_k a_k k!/pi^k L_k(pi x) = -L_0(pi x) = -1
      minimise   vol(B(r/2)) (L_0(0) + sum_k a_k k!/pi^k L_k(0))
    Two clusters coupled by the free variables; SOS31 is given as a 1 x 1 matrix (the dense path), the others as rank-1 terms;
    the basis is re-orthogonalised on the shifted samples for the second constraint, as the example does."""
    with mpmath.workprec(prec + 64):
        r = mpf(r.numerator) / r.denominator if isinstance(r, Fraction) else mpf(r)
        deg = 2 * d + 1
        ns = deg + 1
        alpha = mpf(n) / 2 - 1
        N = deg
        vol = mpmath.sqrt(mpmath.pi) ** n / mpmath.gamma(mpf(n) / 2 + 1) * (r / 2) ** n
        polys = laguerre_coefficients(deg, alpha, 2 * mpmath.pi)
        maxc = [max(pl) for pl in polys]
        evalb = lambda xs: [[mpmath.polyval(list(reversed(polys[k])), x) / maxc[k] for k in range(deg + 1)] for x in xs]
        fact = [mpmath.factorial(k) / mpmath.pi ** k for k in range(deg + 1)]
        # constraint 1
        x1 = sample_points_rescaled_laguerre(deg)
        V1, x1 = approximatefekete(evalb(x1), x1)
        b21 = PSDBlock(m=1, delta=d + 1, high_rank=False, C=wire.wire_zeros((d + 1, d + 1), prec), name="SOS21")
        b22 = PSDBlock(m=1, delta=d + 1, high_rank=False, C=wire.wire_zeros((d + 1, d + 1), prec), name="SOS22")
        B1 = []
        for p, x in enumerate(x1):
            vec = [V1[p, k] for k in range(d + 1)]
            b21.lowrank.append(_rank1(0, 0, p, 1, vec, prec))
            b22.lowrank.append(_rank1(0, 0, p, x, vec, prec))
            B1.append([-x ** k for k in range(1, deg + 1)])
        c1 = Cluster(B=_w(B1, prec), c=_w([1] * ns, prec), blocks=[b21, b22])
        # constraint 2
        x2 = [x + r ** 2 for x in sample_points_rescaled_laguerre(deg)]
        V2, x2 = approximatefekete(evalb(x2), x2)
        b31 = PSDBlock(m=1, delta=1, high_rank=True, C=wire.wire_zeros((1, 1), prec), name="SOS31")
        b32 = PSDBlock(m=1, delta=d + 1, high_rank=False, C=wire.wire_zeros((d + 1, d + 1), prec), name="SOS32")
        B2 = []
        for p, x in enumerate(x2):
            b31.dense[p] = _w([[V2[p, 0] ** 2]], prec)
            b32.lowrank.append(_rank1(0, 0, p, x - r ** 2, [V2[p, k] for k in range(d + 1)], prec))
            Lv = laguerre_values(deg, alpha, mpmath.pi * x)
            B2.append([fact[k] * Lv[k] for k in range(1, deg + 1)])
        c2 = Cluster(B=_w(B2, prec), c=_w([-1] * ns, prec), blocks=[b31, b32])
        L0 = laguerre_values(deg, alpha, mpf(0))
        bvec = [vol * fact[k] * L0[k] for k in range(1, deg + 1)]
        return ClusteredSDP(prec=prec, maximize=False, constant=_w(vol * L0[0], prec), b=_w(bvec, prec), clusters=[c1, c2],
                            name=f"cohnelkies(n={n},d={d})")


# ---------------------------------------------------------------------------
# config 4: three-point bound for spherical codes (examples/ThreePointBound.jl:45-169)
# ---------------------------------------------------------------------------
def gegenbauer_coefficients(d, n):
    """Coefficient lists (low degree first) of the Gegenbauer polynomials of basis_gegenbauer(d, n, x)."""
    polys = [[mpf(1)]]
    if d >= 1:
        polys.append([mpf(0), mpf(1)])
    for l in range(2, d + 1):
        a, b = polys[-1], polys[-2]
        new = [mpf(0)] * (l + 1)
        for i, c in enumerate(a):
            new[i + 1] += mpf(2 * l + n - 4) / (l + n - 3) * c
        for i, c in enumerate(b):
            new[i] -= mpf(l - 1) / (l + n - 3) * c
        polys.append(new)
    return polys


def three_point_bound(n, costheta, d2, d3, prec=256, seed=1935):
    """three_point_spherical_codes(n, costheta, d2, d3).  One cluster (both constraints share the
    (:F,k) blocks): dense F_k blocks, rank-1 univariate SOS / a_k blocks, rank-1 and rank-2
    S_3-invariant trivariate SOS blocks.  The trivariate sample subset is drawn with numpy
    default_rng(seed) (Julia's shuffle stream cannot be reproduced without Julia)."""
    with mpmath.workprec(prec + 64):
        ct = mpf(costheta.numerator) / costheta.denominator if isinstance(costheta, Fraction) else mpf(costheta)
        N2, N3 = max(d2, d3), d3
        geg = gegenbauer_coefficients(max(d3, 1), n - 1)
        floor4 = lambda x: mpf(int(mpmath.floor(10 ** 4 * x))) / 10 ** 4

        def Qf(k, u, v, t):
            cf = geg[k]
            return mpmath.fsum(cf[i] * ((1 - u * u) * (1 - v * v)) ** ((k - i) // 2) * (t - u * v) ** i for i in range(len(cf)) if cf[i] != 0)

        def mvec(w, d):
            return [w ** k for k in range(d + 1)]

        def Smat(k, d, u, v, t):
            mu, mv, mt = mvec(u, d - k), mvec(v, d - k), mvec(t, d - k)
            q1, q2, q3 = Qf(k, u, v, t), Qf(k, t, u, v), Qf(k, t, v, u)
            sz = d - k + 1
            return [[(q1 * (mv[a] * mu[b] + mu[a] * mv[b]) + q2 * (mt[a] * mu[b] + mu[a] * mt[b]) + q3 * (mt[a] * mv[b] + mv[a] * mt[b])) / 6
                     for b in range(sz)] for a in range(sz)]

        pw = lambda u: (u + 1) * (ct - u)
        # ---- samples
        s1 = [floor4(x) for x in sample_points_chebyshev(2 * N2, -1, 1)]
        ntri = len([1 for deg in range(2 * N3 + 1) for k in range(deg // 3 + 1) for j in range((deg - 3 * k) // 2 + 1)])
        cheb = [sample_points_chebyshev(2 * N3 + k, -1, 1) for k in range(3)]
        grid = [(cheb[0][i], cheb[1][j], cheb[2][k]) for i in range(2 * N3 + 1) for j in range(2 * N3 + 2) for k in range(2 * N3 + 3)]
        rng = np.random.default_rng(seed)
        pick = sorted(rng.permutation(len(grid))[:ntri].tolist(), key=lambda i: grid[i])
        s3 = [tuple(floor4(x) for x in grid[i]) for i in pick]
        P1, P = len(s1), len(s1) + len(s3)
        blocks = []
        # ---- (:F, k): dense blocks in both constraints
        for k in range(d3 + 1):
            sz = d3 - k + 1
            Cm = [[mpf(1) if k == 0 else mpf(0)] * sz for _ in range(sz)]
            blk = PSDBlock(m=1, delta=sz, high_rank=True, C=_w(Cm, prec), name=("F", k))
            for p, w in enumerate(s1):
                blk.dense[p] = _w([[3 * v for v in row] for row in Smat(k, d3, w, w, mpf(1))], prec)
            for p, (u, v, t) in enumerate(s3):
                blk.dense[P1 + p] = _w(Smat(k, d3, u, v, t), prec)
            blocks.append(blk)
        # ---- (:a, k): 1x1 rank-one blocks, univariate constraint only
        if d2 >= 0:
            for k in range(2 * d2 + 1):
                blk = PSDBlock(m=1, delta=1, high_rank=False, C=_w([[1]], prec), name=("a", k))
                for p, w in enumerate(s1):
                    blk.lowrank.append(_rank1(0, 0, p, gegenbauer_values(2 * d2, n, w)[k], [1], prec))
                blocks.append(blk)
        # ---- univariate SOS blocks
        if N2 >= 0:
            b1 = PSDBlock(m=1, delta=N2 + 1, high_rank=False, C=wire.wire_zeros((N2 + 1, N2 + 1), prec), name=("univariatesos", 1))
            for p, w in enumerate(s1):
                b1.lowrank.append(_rank1(0, 0, p, 1, chebyshev_values(2 * N2, w)[:N2 + 1], prec))
            blocks.append(b1)
        if N2 >= 1:
            b2 = PSDBlock(m=1, delta=N2, high_rank=False, C=wire.wire_zeros((N2, N2), prec), name=("univariatesos", 2))
            for p, w in enumerate(s1):
                b2.lowrank.append(_rank1(0, 0, p, pw(w), chebyshev_values(2 * N2, w)[:N2], prec))
            blocks.append(b2)
        # ---- trivariate invariant SOS blocks
        basis_idx = [(deg, k, j) for deg in range(N3 + 1) for k in range(deg // 3 + 1) for j in range((deg - 3 * k) // 2 + 1)]
        equivariants = [([lambda u, v, t: mpf(1)], [0]),
                        ([lambda u, v, t: (u - v) * (v - t) * (t - u)], [3]),
                        ([lambda u, v, t: 2 * u - v - t, lambda u, v, t: 2 * v * t - u * t - u * v], [1, 2]),      # eqi = 3, r = 1
                        ([lambda u, v, t: v - t, lambda u, v, t: u * t - u * v], [1, 2])]                          # eqi = 3, r = 2
        eq_groups = [[0], [1], [2, 3]]
        factors = [[mpf(1)], [mpf(1)], [mpf(1) / 2, mpf(3) / 2]]
        weights = [(lambda u, v, t: mpf(1), 0), (lambda u, v, t: pw(u) + pw(v) + pw(t), 2),
                   (lambda u, v, t: pw(u) * pw(v) + pw(v) * pw(t) + pw(t) * pw(u), 4), (lambda u, v, t: pw(u) * pw(v) * pw(t), 6),
                   (lambda u, v, t: 2 * u * v * t + 1 - u * u - v * v - t * t, 3)]
        for wi, (wf, wdeg) in enumerate(weights):
            if wdeg > 2 * N3:
                continue
            for eqi, rows in enumerate(eq_groups):
                # which (eq, basis) pairs enter row r
                sel = [[(e, bi) for e in range(len(equivariants[r][0])) for bi, (deg, _, _) in enumerate(basis_idx)
                        if wdeg + 2 * equivariants[r][1][e] + 2 * deg <= 2 * N3] for r in rows]
                sel = [s for s in sel if s]
                if not sel:
                    continue
                rows_used = [r for r, s in zip(rows, [[(e, bi) for e in range(len(equivariants[r][0])) for bi, (deg, _, _) in enumerate(basis_idx)
                                                      if wdeg + 2 * equivariants[r][1][e] + 2 * deg <= 2 * N3] for r in rows]) if s]
                delta = len(sel[0])
                assert all(len(s) == delta for s in sel)
                blk = PSDBlock(m=1, delta=delta, high_rank=False, C=wire.wire_zeros((delta, delta), prec), name=("trivariatesos", wi, eqi))
                for p, (u, v, t) in enumerate(s3):
                    s1v, s2v, s3v = u + v + t, u * v + v * t + u * t, u * v * t
                    bas = [s1v ** (deg - 3 * k - 2 * j) * s2v ** j * s3v ** k for (deg, k, j) in basis_idx]
                    wv = wf(u, v, t)
                    lam, vecs = [], []
                    for r, s in zip(rows_used, sel):
                        eqv = [f(u, v, t) for f in equivariants[r][0]]
                        lam.append(wv * factors[eqi][rows.index(r)])
                        vecs.append([eqv[e] * bas[bi] for (e, bi) in s])
                    blk.lowrank.append(LowRankTerm(0, 0, P1 + p, _w(lam, prec), _w(vecs, prec), _w(vecs, prec)))
                blocks.append(blk)
        c = _w([-1] * P1 + [0] * len(s3), prec)
        cl = Cluster(B=wire.wire_zeros((P, 0), prec), c=c, blocks=blocks)
        return ClusteredSDP(prec=prec, maximize=False, constant=_w(1, prec), b=wire.wire_zeros((0,), prec), clusters=[cl],
                            name=f"three_point_bound(n={n},d2={d2},d3={d3})")

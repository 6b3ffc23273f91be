"""Column-pivoted QR and the dependent-constraint part of preprocess! (src/pre_postprocessing.jl:4-137,183-213)
on the CPU oracle; the device runs the same cases in tests/test_gpu_preprocess.py."""
import mpmath
import numpy as np
import pytest

from clrs_b200 import ClusteredSDP, Cluster, PSDBlock, Solver, solvesdp, wire, workloads
from clrs_b200 import preprocess as pp

PREC = 256


def w(v, prec=PREC):
    return wire.to_wire(v, prec)


def toy(cvals, Bvals=None, avals=None, maximize=False):
    """min <1, X> + ... with one 1 x 1 PSD variable X and constraints avals[p] X + B[p] . y = cvals[p] (test/runtests_solver.jl:249-314)."""
    P = len(cvals)
    avals = avals or [1] * P
    N = len(Bvals[0]) if Bvals else 0
    blk = PSDBlock(m=1, delta=1, high_rank=True, C=w([[1]]))
    for p in range(P):
        blk.dense[p] = w([[avals[p]]])
    B = w(Bvals) if N else wire.wire_zeros((P, 0), PREC)
    return ClusteredSDP(prec=PREC, maximize=maximize, constant=w(0), b=w([0] * N) if N else wire.wire_zeros((0,), PREC),
                        clusters=[Cluster(B=B, c=w(cvals), blocks=[blk])], name="toy")


@pytest.fixture(scope="module")
def handle():
    s = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="oracle")
    yield s
    s.close()


def qr_case(seed=0, m=14, n=8, spread=5):
    rng = np.random.default_rng(seed)
    A = rng.integers(-spread, spread + 1, size=(m, n - 2)).astype(object)
    A = np.hstack([A, (2 * A[:, [0]] - A[:, [3]]), (A[:, [1]] + A[:, [2]] + A[:, [4]])])      # two exactly dependent columns
    return A


def check_qr(solver, A, rank):
    m, n = A.shape
    R, perm = solver.mp_qr_pivot(w(A.tolist()))
    assert sorted(perm.tolist()) == list(range(n))
    with mpmath.workprec(PREC + 64):
        Rm = wire.from_wire(R, PREC)
        AP = mpmath.matrix(A[:, perm].tolist())
        G = AP.T * AP
        RR = mpmath.matrix(Rm.tolist())
        E = RR.T * RR - G
        scale = max(abs(G[i, i]) for i in range(n))
        assert max(abs(E[i, j]) for i in range(n) for j in range(n)) < scale * mpmath.mpf(2) ** -240      # R^T R = (A P)^T (A P)
        for i in range(min(m, n)):
            assert all(Rm[i, j] == 0 for j in range(i))                                                  # upper triangular
            assert Rm[i, i] >= 0
            if i + 1 < min(m, n):
                assert abs(Rm[i, i]) >= abs(Rm[i + 1, i + 1]) * (1 - mpmath.mpf(2) ** -200)              # pivoting: non-increasing diagonal
        tol = mpmath.sqrt(mpmath.mpf(2) ** (1 - PREC))
        assert sum(1 for i in range(min(m, n)) if abs(Rm[i, i]) >= tol) == rank
    return R, perm


def test_pivoted_qr_reproduces_the_gram_matrix_and_finds_the_rank(handle):
    check_qr(handle, qr_case(), 6)
    check_qr(handle, qr_case(3, 5, 9), 5)                     # more columns than rows


def test_duplicate_constraint_is_removed_and_the_optimum_is_one(handle):
    """`[cs10, cs10]` of test/runtests_solver.jl:309-313: min X s.t. X = 1 twice -> objective 1."""
    sdp = toy([1, 1])
    new, cs, _ = pp.preprocess(sdp, handle)
    assert len(cs) == 1 and new.clusters[0].P == 1
    r = solvesdp(new, lib="oracle")
    assert r.status == "Optimal" and abs(r.p_obj - 1) < mpmath.mpf(10) ** -12


def test_contradicting_duplicates_are_infeasible(handle):
    """`[cs10, cs11]`, test/runtests_solver.jl:305-308: X = 1 and X = 0 -> the reference raises (src/pre_postprocessing.jl:91-99)."""
    with pytest.raises(ValueError):
        pp.preprocess(toy([1, 0]), handle)


def reference_toy(cons, obj_free=(0, 0, 0), cY=0):
    """The toy problems of test/runtests_solver.jl:249-303 in container form: minimise 1 + X + cY Y + obj_free . (x, y, z) subject to
    a_X X + a_Y Y + beta . (x, y, z) = c for cons = [(c, a_X, a_Y, (beta_x, beta_y, beta_z))].  X, Y are 1 x 1 PSD variables; constraints
    that share a PSD variable form a cluster (constraints without one get a cluster without blocks)."""
    groups = {}
    for con in cons:
        key = "X" if con[1] else ("Y" if con[2] else "none")
        groups.setdefault(key, []).append(con)
    clusters = []
    for key, cl in groups.items():
        blocks = []
        if key != "none":
            blk = PSDBlock(m=1, delta=1, high_rank=True, C=w([[1 if key == "X" else cY]]))
            for p, con in enumerate(cl):
                blk.dense[p] = w([[con[1] if key == "X" else con[2]]])
            blocks = [blk]
        clusters.append(Cluster(B=w([list(con[3]) for con in cl]), c=w([con[0] for con in cl]), blocks=blocks))
    return ClusteredSDP(prec=PREC, maximize=False, constant=w(1), b=w(list(obj_free)), clusters=clusters, name="reference toy")


CS1 = (1, 1, 0, (1, 1, 0)); CS2 = (2, 1, 0, (2, 3, 0)); CS3 = (1, 0, 0, (2, 0, 0)); CS4 = (2, 2, 0, (2, 2, 0)); CS5 = (4, 1, 0, (1, 1, 0))
CS6 = (1, 0, 1, (2, 2, 0)); CS7 = (1, 0, 1, (0, 0, 1)); CS8 = (0, 0, 0, (1, 0, 0)); CS9 = (mpmath.mpf(1) / 2, 0, 0, (0, 1, 0))


def solve_reduced(sdp, handle, expect):
    """preprocess, solve the reduced SDP on the oracle, put the removed variables back and check the objective and the constraints of the
    ORIGINAL SDP (objvalue and slacks of the reference's tests, atol 1e-5 there)."""
    new, cs, rel = pp.preprocess(sdp, handle)
    r = solvesdp(new, lib="oracle", keep_solver=True, omega_p=10, omega_d=10)
    x, X, y, Y = r.solver.get_state()
    r.solver.close()
    assert r.status in ("Optimal", "NearOptimal"), r
    with mpmath.workprec(PREC + 64):
        assert abs(r.p_obj - expect) < mpmath.mpf(10) ** -8, (r, expect)
        xf, yf = pp.postprocess(sdp, list(wire.from_wire(x, PREC)), list(wire.from_wire(y, PREC)) if new.N else [], cs, rel)
        assert len(xf) == sdp.num_constraints and len(yf) == sdp.N
        # the free variables of the original problem satisfy the constraints without PSD part exactly
        for cl in sdp.clusters:
            if not cl.blocks:
                B = wire.from_wire(cl.B, PREC).reshape(cl.P, sdp.N); cv = wire.from_wire(cl.c, PREC)
                for p in range(cl.P):
                    assert abs(sum(B[p, k] * yf[k] for k in range(sdp.N)) - cv[p]) < mpmath.mpf(10) ** -8
    return new, cs, rel


def test_reference_toy_problems_with_dependent_constraints_and_free_variables(handle):
    """test/runtests_solver.jl:249-303: known objectives 1, 5/4, 1, 3/2, 5/4, 0 and the two infeasible cases."""
    new, cs, rel = solve_reduced(reference_toy([CS1, CS2]), handle, 1)                       # :250-256   x + 2y = 1 between the free variables
    assert len(cs) == 1 and len(rel["nf_vars"]) == 1 and new.N == 1
    solve_reduced(reference_toy([CS1, CS2, CS3]), handle, mpmath.mpf(5) / 4)                 # :259-263   no PSD variable in cs3
    solve_reduced(reference_toy([CS1, CS2, CS4]), handle, 1)                                 # :266-270   a multiple of cs1: 0 = 0 in the free part
    with pytest.raises(ValueError):
        pp.preprocess(reference_toy([CS1, CS2, CS5]), handle)                                # :273-275   same parts, different constant
    new, cs, rel = solve_reduced(reference_toy([CS1, CS6]), handle, mpmath.mpf(3) / 2)       # :278-282   linearly dependent free variables
    assert cs == [] and len(rel["fv_zeros"]) == 2 and new.N == 1                          # y duplicates x, z does not occur at all
    solve_reduced(reference_toy([CS1, CS2, CS3, CS4, CS7], cY=1), handle, mpmath.mpf(5) / 4)   # :286-290   objective 1 + X + Y
    solve_reduced(reference_toy([CS1, CS2], obj_free=(-1, -2, 0)), handle, 0)                # :293-297   free variables in the objective
    with pytest.raises(ValueError):
        pp.preprocess(reference_toy([CS1, CS2, CS8, CS9, CS5]), handle)                      # :300-303   incompatible constraints on the free variables


def test_redundant_trace_constraint_of_maxcut_is_removed_without_changing_the_optimum(handle):
    """MAX-CUT of C_5 plus the sum of its own constraints (<I, X> = 5): dependent, consistent, removed; the optimum is that of
    the plain SDP (n/4 (2 + 2 cos(pi/n)))."""
    n = 5
    sdp = workloads.maxcut(workloads.laplacian_cycle(n))
    cl = sdp.clusters[0]
    blk = cl.blocks[0]
    blk.dense = {p: np.asarray(A) for p, A in blk.dense.items()}
    blk.dense[n] = wire.wire_eye_scaled(n, 1, PREC)
    sdp.clusters[0] = Cluster(B=wire.wire_zeros((n + 1, 0), PREC), c=w([1] * n + [n]), blocks=[blk])
    new, cs, _ = pp.preprocess(sdp, handle)
    assert len(cs) == 1 and new.clusters[0].P == n
    assert pp.preprocess(workloads.maxcut(workloads.laplacian_cycle(n)), handle)[1] == []      # nothing to remove in the plain SDP
    r = solvesdp(new, lib="oracle", duality_gap_threshold=1e-30)
    with mpmath.workprec(300):
        assert abs(r.p_obj - mpmath.mpf(n) / 4 * (2 + 2 * mpmath.cos(mpmath.pi / n))) < mpmath.mpf(10) ** -28


def test_low_rank_constraints_are_vectorised_like_the_reference(handle):
    """Delsarte's constraints are independent: nothing is removed; the vectorisation has one row per packed entry of every block."""
    from fractions import Fraction
    sdp = workloads.delsarte(8, 3, Fraction(1, 2))
    M, cols = pp.vectorize_constraints(sdp)
    assert M.shape == (sum(b.m * (b.delta * (b.delta + 1) // 2) + (b.m * (b.m - 1) // 2) * b.delta ** 2 for c in sdp.clusters for b in c.blocks), sdp.num_constraints)
    assert pp.find_dependent_constraints(sdp, handle) == []

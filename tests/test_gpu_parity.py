"""Parity of the CUDA path against the CPU oracle (and, for the integer slice
pipeline, bit-exactly against the host build of the same arithmetic).  All calls
go through the C ABI (ctypes)."""
import ctypes as C
import os
import random
from fractions import Fraction

import mpmath
import numpy as np
import pytest

import clrs_b200
from clrs_b200 import wire, workloads, Solver, solvesdp

pytestmark = pytest.mark.gpu
PREC = 256
TOL_OBJ = mpmath.mpf(10) ** -25      # north star: objectives agree to a relative 1e-25 at 256 bit


def rnd_matrix(rng, r, c, spread, zero_frac=0.05):
    out = []
    for _ in range(r):
        row = []
        for _ in range(c):
            if rng.random() < zero_frac:
                row.append(mpmath.mpf(0)); continue
            m = mpmath.mpf(rng.getrandbits(300)) / 2 ** 300 + mpmath.mpf(1) / 7
            row.append(rng.choice([1, -1]) * m * mpmath.mpf(2) ** rng.randint(-spread, spread))
        out.append(row)
    return out


def rand_wire(seed, shape, spread, zero_frac=0.02):
    """Random wire matrix built directly in the wire format (fast, numpy)."""
    rng = np.random.default_rng(seed)
    out = wire.wire_zeros(shape, PREC)
    limbs = rng.integers(0, 2 ** 63, size=shape + (4,), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape + (4,), dtype=np.uint64)
    limbs[..., 3] |= np.uint64(1) << np.uint64(63)
    out["limb"] = limbs
    out["exp"] = rng.integers(-spread, spread + 1, size=shape)
    out["sign"] = np.where(rng.random(shape) < zero_frac, 0, rng.choice([-1, 1], size=shape))
    return out


@pytest.fixture(scope="module")
def tiny():
    sdp = workloads.maxcut(workloads.laplacian_cycle(3))
    s = Solver(sdp, lib="device")
    yield s
    s.close()


@pytest.fixture(scope="module")
def tiny_oracle():
    sdp = workloads.maxcut(workloads.laplacian_cycle(3))
    s = Solver(sdp, lib="oracle")
    yield s
    s.close()


def hostcheck():
    return C.CDLL(os.path.join(os.path.dirname(clrs_b200.DEVICE_LIB), "libclrs_hostcheck.so"))


@pytest.mark.parametrize("M,N,K,spread", [(1, 1, 1, 0), (3, 4, 5, 0), (17, 19, 33, 40), (40, 33, 70, 300), (5, 5, 8, 2), (64, 48, 130, 10)])
def test_gemm_bitexact_vs_host_arithmetic(tiny, M, N, K, spread):
    """The int8 slice pipeline is exact integer arithmetic: the device result must equal
    the host build of the same split/recombine code bit for bit."""
    rng = random.Random(M * 1000 + N * 10 + K)
    with mpmath.workprec(400):
        A = wire.to_wire(rnd_matrix(rng, M, K, spread), PREC)
        B = wire.to_wire(rnd_matrix(rng, K, N, spread), PREC)
    Cd, _ = tiny.mp_gemm(A, B)
    Ch = wire.wire_zeros((M, N), PREC)
    hostcheck().hc_gemm(M, N, K, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), Ch.ctypes.data_as(C.c_void_p))
    assert Cd.tobytes() == Ch.tobytes()


def test_gemm_vs_oracle(tiny, tiny_oracle):
    rng = random.Random(7)
    M, N, K = 20, 24, 50
    with mpmath.workprec(400):
        A = wire.to_wire(rnd_matrix(rng, M, K, 30), PREC)
        B = wire.to_wire(rnd_matrix(rng, K, N, 30), PREC)
    Cd, _ = tiny.mp_gemm(A, B)
    Co, _ = tiny_oracle.mp_gemm(A, B)
    with mpmath.workprec(600):
        a, b, cd, co = (wire.from_wire(v, PREC) for v in (A, B, Cd, Co))
        for i in range(M):
            for j in range(N):
                scale = max(abs(v) for v in a[i, :]) * max(abs(v) for v in b[:, j])
                assert abs(cd[i, j] - co[i, j]) <= scale * K * mpmath.mpf(2) ** -250


@pytest.mark.parametrize("n", [1, 2, 7, 32, 33, 70, 100])
def test_cholesky_vs_oracle(tiny, tiny_oracle, n):
    rng = random.Random(n)
    with mpmath.workprec(400):
        G = mpmath.matrix(rnd_matrix(rng, n, n, 3, 0.0))
        A = G * G.T + mpmath.eye(n) * n
        Aw = wire.to_wire(A, PREC)
    Ld = tiny.mp_cholesky(Aw)
    Lo = tiny_oracle.mp_cholesky(Aw)
    with mpmath.workprec(600):
        ld, lo = wire.from_wire(Ld, PREC), wire.from_wire(Lo, PREC)
        scale = max(abs(v) for v in lo.reshape(-1))
        err = max(abs(x - y) for x, y in zip(ld.reshape(-1), lo.reshape(-1)))
        assert err <= scale * mpmath.mpf(2) ** -235     # kappa-amplified rounding only
        for i in range(n):
            for j in range(i + 1, n):
                assert ld[i, j] == 0


def test_cholesky_reports_nonpositive_pivot(tiny):
    A = wire.to_wire([[1, 2], [2, 1]], PREC)
    with pytest.raises(clrs_b200.SolverFailure):
        tiny.mp_cholesky(A)


def _compare(sdp, **kw):
    dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30, **kw)
    ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30, **kw)
    assert dev.status == ref.status == "Optimal", (dev, ref)
    with mpmath.workprec(400):
        assert abs(dev.p_obj - ref.p_obj) <= TOL_OBJ * max(1, abs(ref.p_obj)), (dev, ref)
        assert abs(dev.d_obj - ref.d_obj) <= TOL_OBJ * max(1, abs(ref.d_obj)), (dev, ref)
        assert abs(dev.gap - ref.gap) <= TOL_OBJ
    assert abs(dev.iterations - ref.iterations) <= 1, (dev.iterations, ref.iterations)
    return dev, ref


def test_first_iteration_intermediates_match_oracle():
    """Kernel-level parity after one iteration: S, residuals and directions."""
    sdp = workloads.delsarte(8, 3, Fraction(1, 2))
    d = Solver(sdp, lib="device"); o = Solver(sdp, lib="oracle")
    d.iterate(); o.iterate()
    with mpmath.workprec(400):
        for what, j, l in [("S", 0, 0), ("d", 0, 0), ("p", 0, 0), ("dx", 0, 0), ("dy", 0, 0), ("dX", 0, 6), ("dY", 0, 7), ("X", 0, 6), ("Y", 0, 7)]:
            a = wire.from_wire(d.debug_get(what, j, l), PREC); b = wire.from_wire(o.debug_get(what, j, l), PREC)
            scale = max([abs(v) for v in b] + [mpmath.mpf(2) ** -200])
            err = max(abs(x - y) for x, y in zip(a, b))
            # X, Y contain the step lengths, which come from a Float64 eigenvalue (src/solver.jl:1659-1662):
            # the reference itself only reproduces them to its Lanczos tolerance 1e-5
            tol = mpmath.mpf(10) ** (-9 if what in ("X", "Y") else -55)
            assert err <= scale * tol, (what, float(err / scale))
    d.close(); o.close()


def test_maxcut_three_cycle_known_answer():
    dev, _ = _compare(workloads.maxcut(workloads.laplacian_cycle(3)))
    assert abs(dev.p_obj - mpmath.mpf(9) / 4) < mpmath.mpf(10) ** -29      # README.md:70-72


def test_polyopt_x2_plus_1():
    dev, _ = _compare(workloads.polyopt(lambda x: x * x + 1, 1))
    assert abs(dev.p_obj - 1) < mpmath.mpf(10) ** -29                       # README.md:146-150


def test_delsarte_e8_is_240():
    dev, _ = _compare(workloads.delsarte(8, 3, Fraction(1, 2)))
    assert abs(dev.p_obj - 240) < mpmath.mpf(10) ** -25                     # test/runtests_solver.jl:86-87


def test_maxcut_complete_graph_dense_path():
    n = 12
    dev, _ = _compare(workloads.maxcut(workloads.laplacian_complete(n)))
    assert abs(dev.p_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -25   # K_n: n^2/4


def test_polyopt_config1_shape():
    """BASELINE config 1: degree-40 polynomial, Chebyshev basis/samples, one rank-1 block of size 21."""
    _compare(workloads.polyopt_random(20, seed=0))


def test_delsarte_config3_d16():
    dev, _ = _compare(workloads.delsarte(8, 16, Fraction(1, 2)))
    assert abs(dev.p_obj - 240) < mpmath.mpf(10) ** -20


def test_sphere_packing_two_radii_small():
    """Config 5 family at small degree: several clusters, subblocks m=2, 50 free variables."""
    sdp = workloads.sphere_packing(8, 7, [Fraction(1, 2), Fraction(1, 2)])
    # at 256 bit this ill-conditioned family runs out of precision near gap 1e-28 in BOTH arms
    # (the reference's own test uses prec=300, test/runtests_solver.jl:21); compare at 1e-20
    dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-20)
    ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-20)
    assert dev.status == ref.status == "Optimal", (dev, ref)
    assert abs(dev.p_obj - ref.p_obj) <= mpmath.mpf(10) ** -19 and abs(dev.d_obj - ref.d_obj) <= mpmath.mpf(10) ** -19
    assert abs(dev.iterations - ref.iterations) <= 1


@pytest.mark.parametrize("M,N,K", [(128, 32, 96), (130, 40, 100), (257, 100, 200)])
def test_gemm_tcgen05_bitexact_vs_host_arithmetic(tiny, M, N, K):
    A = rand_wire(M + N, (M, K), 20); B = rand_wire(K + N, (K, N), 20)
    Cd, _ = tiny.mp_gemm(A, B, path=2)
    Ch = wire.wire_zeros((M, N), PREC)
    hostcheck().hc_gemm(M, N, K, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), Ch.ctypes.data_as(C.c_void_p))
    assert Cd.tobytes() == Ch.tobytes()


@pytest.mark.parametrize("M,N,K,spread", [(19200, 112, 100, 5), (1, 1, 1, 0), (129, 17, 33, 3), (200, 64, 224, 10),
                                          (19200, 300, 100, 5),      # 160 + 144 column tiles, three diagonals per group
                                          (300, 300, 300, 5), (130, 290, 140, 30), (257, 160, 128, 3), (128, 128, 3500, 8)])   # few tiles: one CTA per (tile, K range, diagonal group), raw sums added exactly
def test_gemm_tcgen05_equals_cuda_core_path(tiny, M, N, K, spread):
    """Both GEMM paths compute the same exact integer slice-pair sums: identical bits (shapes whose K ranges are combined
    exactly: one K range per tile, or the raw-sum mode of products with few tiles)."""
    A = rand_wire(1, (M, K), spread); B = rand_wire(2, (K, N), spread)
    C1, _ = tiny.mp_gemm(A, B, path=1)
    C2, _ = tiny.mp_gemm(A, B, path=2)
    assert C1.tobytes() == C2.tobytes()


@pytest.mark.parametrize("M,N,K,spread", [(32, 300, 200, 5), (17, 129, 96, 50), (64, 481, 224, 3)])
def test_gemm_short_left_operand_runs_swapped_on_tensor_cores(tiny, M, N, K, spread):
    """A short panel against a long right-hand side (block rows of the triangular solves) is routed to the tcgen05
    kernel with the operands swapped and the result written transposed: same bits as the CUDA-core path."""
    A = rand_wire(3, (M, K), spread); B = rand_wire(4, (K, N), spread)
    C0, _ = tiny.mp_gemm(A, B, path=0)
    C1, _ = tiny.mp_gemm(A, B, path=1)
    assert C0.tobytes() == C1.tobytes()


@pytest.mark.parametrize("M,N,K,spread", [(300, 300, 300, 5), (512, 304, 300, 200), (256, 48, 3800, 10)])
def test_gemm_tcgen05_k_split(tiny, M, N, K, spread):
    """Products with few output tiles or K beyond the int32 headroom (35 K 2^14 < 2^31) run as split-K over
    blockIdx.z; the truncated partial results are added, so the result equals the single-pass CUDA-core
    result up to the rounding of those additions (normwise, relative to rowmax * colmax)."""
    A = rand_wire(1, (M, K), spread); B = rand_wire(2, (K, N), spread)
    C1, _ = tiny.mp_gemm(A, B, path=1)
    C2, _ = tiny.mp_gemm(A, B, path=2)
    with mpmath.workprec(400):
        a, b = wire.from_wire(C1, PREC), wire.from_wire(C2, PREC)
        am, bm = wire.from_wire(A, PREC), wire.from_wire(B, PREC)
        rmax = [max(abs(v) for v in am[i, :]) for i in range(M)]
        cmax = [max(abs(v) for v in bm[:, j]) for j in range(N)]
        for i in range(0, M, 7):
            for j in range(0, N, 5):
                assert abs(a[i, j] - b[i, j]) <= rmax[i] * cmax[j] * K * mpmath.mpf(2) ** -250


def test_maxcut_complete_graph_tensor_core_block():
    """n = 130 > 128: the X/Y block products and the dense Schur path run on the tcgen05 kernel."""
    n = 130
    dev = solvesdp(workloads.maxcut(workloads.laplacian_complete(n)), lib="device", duality_gap_threshold=1e-30)
    assert dev.status == "Optimal"
    assert abs(dev.p_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -24
    assert abs(dev.d_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -24


def _golden(name):
    import json
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", name + ".json")))


@pytest.mark.parametrize("name,make", [
    ("maxcut300_seed0", lambda: workloads.maxcut(workloads.laplacian_random(300, 0.5, 0))),      # BASELINE config 2 at full size
    ("maxcut130_seed1", lambda: workloads.maxcut(workloads.laplacian_random(130, 0.5, 1))),
    ("polyopt20_seed0", lambda: workloads.polyopt_random(20, 0)),                                # config 1
    ("delsarte_8_16", lambda: workloads.delsarte(8, 16, Fraction(1, 2))),                        # config 3
    ("delsarte_8_32", lambda: workloads.delsarte(8, 32, Fraction(1, 2))),                        # config 3 at full size (dimension 8, degree 32)
    ("sphere_2_31_prec512", lambda: workloads.sphere_packing(8, 31, [Fraction(1, 2), Fraction(1, 2)], prec=512)),   # config 5, reference default precision
    ("threepoint_4_10_10", lambda: workloads.three_point_bound(4, Fraction(1, 6), 10, 10)),      # config 4
])
def test_full_size_configs_against_golden_oracle_values(name, make):
    """Objectives and gap to a relative 1e-25, iteration count within +-1 (the north-star parity bar);
    the golden values come from the CPU oracle (tests/golden/make_golden.py)."""
    g = _golden(name)
    kw = {k: int(v) for k, v in g["options"].items() if k.startswith("omega")}
    dev = solvesdp(make(), lib="device", duality_gap_threshold=1e-30, **kw)
    assert dev.status == g["status"] == "Optimal"
    with mpmath.workprec(700):
        for key, val in (("p_obj", dev.p_obj), ("d_obj", dev.d_obj)):
            ref = mpmath.mpf(g[key])
            assert abs(val - ref) <= TOL_OBJ * max(1, abs(ref)), (key, val, ref)
        assert abs(dev.gap - mpmath.mpf(g["gap"])) <= TOL_OBJ
    assert abs(dev.iterations - g["iterations"]) <= 1


def test_dense_schur_equals_hadamard_identity_at_full_size():
    """Size-independent property of config 2: with A_p = E_pp the Schur complement is
    S = X^-1 o Y (Hadamard); checked on the n = 300 instance after two iterations."""
    sdp = workloads.maxcut(workloads.laplacian_random(300, 0.5, 0))
    s = Solver(sdp, lib="device")
    s.iterate(); s.iterate()
    x, X, y, Y = s.get_state()
    s.iterate()      # S and X^-1 of this iteration are built from the (X, Y) just read
    n = 300
    Ls = wire.from_wire(s.debug_get("S", 0, 0), PREC).reshape(n, n)      # S is held as its Cholesky factor
    Xinv = wire.from_wire(s.debug_get("Xinv", 0, 0), PREC).reshape(n, n)
    Yw = wire.from_wire(Y, PREC).reshape(n, n)
    Xw = wire.from_wire(X, PREC).reshape(n, n)
    rng = random.Random(0)
    with mpmath.workprec(400):
        scale = max(abs(Xinv[i, i] * Yw[i, i]) for i in range(n))
        for _ in range(25):
            i, j = rng.randrange(n), rng.randrange(n)
            sij = mpmath.fsum(Ls[i, k] * Ls[j, k] for k in range(min(i, j) + 1))
            assert abs(sij - Xinv[i, j] * Yw[i, j]) <= scale * mpmath.mpf(10) ** -60, (i, j)
        for _ in range(3):      # X^-1 really inverts X (sampled rows)
            i = rng.randrange(n)
            row = [mpmath.fsum(Xinv[i, k] * Xw[k, j] for k in range(n)) for j in range(n)]
            assert max(abs(v - (1 if j == i else 0)) for j, v in enumerate(row)) < mpmath.mpf(10) ** -50
    s.close()


def test_prec_300_uses_ten_limbs_and_matches_oracle():
    """prec = 300 (the reference's setting for the sphere-packing family, test/runtests_solver.jl:21):
    the 10-limb instantiation; both arms reach gap 1e-30 here."""
    sdp = workloads.sphere_packing(8, 7, [Fraction(1, 2), Fraction(1, 2)], prec=300)
    dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30)
    ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30)
    assert dev.status == ref.status == "Optimal", (dev, ref)
    with mpmath.workprec(500):
        assert abs(dev.p_obj - ref.p_obj) <= TOL_OBJ * abs(ref.p_obj)
        assert abs(dev.d_obj - ref.d_obj) <= TOL_OBJ * abs(ref.d_obj)
    assert abs(dev.iterations - ref.iterations) <= 1


def test_prec_300_dense_path_known_answer():
    n = 9
    dev = solvesdp(workloads.maxcut(workloads.laplacian_complete(n), prec=300), lib="device", duality_gap_threshold=1e-40)
    assert dev.status == "Optimal" and abs(dev.p_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -35


def test_prec_512_sixteen_limbs_sphere_packing():
    """prec = 512 is the default of the reference's examples/SpherePacking.jl:13 (16-limb instantiation).  Degree 15 with
    two radii; both arms reach gap 1e-30 and must agree far below it."""
    sdp = workloads.sphere_packing(8, 15, [Fraction(1, 2), Fraction(1, 2)], prec=512)
    dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30)
    ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30)
    assert dev.status == ref.status == "Optimal", (dev, ref)
    with mpmath.workprec(700):
        assert abs(dev.p_obj - ref.p_obj) <= mpmath.mpf(10) ** -40 * abs(ref.p_obj)
        assert abs(dev.d_obj - ref.d_obj) <= mpmath.mpf(10) ** -40 * abs(ref.d_obj)
    assert abs(dev.iterations - ref.iterations) <= 1


def test_prec_512_dense_path_known_answer():
    n = 40                                                  # > 32: blocked Cholesky and the tensor-core split at 67 slices
    dev = solvesdp(workloads.maxcut(workloads.laplacian_complete(n), prec=512), lib="device", duality_gap_threshold=1e-60)
    assert dev.status == "Optimal" and abs(dev.p_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -55


def test_warp_cooperative_arithmetic_selftest(tiny):
    """mpw.cuh (limb-per-lane multiply with ballot carry resolution, used by the Cholesky pivot chain) must give
    the same bits as the single-thread routines on 4096 random operand pairs."""
    fn = tiny.lib.clrs_debug_selftest
    fn.restype = C.c_int
    assert fn(tiny.h) == 0
    s16 = Solver(workloads.maxcut(workloads.laplacian_cycle(3), prec=512), lib="device")      # 16 limbs: 32 product columns fill the warp
    try:
        assert fn(s16.h) == 0
    finally:
        s16.close()


def test_three_point_bound_config4():
    """BASELINE config 4 at the reference's test size: dense F_k blocks, rank-1 and rank-2 invariant SOS blocks."""
    sdp = workloads.three_point_bound(4, Fraction(1, 6), -1, 4)
    dev, _ = _compare(sdp, omega_p=10 ** 3, omega_d=10 ** 3)
    assert abs(dev.p_obj - 10) < mpmath.mpf(10) ** -24           # test/runtests_solver.jl:26-27


def test_three_point_bound_with_univariate_part():
    """d2 >= 0 adds the 1x1 a_k blocks with sample-dependent eigenvalues."""
    _compare(workloads.three_point_bound(4, Fraction(1, 6), 3, 3), omega_p=10 ** 3, omega_d=10 ** 3)


# ---- solver options and stop reasons (src/solver.jl:100-137 keyword arguments, terminate() :921-950) ----

def _hist(r):
    return [(h["iter"], h["alpha_p"], h["alpha_d"], h["beta_c"]) for h in r.history]


@pytest.mark.parametrize("kw", [
    dict(need_dual_feasible=True),
    dict(need_primal_feasible=True),
    dict(correctoronly=True),
    dict(safe_step=False),
    dict(gamma=Fraction(7, 10), beta_feasible=Fraction(1, 5), beta_infeasible=Fraction(2, 5)),
    dict(omega_p=10 ** 4, omega_d=10 ** 6),
    dict(maxiterations=7),
])
def test_solver_options_take_the_same_path_as_the_oracle(kw):
    """Every keyword of solvesdp that changes the iteration must change it the same way in both arms: same status,
    same error code, same iteration count, same step lengths and centering parameters (to double rounding)."""
    sdp = workloads.polyopt_random(8, 3)
    dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30, **kw)
    ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30, **kw)
    assert (dev.status, dev.error_code, dev.iterations) == (ref.status, ref.error_code, ref.iterations), (dev, ref)
    for a, b in zip(_hist(dev), _hist(ref)):
        assert a[0] == b[0] and all(abs(x - y) <= 1e-9 * max(1.0, abs(y)) for x, y in zip(a[1:], b[1:])), (a, b)
    # converged runs agree far below the gap; runs stopped early differ by the Float64 step lengths (alpha comes from a
    # double-precision eigenvalue, so the iterates agree to ~1e-16 relative until the optimum pulls them together)
    tol = mpmath.mpf(10) ** -20 if ref.status == "Optimal" else mpmath.mpf(10) ** -8
    with mpmath.workprec(400):
        assert abs(dev.p_obj - ref.p_obj) <= tol * max(1, abs(ref.p_obj))


def test_indefinite_iterate_raises_the_same_solver_failure_as_the_oracle():
    """A warm start whose X is not positive definite must stop in the Cholesky factorisation of X with the reference's
    SolverFailure (src/solver.jl:389-392, src/tools.jl:92-95) in both arms, not return numbers."""
    sdp = workloads.polyopt_random(6, 1)
    msgs = []
    for lib in ("device", "oracle"):
        s = Solver(sdp, lib=lib)
        x, X, y, Y = s.get_state()
        X = X.copy(); X["sign"] = -X["sign"]                      # X = -omega_p I
        s.set_state(x, X, y, Y)
        with pytest.raises(clrs_b200.SolverFailure) as e:
            s.iterate()
        msgs.append(str(e.value))
        # the reference throws before the step (src/solver.jl:394-398) and hands back the last good iterate (:594-628):
        # the state after the failed call is the state that was set, bit for bit
        x2, X2, y2, Y2 = s.get_state()
        assert X2.tobytes() == X.tobytes() and Y2.tobytes() == Y.tobytes() and x2.tobytes() == x.tobytes() and y2.tobytes() == y.tobytes(), lib
        s.close()
    assert msgs[0][:4] == msgs[1][:4] and msgs[0].startswith("[1"), msgs      # same failure site code


def test_failed_schur_factorisation_takes_no_step():
    """A Cholesky failure later in the iteration (of S or Q: sphere packing (4,31) cannot be solved at 256 bit, the oracle
    stops with "Q was not decomposed correctly" too) must not apply the step built from the broken factor: the iterate
    stays the one before the call."""
    sdp = workloads.sphere_packing(8, 31, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4), Fraction(1)], prec=256)
    s = Solver(sdp, lib="device")
    before = None
    with pytest.raises(clrs_b200.SolverFailure):
        for _ in range(200):
            before = s.get_state()
            s.iterate()
    after = s.get_state()
    for a, b in zip(before, after):
        assert a.tobytes() == b.tobytes()
    s.close()


def test_cuda_graph_replay_is_bit_identical_to_eager_launches():
    """From its second iteration on a handle replays the iteration as a CUDA graph; the numbers must not depend on it."""
    for make in (lambda: workloads.sphere_packing(8, 5, [Fraction(1, 2), Fraction(1, 2)]), lambda: workloads.maxcut(workloads.laplacian_cycle(7)),
                 lambda: workloads.maxcut(workloads.laplacian_complete(130))):
        sdp = make()
        a = Solver(sdp, lib="device"); b = Solver(sdp, lib="device"); b.use_graph(False)
        for _ in range(7):
            ia, ib = a.iterate(), b.iterate()
            assert ia.stop == ib.stop == 0
            assert ia.p_obj_new == ib.p_obj_new and ia.d_obj_new == ib.d_obj_new and ia.alpha_p == ib.alpha_p and ia.alpha_d == ib.alpha_d and ia.mu == ib.mu
        for u, v in zip(a.get_state(), b.get_state()):
            assert u.tobytes() == v.tobytes()
        assert all(t >= 0 for t in ia.phase_ms) and sum(ia.phase_ms[12:17]) > 0          # the phase timers survive the replay
        a.close(); b.close()


def test_warm_start_continues_identically_on_the_device():
    """clrs_get_state / clrs_set_state (src/solver.jl:202-239): a solver restarted from a saved iterate takes the same
    iterations as the uninterrupted one (the state crosses the ABI as exact wire numbers)."""
    sdp = workloads.sphere_packing(8, 5, [Fraction(1, 2), Fraction(1, 2)])
    a = Solver(sdp, lib="device")
    for _ in range(6):
        a.iterate()
    x, X, y, Y = a.get_state()
    b = Solver(sdp, lib="device")
    b.set_state(x, X, y, Y)
    for _ in range(5):
        ia, ib = a.iterate(), b.iterate()
        assert ia.p_obj == ib.p_obj and ia.d_obj == ib.d_obj and ia.alpha_p == ib.alpha_p and ia.mu == ib.mu
    xa, Xa, ya, Ya = a.get_state(); xb, Xb, yb, Yb = b.get_state()
    assert Xa.tobytes() == Xb.tobytes() and xa.tobytes() == xb.tobytes() and Ya.tobytes() == Yb.tobytes()
    a.close(); b.close()


def test_maxcut_complete_graph_n100_dense_schur_on_tensor_cores():
    """n = 100 < 128: the block products stay small, but the (P n) x n x n dense Schur products are tensor-core shaped and
    decide the panel layout of the block."""
    n = 100
    dev = solvesdp(workloads.maxcut(workloads.laplacian_complete(n)), lib="device", duality_gap_threshold=1e-30)
    assert dev.status == "Optimal"
    assert abs(dev.p_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -24
    assert abs(dev.d_obj - mpmath.mpf(n * n) / 4) < mpmath.mpf(10) ** -24


def test_delsarte_irrational_angles_known_answers_on_device():
    """test/runtests_solver.jl:98-111 and :124-125: the Delsarte bound is 120 for (n=4, d=9, cos = 1/(sqrt5 - 1)) and 12 for
    (n=3, d=2, cos = 1/sqrt5)."""
    with mpmath.workprec(400):
        for n, d, ct, ans in ((4, 9, 1 / (mpmath.sqrt(5) - 1), 120), (3, 2, 1 / mpmath.sqrt(5), 12)):
            dev = solvesdp(workloads.delsarte(n, d, ct), lib="device", duality_gap_threshold=1e-30)
            assert dev.status == "Optimal" and abs(dev.p_obj - ans) < mpmath.mpf(10) ** -26, (n, d, dev)


def test_lovasz_theta_cycles_on_device():
    """General dense constraint matrices (identity + off-diagonal pairs): theta(C_5) = sqrt 5 (test/moi_tests.jl:7-8), and a
    40-cycle whose block runs the blocked Cholesky."""
    with mpmath.workprec(400):
        dev = solvesdp(workloads.lovasz_theta_cycle(5), lib="device", duality_gap_threshold=1e-30)
        assert dev.status == "Optimal" and abs(dev.p_obj - mpmath.sqrt(5)) < mpmath.mpf(10) ** -26
        n = 41
        dev = solvesdp(workloads.lovasz_theta_cycle(n), lib="device", duality_gap_threshold=1e-30)
        assert dev.status == "Optimal" and abs(dev.p_obj - n * mpmath.cos(mpmath.pi / n) / (1 + mpmath.cos(mpmath.pi / n))) < mpmath.mpf(10) ** -24


@pytest.mark.gpu
def test_device_reproduces_the_reference_solver_log_of_min_f_2():
    """The same comparison as tests/test_oracle_pins.py, on the device: 56 iterations, the printed rows 1-3 and 55-56 of
    docs/src/solving.md:38-52 and the reference's final objectives."""
    from test_oracle_pins import check_against_reference_log
    r = solvesdp(workloads.min_f(2), lib="device")
    check_against_reference_log(r)


@pytest.mark.gpu
def test_two_radii_sphere_packing_d15_prec300_reference_test_on_device():
    """test/runtests_solver.jl:21-22 on the device (10 limbs): pi^4/384 to 1e-4."""
    r = solvesdp(workloads.sphere_packing(8, 15, [Fraction(1, 2), Fraction(1, 2)], prec=300), lib="device")
    with mpmath.workprec(200):
        assert r.status == "Optimal" and 0 < r.p_obj - mpmath.pi ** 4 / 384 < mpmath.mpf(10) ** -4
        assert abs(r.p_obj - mpmath.pi ** 4 / 384 - mpmath.mpf("7.0919e-5")) < mpmath.mpf(10) ** -8


@pytest.mark.gpu
def test_cohn_elkies_8_15_reference_test_on_device():
    """test/runtests_solver.jl:19-20 on the device: cohnelkies(8, 15, prec=256) = pi^4/384 to 1e-4, and the oracle's value and iteration count."""
    sdp = workloads.cohnelkies(8, 15)
    r = solvesdp(sdp, lib="device")
    o = solvesdp(sdp, lib="oracle")
    with mpmath.workprec(200):
        assert r.status == o.status == "Optimal" and 0 < r.p_obj - mpmath.pi ** 4 / 384 < mpmath.mpf(10) ** -4
        assert abs(r.p_obj - o.p_obj) < mpmath.mpf(10) ** -14 and abs(r.iterations - o.iterations) <= 1


@pytest.mark.gpu
def test_three_point_bound_d14_first_two_iterations_match_the_oracle():
    """BASELINE config 4 at its largest survey shape (d2 = d3 = 14: P = 894, 61 blocks up to n = 225).  A full oracle solve of it
    takes hours on the CPUs of the build container (~9 minutes per iteration), so the golden holds the first two iterations:
    the printed row of each and the objectives of the iterate after them.  The step lengths carry the Float64 eigenvalue, so the
    comparison is to 1e-7, not to the last bit."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "threepoint_4_14_14_first2.json")
    if not os.path.exists(path):
        pytest.skip("golden not generated (tests/golden/make_golden.py threepoint_4_14_14_first2)")
    import json
    g = json.load(open(path))
    sdp = workloads.three_point_bound(4, Fraction(1, 6), 14, 14)
    assert sdp.describe() == g["describe"]
    r = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30, maxiterations=2, omega_p=10 ** 3, omega_d=10 ** 3)
    assert r.iterations == g["iterations"] == 2 and r.error_code == 2
    for hd, hg in zip(r.history, g["history"]):
        for k in ("mu", "d_obj", "p_obj", "gap", "err_P", "err_p", "err_d", "alpha_d", "alpha_p", "beta_c", "d_obj_new", "p_obj_new", "gap_new"):
            assert abs(hd[k] - hg[k]) <= 1e-7 * max(abs(hg[k]), 1e-300), (k, hd[k], hg[k])
    with mpmath.workprec(300):
        assert abs(r.d_obj - mpmath.mpf(g["d_obj"])) <= mpmath.mpf(10) ** -7 * abs(mpmath.mpf(g["d_obj"]))
        assert abs(r.p_obj - mpmath.mpf(g["p_obj"])) <= mpmath.mpf(10) ** -7 * abs(mpmath.mpf(g["p_obj"]))


@pytest.mark.gpu
def test_povm_two_states_on_device():
    """test/moi_tests.jl:9-10 on the device: 1/2 + sqrt(2)/4 to 1e-30, same iteration count as the oracle (two dense blocks sharing constraints)."""
    sdp = workloads.povm_two_states()
    dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30)
    ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30)
    with mpmath.workprec(300):
        target = mpmath.mpf(1) / 2 + mpmath.sqrt(2) / 4
        assert dev.status == "Optimal" and abs(dev.p_obj - target) < mpmath.mpf(10) ** -30 and abs(dev.d_obj - target) < mpmath.mpf(10) ** -30
        assert abs(dev.iterations - ref.iterations) <= 1 and abs(dev.p_obj - ref.p_obj) < mpmath.mpf(10) ** -25

"""Device tests of the triplet upload (clrs_add_sparse_term), the SDPA reader and the opt-in sparsity shortcut of the dense Schur
path (clrs_options.sparse_schur; SURVEY.md §8(f)2,4), against the CPU oracle, which always takes the reference's dense route."""
import ctypes as C
import json
import os

import mpmath
import numpy as np
import pytest

from clrs_b200 import sdpa, solvesdp, Solver, wire, workloads

pytestmark = pytest.mark.gpu
PREC = 256
TOL_OBJ = mpmath.mpf(10) ** -25


def _is_sparse(S, j=0, l=0):
    fn = S._fn("debug_get"); fn.restype = C.c_int64
    return int(fn(S.h, b"sparse?", C.c_int32(j), C.c_int32(l), None, C.c_int64(0)))


def test_triplet_upload_is_bit_identical_to_dense_upload():
    L = workloads.laplacian_random(40, 0.5, 3)
    a = Solver(workloads.maxcut(L), lib="device"); b = Solver(sdpa.sdpa_sparse_to_sdp(sdpa.maxcut_sdpa_text(L)), lib="device")
    for _ in range(3):
        a.iterate(); b.iterate()
    for u, v in zip(a.get_state(), b.get_state()):
        assert u.tobytes() == v.tobytes()
    assert _is_sparse(a) == 0 and _is_sparse(b) == 0          # the shortcut is opt-in
    a.close(); b.close()


@pytest.mark.parametrize("n,prec", [(5, 256), (9, 256), (7, 512)])
def test_sparse_schur_first_iteration_matches_oracle(n, prec):
    """S, residuals and directions after one iteration; constraint matrices with 1, 2 and n nonzeros (identity + edge pairs)."""
    PREC = prec
    sdp = workloads.lovasz_theta_cycle(n, prec=prec)
    d = Solver(sdp, lib="device", sparse_schur=True); o = Solver(sdp, lib="oracle")
    assert _is_sparse(d) == 1
    d.iterate(); o.iterate()
    with mpmath.workprec(prec + 200):
        for what in ("S", "d", "dx", "dX", "dY"):
            a = wire.from_wire(d.debug_get(what, 0, 0), PREC); b = wire.from_wire(o.debug_get(what, 0, 0), PREC)
            scale = max(abs(v) for v in b)
            assert max(abs(x - y) for x, y in zip(a, b)) <= scale * mpmath.mpf(10) ** (-55 if prec == 256 else -130), what
    d.close(); o.close()


def test_sparse_schur_solves_theta_and_maxcut_like_the_oracle():
    for sdp, known in ((workloads.lovasz_theta_cycle(5), 5), (workloads.maxcut(workloads.laplacian_cycle(7)), None)):
        dev = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30, sparse_schur=True)
        ref = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30)
        assert dev.status == ref.status == "Optimal"
        with mpmath.workprec(400):
            assert abs(dev.p_obj - ref.p_obj) <= TOL_OBJ * max(1, abs(ref.p_obj)) and abs(dev.d_obj - ref.d_obj) <= TOL_OBJ * max(1, abs(ref.d_obj))
            if known is not None:
                assert abs(dev.p_obj - mpmath.sqrt(known)) < mpmath.mpf(10) ** -25       # theta(C_5) = sqrt 5 (test/moi_tests.jl:7-8)
        assert abs(dev.iterations - ref.iterations) <= 1


def test_maxcut300_from_sdpa_with_sparse_schur_matches_the_golden_solution():
    """BASELINE config 2 read from SDPA text, triplet upload, Schur complement as X^-1 o Y: same optimum and iteration count
    as the oracle's dense solve (tests/golden/maxcut300_seed0.json)."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "maxcut300_seed0.json")))
    L = workloads.laplacian_random(300, 0.5, 0)
    res = solvesdp(sdpa.sdpa_sparse_to_sdp(sdpa.maxcut_sdpa_text(L)), lib="device", duality_gap_threshold=1e-30, sparse_schur=True, keep_solver=True)
    assert _is_sparse(res.solver) == 1
    res.solver.close()
    assert res.status == "Optimal"
    with mpmath.workprec(400):
        p = mpmath.mpf(g["p_obj"]); d = mpmath.mpf(g["d_obj"])
        assert abs(res.p_obj - p) <= TOL_OBJ * abs(p) and abs(res.d_obj - d) <= TOL_OBJ * abs(d)
    assert abs(res.iterations - g["iterations"]) <= 1


def test_matmul_prec_produces_fewer_diagonals_and_still_converges():
    """matmul_prec (src/solver.jl:93,125): the T Y product of the dense path at 128 bit = 19 of the 35 slice-pair diagonals on the
    tensor cores.  S of the second iteration moves by about 2^-146 relative (not 0: the option is active; not more: only the
    least significant diagonals are dropped), and the solve still reaches gap 1e-15 at the golden optimum."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "maxcut130_seed1.json")))
    sdp = workloads.maxcut(workloads.laplacian_random(130, 0.5, 1))
    full = Solver(sdp, lib="device"); low = Solver(sdp, lib="device", matmul_prec=128)
    for _ in range(2):                      # (the start X = omega_p I, Y = omega_d I has so few digits that the dropped diagonals are exact zeros)
        full.iterate(); low.iterate()
    with mpmath.workprec(400):
        a = wire.from_wire(full.debug_get("S", 0, 0), PREC); b = wire.from_wire(low.debug_get("S", 0, 0), PREC)
        scale = max(abs(v) for v in a)
        rel = max(abs(x - y) for x, y in zip(a, b)) / scale
        assert mpmath.mpf(2) ** -200 < rel < mpmath.mpf(2) ** -110, mpmath.nstr(rel, 5)
    full.close(); low.close()
    res = solvesdp(sdp, lib="device", duality_gap_threshold=1e-15, matmul_prec=128)
    assert res.status == "Optimal"
    with mpmath.workprec(400):
        p = mpmath.mpf(g["p_obj"])
        assert abs(res.p_obj - p) <= mpmath.mpf(10) ** -13 * abs(p)


def test_sparse_term_arguments_are_validated():
    """Out-of-range entries and repeated positions are refused (the scatter kernel writes without ordering)."""
    from clrs_b200 import PSDBlock, Cluster, ClusteredSDP
    one = wire.to_wire(1, PREC)

    def make(rows, cols, mirror):
        blk = PSDBlock(m=1, delta=3, high_rank=True, C=wire.wire_eye_scaled(3, 1, PREC))
        blk.sparse[0] = (np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32), wire.to_wire([1] * len(rows), PREC), mirror)
        return ClusteredSDP(prec=PREC, maximize=True, constant=wire.to_wire(0, PREC), b=wire.wire_zeros((0,), PREC),
                            clusters=[Cluster(B=wire.wire_zeros((1, 0), PREC), c=wire.to_wire([1], PREC), blocks=[blk])])
    for rows, cols, mirror in (([0, 0], [1, 1], False), ([0, 1], [1, 0], True), ([3], [0], False), ([0], [-1], False)):
        with pytest.raises(RuntimeError):
            Solver(make(rows, cols, mirror), lib="device")
    Solver(make([0, 1], [1, 0], False), lib="device").close()           # both triangles listed explicitly: fine without mirroring

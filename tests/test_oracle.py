"""Internal consistency of the oracle and of the committed golden fixtures."""
import json
import os
from fractions import Fraction

import mpmath

import clrs_b200
from clrs_b200 import workloads, Solver, solvesdp, wire

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_zero_skipping_mode_is_bit_identical_to_the_plain_dense_path():
    sdp = workloads.maxcut(workloads.laplacian_random(9, seed=5))
    a = Solver(sdp, lib="oracle"); b = Solver(sdp, lib="oracle", oracle_skip_zeros=True)
    for it in range(4):
        a.iterate(); b.iterate()
        for what in ("S", "X", "Y", "dX"):
            assert a.debug_get(what, 0, 0).tobytes() == b.debug_get(what, 0, 0).tobytes(), (what, it)
    a.close(); b.close()


def test_golden_fixture_polyopt20_reproduces():
    g = json.load(open(os.path.join(GOLD, "polyopt20_seed0.json")))
    r = solvesdp(workloads.polyopt_random(20, 0), lib="oracle", duality_gap_threshold=1e-30)
    assert r.iterations == g["iterations"] and r.status == g["status"]
    with mpmath.workprec(300):
        assert abs(r.p_obj - mpmath.mpf(g["p_obj"])) <= abs(r.p_obj) * mpmath.mpf(10) ** -60


def test_golden_fixtures_are_complete():
    for name in ("maxcut300_seed0", "maxcut130_seed1", "polyopt20_seed0", "delsarte_8_16"):
        g = json.load(open(os.path.join(GOLD, name + ".json")))
        assert g["status"] == "Optimal" and float(g["gap"]) < 1e-30
    g = json.load(open(os.path.join(GOLD, "delsarte_8_16.json")))
    with mpmath.workprec(300):
        assert abs(mpmath.mpf(g["p_obj"]) - 240) < mpmath.mpf(10) ** -27     # test/runtests_solver.jl:86-87


def test_warm_start_round_trip():
    """get_state / set_state (src/solver.jl:202-239): restarting from a saved iterate continues identically."""
    sdp = workloads.delsarte(8, 3, Fraction(1, 2))
    a = Solver(sdp, lib="oracle")
    for _ in range(5):
        a.iterate()
    x, X, y, Y = a.get_state()
    b = Solver(sdp, lib="oracle"); b.set_state(x, X, y, Y)
    ia, ib = a.iterate(), b.iterate()
    assert abs(ia.mu - ib.mu) <= 1e-12 * abs(ia.mu) and abs(ia.p_obj_new - ib.p_obj_new) <= 1e-12 * abs(ia.p_obj_new)
    a.close(); b.close()

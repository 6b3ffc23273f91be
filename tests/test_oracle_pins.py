"""Pin the CPU oracle against the known-answer values the reference's own tests and README
hold for this path; Julia/Arb cannot run in this image."""
from fractions import Fraction

import mpmath
import pytest

import clrs_b200
from clrs_b200 import workloads, solvesdp


def solve(sdp, gap=1e-30, **kw):
    r = solvesdp(sdp, lib="oracle", duality_gap_threshold=gap, oracle_skip_zeros=True, **kw)
    assert r.status == "Optimal" and r.error_code == 0, r
    return r


def test_maxcut_three_cycle_is_nine_quarters():          # README.md:70-72, 101-103
    r = solve(workloads.maxcut(workloads.laplacian_cycle(3)))
    assert abs(r.p_obj - mpmath.mpf(9) / 4) < mpmath.mpf(10) ** -29
    assert abs(r.d_obj - mpmath.mpf(9) / 4) < mpmath.mpf(10) ** -29


def test_maxcut_complete_graph_and_odd_cycle():           # derived family, SURVEY.md §8(c) last row
    r = solve(workloads.maxcut(workloads.laplacian_complete(10)))
    assert abs(r.p_obj - 25) < mpmath.mpf(10) ** -27
    n = 7
    r = solve(workloads.maxcut(workloads.laplacian_cycle(n)))
    with mpmath.workprec(300):
        assert abs(r.p_obj - mpmath.mpf(n) / 4 * (2 + 2 * mpmath.cos(mpmath.pi / n))) < mpmath.mpf(10) ** -27


def test_polyopt_min_of_x2_plus_1_is_1():                  # README.md:146-150
    r = solve(workloads.polyopt(lambda x: x * x + 1, 1))
    assert abs(r.p_obj - 1) < mpmath.mpf(10) ** -29


def test_polyopt_chebyshev_square_has_minimum_zero():      # config-1 shape, f = T_5(x)^2
    d = 5
    r = solve(workloads.polyopt(lambda x: workloads.chebyshev_values(d, x)[d] ** 2, d), gap=1e-25)
    assert abs(r.p_obj) < mpmath.mpf(10) ** -22


def test_delsarte_e8_kissing_number_240():                 # test/runtests_solver.jl:86-87
    r = solve(workloads.delsarte(8, 3, Fraction(1, 2)))
    assert abs(r.p_obj - 240) < mpmath.mpf(10) ** -26


def test_delsarte_24_cell_and_icosahedron_irrational_angles():   # test/runtests_solver.jl:98-111 (n=4, d=9 -> 120), :124-125,158 (n=3, d=2 -> 12)
    with mpmath.workprec(400):
        r = solve(workloads.delsarte(4, 9, 1 / (mpmath.sqrt(5) - 1)))
        assert abs(r.p_obj - 120) < mpmath.mpf(10) ** -26
        r = solve(workloads.delsarte(3, 2, 1 / mpmath.sqrt(5)))
        assert abs(r.p_obj - 12) < mpmath.mpf(10) ** -26


def test_delsarte_3_10_half():                             # test/runtests_solver.jl:15
    r = solve(workloads.delsarte(3, 10, Fraction(1, 2)))
    assert abs(r.p_obj - mpmath.mpf("13.158314")) < mpmath.mpf(10) ** -5


def test_two_radii_sphere_packing_near_cohn_elkies():      # test/runtests_solver.jl:19-22 (d = 15 there; the bound decreases with d)
    r = solve(workloads.sphere_packing(8, 7, [Fraction(1, 2), Fraction(1, 2)]), gap=1e-15)
    with mpmath.workprec(200):
        target = mpmath.pi ** 4 / 384
        assert target - mpmath.mpf(10) ** -12 < r.p_obj < target + mpmath.mpf(4) / 1000


def test_status_codes_and_iteration_limit():
    r = solvesdp(workloads.maxcut(workloads.laplacian_cycle(3)), lib="oracle", maxiterations=5)
    assert r.error_code == 2 and r.iterations == 5 and r.status != "Optimal"     # src/solver.jl:362-366
    r = solvesdp(workloads.maxcut(workloads.laplacian_cycle(3)), lib="oracle", need_primal_feasible=True)
    assert r.status in ("PrimalFeasible", "Feasible", "NearOptimal", "Optimal") and r.iterations < 30


def test_three_point_bound_n4_is_10():                       # test/runtests_solver.jl:26-27, 89-93
    sdp = workloads.three_point_bound(4, Fraction(1, 6), -1, 4)
    assert sdp.num_constraints == 50 and len(sdp.clusters[0].blocks) == 19      # P=50, 19 blocks, K=79
    r = solve(sdp, omega_p=10 ** 3, omega_d=10 ** 3)
    assert abs(r.p_obj - 10) < mpmath.mpf(10) ** -25


def test_lovasz_theta_of_the_five_cycle_is_sqrt5():          # test/moi_tests.jl:7-8 (example_theta_problem); odd cycles: n cos(pi/n) / (1 + cos(pi/n))
    with mpmath.workprec(400):
        r = solve(workloads.lovasz_theta_cycle(5))
        assert abs(r.p_obj - mpmath.sqrt(5)) < mpmath.mpf(10) ** -26
        r = solve(workloads.lovasz_theta_cycle(7))
        assert abs(r.p_obj - 7 * mpmath.cos(mpmath.pi / 7) / (1 + mpmath.cos(mpmath.pi / 7))) < mpmath.mpf(10) ** -26

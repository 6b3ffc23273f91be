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


def test_cohn_elkies_8_15_as_in_the_reference_test():
    """test/runtests_solver.jl:19-20: cohnelkies(8, 15, prec=256) is pi^4/384 to 1e-4 (default gap 1e-15); one dense 1 x 1 block,
    three rank-1 SOS blocks, two clusters coupled by 31 free variables."""
    r = solve(workloads.cohnelkies(8, 15), gap=1e-15)
    with mpmath.workprec(200):
        target = mpmath.pi ** 4 / 384
        assert 0 < r.p_obj - target < mpmath.mpf(10) ** -4
        assert abs(r.p_obj - target - mpmath.mpf("7.0919e-5")) < mpmath.mpf(10) ** -8     # the same degree-15 bound as the two-radii SDP below


def test_two_radii_sphere_packing_d15_prec300_as_in_the_reference_test():
    """test/runtests_solver.jl:21-22: Nsphere_packing(8, 15, [1//2, 1//2], 2, prec=300) is pi^4/384 to 1e-4 (default gap 1e-15)."""
    r = solve(workloads.sphere_packing(8, 15, [Fraction(1, 2), Fraction(1, 2)], prec=300), gap=1e-15)
    with mpmath.workprec(200):
        target = mpmath.pi ** 4 / 384
        assert 0 < r.p_obj - target < mpmath.mpf(10) ** -4
        assert abs(r.p_obj - target - mpmath.mpf("7.0919e-5")) < mpmath.mpf(10) ** -8     # the bound of this degree (device and oracle agree on it)


# The reference prints the solver log of min_f(2) in docs/src/solving.md:38-52: columns iter, time, mu, D-obj, P-obj, gap,
# D-error, d-error, p-error, alpha_d, alpha_p, beta for rows 1-3 and 55-56, "Optimal solution found" after 56 iterations,
# and the final objectives to 78 digits.  The step lengths depend on a Float64 Lanczos with a random start vector
# (src/solver.jl:1659), so the reference reproduces its own log to ~1e-5 in alpha at best; the printed 3-4 digits are compared.
REFERENCE_LOG_MIN_F_2 = {
    1: (1.000e+20, 0.000e+00, 0.000e+00, 0.00e+00, 1.00e+10, 1.00e+00, 1.95e+10, 7.42e-01, 7.10e-01, 3.00e-01),
    2: (3.995e+19, 1.999e+11, -2.907e+09, 1.03e+00, 2.58e+09, 2.58e-01, 5.65e+09, 7.46e-01, 7.17e-01, 3.00e-01),
    3: (1.576e+19, 3.079e+11, -4.779e+09, 1.03e+00, 6.53e+08, 6.53e-02, 1.60e+09, 7.32e-01, 7.31e-01, 3.00e-01),
    55: (5.066e-14, -2.113e+00, -2.113e+00, 8.39e-14, None, None, None, 1.00e+00, 1.00e+00, 1.00e-01),
    56: (5.067e-15, -2.113e+00, -2.113e+00, 8.39e-15, None, None, None, 1.00e+00, 1.00e+00, 1.00e-01),
}
REFERENCE_FINAL_MIN_F_2 = ("-2.112913881423601867325289796075301826150007716044362101360781221096092533872562",
                           "-2.112913881423605414349991239275382883067580432169230529548206052006356176913883",
                           "8.393680245626824434313082297089851809408852609517159688543365552836941907249006e-16")


def check_against_reference_log(r):
    """r: SolveResult of min_f(2) with the reference's default options (gap 1e-15)."""
    assert r.status == "Optimal" and r.iterations == 56                     # "Optimal solution found" after row 56
    for it, row in REFERENCE_LOG_MIN_F_2.items():
        h = r.history[it - 1]
        got = (h["mu"], h["d_obj"], h["p_obj"], h["gap"], h["err_P"], h["err_p"], h["err_d"], h["alpha_d"], h["alpha_p"], h["beta_c"])
        for k, (g, want) in enumerate(zip(got, row)):
            if want is None:                                                # errors at the 1e-77 noise floor of 256 bits
                assert abs(g) < 1e-70
            elif want == 0:
                assert g == 0
            else:
                assert abs(g - want) <= (1.5e-3 if k < 3 else 6e-3) * abs(want), (it, k, g, want)   # the printed digits: %.3e for mu and the objectives, %.2e for the rest
    with mpmath.workprec(300):
        d_ref, p_ref, g_ref = (mpmath.mpf(v) for v in REFERENCE_FINAL_MIN_F_2)
        # the final objectives are each within the final gap (8.4e-16) of the optimum; two runs whose step lengths differ by
        # 1e-5 agree much better than that: 1e-18 here
        assert abs(r.d_obj - d_ref) < mpmath.mpf(10) ** -18 and abs(r.p_obj - p_ref) < mpmath.mpf(10) ** -18
        assert abs(r.gap - g_ref) < mpmath.mpf(10) ** -3 * g_ref


def test_oracle_reproduces_the_reference_solver_log_of_min_f_2():
    """Pins the oracle's TRAJECTORY (not only its optimum) to the one solver log the reference ships."""
    sdp = workloads.min_f(2)
    assert sdp.num_constraints == 11 and [b.n for b in sdp.clusters[0].blocks] == [4, 3] and sdp.N == 1
    r = solvesdp(sdp, lib="oracle")
    check_against_reference_log(r)
    assert abs(r.p_obj - mpmath.mpf("-2.113")) < mpmath.mpf(10) ** -2        # test/runtests_solver.jl:10-11


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


def test_povm_two_states_is_half_plus_quarter_sqrt2():       # test/moi_tests.jl:9-10 (example_POVM, atol 1e-30): two dense blocks in one cluster
    sdp = workloads.povm_two_states()
    assert sdp.num_constraints == 16 and [b.n for b in sdp.clusters[0].blocks] == [4, 4]
    r = solve(sdp)
    with mpmath.workprec(300):
        target = mpmath.mpf(1) / 2 + mpmath.sqrt(2) / 4
        assert abs(r.p_obj - target) < mpmath.mpf(10) ** -30 and abs(r.d_obj - target) < mpmath.mpf(10) ** -30

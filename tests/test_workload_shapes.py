"""The generators of the five BASELINE configurations produce the SDP shapes derived from the reference's example code
(clusters J, constraints P, free variables N, blocks, K = sum of block sizes)."""
from fractions import Fraction as F

import clrs_b200
from clrs_b200 import workloads


def shape(sdp):
    blocks = [b.n for c in sdp.clusters for b in c.blocks]
    return dict(J=len(sdp.clusters), P=sdp.num_constraints, N=sdp.N, blocks=len(blocks), K=sum(blocks), max_n=max(blocks))


def test_config1_polyopt_d20():            # examples/PolyOpt.jl:7-30
    s = shape(workloads.polyopt_random(20, 0))
    assert (s["J"], s["P"], s["N"], s["K"]) == (1, 41, 1, 21)


def test_config2_maxcut():                 # README.md:39-65 (small instance of the n = 300 shape: one dense block, P = n, N = 0)
    s = shape(workloads.maxcut(workloads.laplacian_random(12, 0.5, 0)))
    assert s == dict(J=1, P=12, N=0, blocks=1, K=12, max_n=12)


def test_config3_delsarte():               # examples/Delsarte.jl:7-49
    s = shape(workloads.delsarte(8, 16, F(1, 2)))
    assert (s["J"], s["P"], s["N"], s["blocks"], s["K"]) == (1, 34, 1, 35, 66)
    s = shape(workloads.delsarte(8, 32, F(1, 2)))
    assert (s["P"], s["blocks"], s["K"], s["max_n"]) == (66, 67, 130, 33)


def test_config4_three_point_bound():      # examples/ThreePointBound.jl:45-169
    s = shape(workloads.three_point_bound(4, F(1, 6), -1, 4))
    assert (s["J"], s["P"], s["blocks"], s["K"]) == (1, 50, 19, 79)
    s = shape(workloads.three_point_bound(4, F(1, 6), 10, 10))
    assert (s["J"], s["P"], s["blocks"], s["K"], s["max_n"]) == (1, 379, 49, 751, 94)


def test_config5_sphere_packing():         # examples/SpherePacking.jl:13-115: J = 2 + T + N_r, N = T (2d + 2) + 1
    s = shape(workloads.sphere_packing(8, 15, [F(1, 2), F(1, 2)]))
    assert (s["J"], s["N"], s["P"]) == (7, 97, 197)
    s = shape(workloads.sphere_packing(8, 9, [F(1, 2), F(1, 2), F(3, 4), F(1)]))
    T, d, Nr = 10, 9, 4
    assert (s["J"], s["N"]) == (2 + T + Nr, T * (2 * d + 2) + 1)

"""SDPA-sparse reader (mirror of src/SDPAtoCLRS.jl:3-83) and the triplet upload path, on the CPU oracle."""
import mpmath
import numpy as np
import pytest

from clrs_b200 import sdpa, solvesdp, wire, workloads


def test_parse_skips_comments_and_reads_blocks():
    txt = '"a comment"\n* another\n2\n2\n{2, -3}\n1.5 2\n0 1 1 1 1\n0 1 1 2 -0.5\n1 1 2 2 3\n1 2 2 2 4\n2 2 3 3 1e-1\n2 1 1 2 0\n'
    m, bs, c, ent = sdpa.parse_sdpa_sparse(txt)
    assert (m, bs, c) == (2, [2, -3], ["1.5", "2"])
    assert ent[(0, (1,))] == {(1, 1): "1", (1, 2): "-0.5"}
    assert ent[(1, (2, 2))] == {(1, 1): "4"} and ent[(2, (2, 3))] == {(1, 1): "1e-1"}


def test_structure_follows_the_reference_conversion():
    # two independent groups of variables -> two clusters; the all-zero matrix of constraint 3 is dropped with its constraint
    txt = sdpa.write_sdpa_sparse([2, -2, 3], ["1", "2", "0", "5"],
                                 [(0, 1, 1, 1, 1), (0, 3, 1, 2, "0.5"), (1, 1, 1, 2, 1), (1, 2, 1, 1, 2), (2, 3, 1, 1, 1), (2, 3, 2, 3, -1),
                                  (3, 3, 1, 1, 0), (4, 2, 1, 1, 1), (4, 1, 2, 2, 1)])
    s = sdpa.sdpa_sparse_to_sdp(txt, prec=128)
    assert s.maximize and s.N == 0 and len(s.clusters) == 2
    c0, c1 = s.clusters
    assert c0.P == 2 and [b.n for b in c0.blocks] == [2, 1]            # constraints 1 and 4 share block 1 and the (2,1) diagonal entry
    assert c1.P == 1 and [b.n for b in c1.blocks] == [3]
    assert [float(v) for v in wire.from_wire(c0.c, 128)] == [1.0, 5.0] and float(wire.from_wire(c1.c, 128)[0]) == 2.0
    rows, cols, vals, mirror = c1.blocks[0].sparse[0]
    assert sorted(zip(rows.tolist(), cols.tolist())) == [(0, 0), (1, 2)] and mirror
    assert float(wire.from_wire(c1.blocks[0].C, 128)[1, 0]) == 0.5        # objective matrices are mirrored too
    assert set(c0.blocks[0].sparse) == {0, 1} and set(c0.blocks[1].sparse) == {0, 1}


def test_objective_only_block_is_rejected():
    with pytest.raises(ValueError):
        sdpa.sdpa_sparse_to_sdp(sdpa.write_sdpa_sparse([1, 1], ["1"], [(0, 2, 1, 1, 1), (1, 1, 1, 1, 1)]))


def test_maxcut_from_sdpa_matches_the_dense_generator_on_the_oracle():
    """Same SDP, once as dense E_ii matrices (README.md:39-65), once read from SDPA text as triplets: identical trajectory."""
    L = workloads.laplacian_cycle(5)
    a = solvesdp(workloads.maxcut(L), lib="oracle", duality_gap_threshold=1e-30)
    b = solvesdp(sdpa.sdpa_sparse_to_sdp(sdpa.maxcut_sdpa_text(L)), lib="oracle", duality_gap_threshold=1e-30)
    assert a.status == b.status == "Optimal" and a.iterations == b.iterations
    assert a.p_obj == b.p_obj and a.d_obj == b.d_obj
    with mpmath.workprec(300):
        assert abs(a.p_obj - mpmath.mpf(5) / 4 * (2 + 2 * mpmath.cos(mpmath.pi / 5))) < mpmath.mpf(10) ** -28

"""Device side of SURVEY.md §8(f)3: clrs_mp_qr_pivot against the oracle's QR, and the dependent-constraint detection of preprocess!
driven by the device (same cases as tests/test_preprocess.py)."""
import mpmath
import numpy as np
import pytest

from clrs_b200 import Cluster, Solver, solvesdp, wire, workloads
from clrs_b200 import preprocess as pp
from test_preprocess import check_qr, qr_case, toy, w

pytestmark = pytest.mark.gpu
PREC = 256


@pytest.fixture(scope="module")
def dev():
    s = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="device")
    yield s
    s.close()


@pytest.fixture(scope="module")
def ora():
    s = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="oracle")
    yield s
    s.close()


@pytest.mark.parametrize("seed,m,n,rank", [(0, 14, 8, 6), (3, 5, 9, 5), (7, 300, 40, 38), (1, 1, 1, 1)])
def test_device_qr_matches_the_oracle(dev, ora, seed, m, n, rank):
    A = qr_case(seed, m, n, spread=1000 if m > 100 else 5) if n > 2 else np.array([[3]], dtype=object)      # (wide entries: no exact ties between column norms)
    Rd, pd = check_qr(dev, A, rank)
    Ro, po = check_qr(ora, A, rank)
    # Both factorisations satisfy R^T R = (A P)^T (A P) (check_qr).  Their pivot orders may differ where the choice is a mathematical tie:
    # in a group of columns with one dependency (c = a + b + ...) the last two candidates have residuals r and -r, so rounding decides
    # which of them becomes the last independent pivot.  What must agree is the pivot order before any tie and the diagonal of R.
    with mpmath.workprec(PREC + 64):
        a, b = wire.from_wire(Rd, PREC), wire.from_wire(Ro, PREC)
        for i in range(rank):
            assert abs(a[i, i] - b[i, i]) <= abs(b[i, i]) * mpmath.mpf(2) ** -200, i
        same = 0
        while same < rank and pd[same] == po[same]:
            same += 1
        assert same >= rank - 2                                              # (two dependent groups in qr_case: at most their last pivots differ)
        scale = max(abs(v) for v in b.reshape(-1))
        assert max([abs(a[i, j] - b[i, j]) for i in range(same) for j in range(same)] + [0]) <= scale * mpmath.mpf(2) ** -230


def test_device_qr_at_512_bit():
    s = Solver(workloads.maxcut(workloads.laplacian_cycle(3), prec=512), lib="device")
    A = qr_case(5, 40, 12)
    R, perm = s.mp_qr_pivot(wire.to_wire(A.tolist(), 512))
    s.close()
    with mpmath.workprec(600):
        Rm = wire.from_wire(R, 512)
        AP = mpmath.matrix(A[:, perm].tolist()); G = AP.T * AP; RR = mpmath.matrix(Rm.tolist()); E = RR.T * RR - G
        assert max(abs(E[i, j]) for i in range(12) for j in range(12)) < max(abs(G[i, i]) for i in range(12)) * mpmath.mpf(2) ** -490
        assert sum(1 for i in range(12) if abs(Rm[i, i]) > mpmath.mpf(2) ** -250) == 10


def test_preprocess_on_the_device_removes_the_same_constraints_as_on_the_oracle(dev, ora):
    n = 6
    sdp = workloads.maxcut(workloads.laplacian_cycle(n))
    blk = sdp.clusters[0].blocks[0]
    blk.dense = {p: np.asarray(A) for p, A in blk.dense.items()}
    blk.dense[n] = wire.wire_eye_scaled(n, 1, PREC)
    sdp.clusters[0] = Cluster(B=wire.wire_zeros((n + 1, 0), PREC), c=w([1] * n + [n]), blocks=[blk])
    new, cs, _ = pp.preprocess(sdp, dev)
    assert len(cs) == 1 and new.clusters[0].P == n
    assert len(pp.preprocess(sdp, ora)[1]) == 1
    r = solvesdp(new, lib="device", duality_gap_threshold=1e-30)
    with mpmath.workprec(300):
        assert r.status == "Optimal" and abs(r.p_obj - mpmath.mpf(n) / 4 * 4) < mpmath.mpf(10) ** -28        # even cycle: bipartite, the cut takes all n edges
    assert len(pp.preprocess(toy([1, 1]), dev)[1]) == 1
    with pytest.raises(ValueError):
        pp.preprocess(toy([1, 0]), dev)

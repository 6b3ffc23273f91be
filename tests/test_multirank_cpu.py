"""world_size-2 gloo test of the host-side logic of the sharded path: every rank must
derive the same cluster -> rank partition from the SDP alone, the owners must cover all clusters, and
the 128-byte communicator id created on rank 0 must reach every rank unchanged."""
import os
import sys
from fractions import Fraction

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cluster_weights(sdp):
    w = []
    for cl in sdp.clusters:
        v = float(cl.P) ** 3
        for b in cl.blocks:
            v += float(b.n) ** 3 * (2.0 * len(b.dense) + 15 if b.high_rank else 15)
        w.append(v)
    return w


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import clrs_b200
    from clrs_b200 import workloads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sdp = workloads.sphere_packing(8, 5, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4)])
    owners = clrs_b200.partition_clusters(cluster_weights(sdp), world)
    gathered = [None] * world
    dist.all_gather_object(gathered, owners)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8) * 3 + 1
    dist.broadcast(uid, 0)
    q.put((rank, owners, gathered, uid.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_on_partition_and_communicator_id():
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owners = res[0][1]
    for rank, own, gathered, uid in res:
        assert own == owners and all(g == owners for g in gathered)
        assert uid == [(3 * i + 1) % 256 for i in range(128)]
    assert set(owners) == {0, 1}                       # both ranks get work
    assert len(owners) == 11                           # 2 + T + N_r clusters with T = 6, N_r = 3


def test_partition_balances_by_weight():
    import clrs_b200
    own = clrs_b200.partition_clusters([100, 1, 1, 50, 50, 3], 2)
    load = [sum(w for w, o in zip([100, 1, 1, 50, 50, 3], own) if o == r) for r in range(2)]
    assert abs(load[0] - load[1]) <= 3
    assert clrs_b200.partition_clusters([5, 4, 3], 1) == [0, 0, 0]
    assert sorted(clrs_b200.partition_clusters([1] * 8, 8)) == list(range(8))


# ---- the cross-rank sum of multi-limb numbers: exponent max + int64-lane sum (csrc/lanes.cuh), run here with gloo ----
def _lane_worker(rank, world, port, q):
    import ctypes as C
    import numpy as np
    import mpmath
    sys.path.insert(0, ROOT)
    import clrs_b200
    from clrs_b200 import wire
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hc = C.CDLL(os.path.join(ROOT, "clusteredlowranksolver.jl_b200", "csrc", "libclrs_hostcheck.so"))
    n, prec = 64, 256
    rng = np.random.default_rng(100 + rank)
    with mpmath.workprec(prec):
        vals = []
        for i in range(n):
            v = mpmath.mpf(int(rng.integers(1, 2 ** 62))) / 2 ** 62 + mpmath.mpf(int(rng.integers(0, 2 ** 62))) / mpmath.mpf(2) ** 200
            v *= mpmath.mpf(2) ** int(rng.integers(-300, 300)) * (-1 if rng.integers(0, 2) else 1)
            if i % 9 == rank:
                v = mpmath.mpf(0)                      # zeros on some ranks
            if i == 5:
                v = mpmath.mpf(1) if rank == 0 else -mpmath.mpf(1)   # exact cancellation (two ranks)
            if i == 6:
                v = mpmath.mpf(3) * mpmath.mpf(2) ** (40 * rank)      # very different magnitudes
            vals.append(v)
        w = wire.to_wire(vals, prec)
    E = np.zeros(n, dtype=np.int32)
    hc.hc_lane_exp(C.c_int(n), w.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p))
    Et = torch.from_numpy(E)
    dist.all_reduce(Et, op=dist.ReduceOp.MAX)                       # step (1): ncclMax on int32 in the library
    lanes = np.zeros((n, 9), dtype=np.int64)
    hc.hc_to_lanes(C.c_int(n), w.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p), lanes.ctypes.data_as(C.c_void_p))
    Lt = torch.from_numpy(lanes)
    dist.all_reduce(Lt, op=dist.ReduceOp.SUM)                       # step (3): ncclSum on int64 in the library
    out = wire.wire_zeros((n,), prec)
    hc.hc_from_lanes(C.c_int(n), lanes.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    gathered = [None] * world
    dist.all_gather_object(gathered, [(int(v._mpf_[0]), int(v._mpf_[1]), int(v._mpf_[2])) for v in vals])     # exact (sign, mantissa, exponent)
    q.put((rank, out.tobytes(), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_sum_multi_limb_numbers_through_int64_lanes():
    import mpmath
    import numpy as np
    sys.path.insert(0, ROOT)
    import clrs_b200
    from clrs_b200 import wire
    world, port = 2, 29541
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lane_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                                   # every rank holds the same bits
    out = np.frombuffer(res[0][1], dtype=wire.wire_dtype(256))
    with mpmath.workprec(400):
        got = wire.from_wire(out, 256)
        per_rank = [[(-1) ** sg * mpmath.mpf(man) * mpmath.mpf(2) ** ex for (sg, man, ex) in g] for g in res[0][2]]
        for i in range(len(got)):
            exact = sum(pr[i] for pr in per_rank)
            big = max(abs(pr[i]) for pr in per_rank)
            # one 32-bit guard lane below the aligned 256-bit mantissas, then the sum is truncated to 256 bits
            assert abs(got[i] - exact) <= big * mpmath.mpf(2) ** -280 + abs(exact) * mpmath.mpf(2) ** -255, i
        assert got[5] == 0 and got[6] == 3 + 3 * mpmath.mpf(2) ** 40


def test_shard_plan_splits_single_heavy_clusters_by_blocks():
    """config 4 is ONE cluster -> its blocks are spread over the ranks; §8(e)(ii)/(iii): sphere packing keeps whole
    clusters (the big SOS2 cluster takes the column-split path), single-block SDPs are never split."""
    sys.path.insert(0, ROOT)
    import clrs_b200
    from clrs_b200 import workloads, plan_shards
    tp = workloads.three_point_bound(4, Fraction(1, 6), 4, 4)
    own, split, bown = plan_shards(tp, 4)
    assert split == [True] and sorted(set(bown[0])) == [0, 1, 2, 3] and own[0] == bown[0][0]
    w = [float(b.n) ** 3 * (2.0 * len(b.dense) + 15 if b.high_rank else 15) for b in tp.clusters[0].blocks]      # the library's block weights
    load = [sum(wi for wi, o in zip(w, bown[0]) if o == r) for r in range(4)]
    assert max(load) <= sum(w) / 4 + max(w)                              # greedy LPT: no rank exceeds the fair share by more than one item
    assert plan_shards(tp, 1)[1] == [False] and plan_shards(tp, 4, split_mode=0)[1] == [False]
    sp = workloads.sphere_packing(8, 5, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4)])
    own, split, bown = plan_shards(sp, 2, big_cluster=20)
    assert not any(s and c.P >= 20 for s, c in zip(split, sp.clusters))  # column-split clusters keep their blocks together
    for j, c in enumerate(sp.clusters):
        if not split[j]:
            assert all(o == own[j] for o in bown[j])
    assert plan_shards(workloads.maxcut(workloads.laplacian_cycle(5)), 8)[1] == [False]
    assert plan_shards(tp, 4) == plan_shards(tp, 4)                      # deterministic

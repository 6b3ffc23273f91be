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

"""Multi-GPU parity of the cluster-sharded path.  Launch with
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_multi.py
Every rank uploads the same SDP; clusters are partitioned over the ranks inside the library; Q, sum_j u_j,
p and the scalars travel over NCCL.  Rank 0 also solves the SDP alone and the objectives are compared."""
import os, sys, time
from fractions import Fraction
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, mpmath
import clrs_b200
from clrs_b200 import workloads, solvesdp, nccl_unique_id

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def shared_uid():
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())

ok = True
cases = [("sphere_packing(8,7,2 radii): 7 clusters, 50 free variables", lambda: workloads.sphere_packing(8, 7, [Fraction(1, 2), Fraction(1, 2)]), 1e-20),
         ("polyopt d=20: one cluster (rank 1 owns nothing)", lambda: workloads.polyopt_random(20, 0), 1e-30),
         ("sphere_packing(8,9,3 radii): 11 clusters", lambda: workloads.sphere_packing(8, 9, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4)]), 1e-15),
         # single clusters with many blocks: the BLOCKS are spread over the ranks, S_j is all-reduced
         ("three_point_bound(4,1/6,4,4): one cluster, 28 blocks, split by blocks", lambda: workloads.three_point_bound(4, Fraction(1, 6), 4, 4), 1e-30),
         ("delsarte(8,8): one cluster with a free variable, 19 blocks, split by blocks", lambda: workloads.delsarte(8, 8, Fraction(1, 2)), 1e-30),
         ("theta(C_9) + maxcut: dense single block (not split)", lambda: workloads.lovasz_theta_cycle(9), 1e-30)]
if os.environ.get("CLRS_MULTI_CASES"):
    keep = [int(t) for t in os.environ["CLRS_MULTI_CASES"].split(",")]
    cases = [c for i, c in enumerate(cases) if i in keep]
for name, make, gap in cases:
    sdp = make()
    uid = shared_uid()
    t0 = time.time()
    multi = solvesdp(sdp, lib="device", device=local, duality_gap_threshold=gap, comm=(rank, world, uid), keep_solver=True)
    tm = time.time() - t0
    owners = sorted({multi.solver.block_owner(j, l) for j, c in enumerate(sdp.clusters) for l in range(len(c.blocks))})
    multi.solver.close()
    if "split by blocks" in name and len(owners) < min(world, 2):
        ok = False
        print(f"[FAIL] {name}: blocks were not spread over the ranks (owners {owners})", flush=True)
    dist.barrier()
    if rank == 0:
        single = solvesdp(sdp, lib="device", device=local, duality_gap_threshold=gap)
        with mpmath.workprec(400):
            rel = abs(multi.p_obj - single.p_obj) / max(1, abs(single.p_obj))
            reld = abs(multi.d_obj - single.d_obj) / max(1, abs(single.d_obj))
        good = multi.status == single.status == "Optimal" and rel < mpmath.mpf(10) ** -25 and reld < mpmath.mpf(10) ** -25 and abs(multi.iterations - single.iterations) <= 1
        ok = ok and good
        print(f"[{'OK' if good else 'FAIL'}] {name}: {world} ranks {multi.iterations} it in {tm:.2f}s vs 1 rank {single.iterations} it in {single.time:.2f}s; "
              f"rel diff p_obj {float(rel):.2e} d_obj {float(reld):.2e}; {multi}", flush=True)
    dist.barrier()
if rank == 0:
    print("MULTI-GPU PARITY", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)

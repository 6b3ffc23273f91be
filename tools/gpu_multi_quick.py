"""Quick 2-rank run of the sharded path (debug aid): a few iterations of a small sphere-packing SDP."""
import os, sys, time
from fractions import Fraction
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import clrs_b200
from clrs_b200 import workloads, Solver, nccl_unique_id
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    t.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(t, 0)
uid = bytes(t.cpu().tolist())
sdp = workloads.sphere_packing(8, 7, [Fraction(1, 2), Fraction(1, 2)])
S = Solver(sdp, lib="device", device=local, comm=(rank, world, uid), duality_gap_threshold=1e-30)
print(f"rank {rank}: solver up", flush=True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    info = S.iterate()
    print(f"rank {rank}: iter {i} stop {info.stop} mu {info.mu:.3e} ms {S.last_iteration_ms():.2f}", flush=True)
torch.cuda.synchronize()
dist.barrier()
S.close()
print(f"rank {rank}: DONE", flush=True)
dist.destroy_process_group()

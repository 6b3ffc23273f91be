"""Small runs for compute-sanitizer (racecheck / memcheck): a Cholesky, a dense-path solve and a low-rank solve, few iterations."""
import sys
from fractions import Fraction as F
sys.path.insert(0, ".")
import clrs_b200
from clrs_b200 import workloads, Solver
for sdp in (workloads.maxcut(workloads.laplacian_cycle(5)), workloads.sphere_packing(8, 3, [F(1, 2), F(1, 2)]), workloads.maxcut(workloads.laplacian_complete(40))):
    S = Solver(sdp, lib="device")
    for _ in range(2):
        info = S.iterate()
    print(sdp.describe(), info.stop, info.mu, flush=True)
    S.close()
print("done")

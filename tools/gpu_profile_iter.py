"""Two IPM iterations of the BASELINE workload, for `ncu` launch lists (not a test)."""
import sys
sys.path.insert(0, ".")
import clrs_b200
from clrs_b200 import workloads, Solver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
S = Solver(workloads.maxcut(workloads.laplacian_random(n, 0.5, 0)), lib="device")
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    info = S.iterate()
print("ms/iter", S.last_iteration_ms(), list(info.phase_ms)[:12])

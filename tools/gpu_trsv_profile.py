"""A few iterations of a sphere-packing SDP at 512 bit for ncu captures of the substitution kernels (not a test)."""
import sys
from fractions import Fraction as F
sys.path.insert(0, ".")
import clrs_b200
from clrs_b200 import workloads, Solver
S = Solver(workloads.sphere_packing(8, 31, [F(1, 2), F(1, 2)], prec=512), lib="device")
S.use_graph(False)
for _ in range(2):
    info = S.iterate()
print("ms/iter", S.last_iteration_ms())

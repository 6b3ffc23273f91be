"""Cholesky-only run for ncu launch lists (not a test)."""
import sys
sys.path.insert(0, ".")
import numpy as np, mpmath
import clrs_b200
from clrs_b200 import workloads, Solver, wire
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
S = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="device")
rng = np.random.default_rng(0)
G = rng.standard_normal((n, n)); A = G @ G.T + n * np.eye(n)
Aw = wire.wire_zeros((n, n), 256)
with mpmath.workprec(300):
    flat = wire.to_wire([[mpmath.mpf(float(v)) for v in row] for row in A], 256)
for _ in range(2):
    L = S.mp_cholesky(flat)
print("ok")

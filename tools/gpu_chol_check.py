"""Cholesky through the C ABI for a few sizes: prints a hash of the factor's bytes and the wall time per call, so that two builds /
settings (CLRS_CHOL_LOOKAHEAD=0/1, CLRS_MP_CHOL_NOINV=0/1, CLRS_CHOL_PANEL) can be compared bit for bit.  Not a test."""
import hashlib, sys, time
sys.path.insert(0, ".")
import numpy as np, mpmath
import clrs_b200
from clrs_b200 import workloads, Solver, wire
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = Solver(workloads.maxcut(workloads.laplacian_cycle(3), prec=prec), lib="device")
for n in [int(a) for a in sys.argv[2:]] or [100, 300, 500]:
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n)); A = G @ G.T + n * np.eye(n)
    with mpmath.workprec(prec + 64):
        flat = wire.to_wire([[mpmath.mpf(float(v)) for v in row] for row in A], prec)
    L = S.mp_cholesky(flat)
    t0 = time.perf_counter(); L = S.mp_cholesky(flat); dt = time.perf_counter() - t0
    print(f"n={n} prec={prec} sha={hashlib.sha256(L.tobytes()).hexdigest()[:16]} wall_ms={1e3 * dt:.2f}", flush=True)

import sys
sys.path.insert(0, ".")
import numpy as np, mpmath
import clrs_b200
from clrs_b200 import workloads, Solver, wire
PREC = 300
def rand_wire(seed, shape, spread):
    rng = np.random.default_rng(seed)
    out = wire.wire_zeros(shape, PREC)
    limbs = rng.integers(0, 2 ** 63, size=shape + (5,), dtype=np.uint64) * np.uint64(2)
    limbs[..., 0] &= np.uint64(0xFFFFFFFFFFF00000)      # prec 300: low 20 bits of the lowest limb are zero
    limbs[..., 4] |= np.uint64(1) << np.uint64(63)
    out["limb"] = limbs; out["exp"] = rng.integers(-spread, spread + 1, size=shape); out["sign"] = rng.choice([-1, 1], size=shape)
    return out
d = Solver(workloads.maxcut(workloads.laplacian_cycle(3), prec=PREC), lib="device")
o = Solver(workloads.maxcut(workloads.laplacian_cycle(3), prec=PREC), lib="oracle")
for (M, N, K) in [(20, 24, 50), (200, 193, 192), (193, 193, 64), (300, 112, 128)]:
    A = rand_wire(1, (M, K), 5); B = rand_wire(2, (K, N), 5)
    C1, _ = d.mp_gemm(A, B, path=1); C2, _ = d.mp_gemm(A, B, path=2); Co, _ = o.mp_gemm(A, B)
    with mpmath.workprec(500):
        a, b, c = wire.from_wire(C1, PREC).reshape(-1), wire.from_wire(C2, PREC).reshape(-1), wire.from_wire(Co, PREC).reshape(-1)
        sc = max(abs(v) for v in c)
        e1 = max(abs(x - z) for x, z in zip(a, c)) / sc; e2 = max(abs(x - z) for x, z in zip(b, c)) / sc
    print(M, N, K, "dp4a vs oracle 2^%.1f" % float(mpmath.log(e1, 2) if e1 else -999), "tc vs oracle 2^%.1f" % float(mpmath.log(e2, 2) if e2 else -999), "paths equal:", C1.tobytes() == C2.tobytes(), flush=True)
n = 100
rng = np.random.default_rng(0); G = rng.standard_normal((n, n)); Am = G @ G.T + n * np.eye(n)
with mpmath.workprec(400):
    Aw = wire.to_wire([[mpmath.mpf(float(v)) for v in row] for row in Am], PREC)
L1 = d.mp_cholesky(Aw); L2 = o.mp_cholesky(Aw)
with mpmath.workprec(500):
    a, b = wire.from_wire(L1, PREC).reshape(-1), wire.from_wire(L2, PREC).reshape(-1)
    print("chol100 err 2^%.1f" % float(mpmath.log(max(abs(x - y) for x, y in zip(a, b)) / max(abs(v) for v in b), 2)))

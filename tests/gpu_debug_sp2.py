import sys
sys.path.insert(0, ".")
from fractions import Fraction
import numpy as np, mpmath
import clrs_b200
from clrs_b200 import workloads, Solver, wire
prec = 256
sdp = workloads.sphere_packing(8, 23, [Fraction(1, 2), Fraction(1, 2)], prec=prec)
blk = sdp.clusters[1].blocks[1]          # SOS22: lambda = x, basis vectors
seen, vecs = set(), []
for t in blk.lowrank:
    key = t.vs.tobytes()
    if key not in seen:
        seen.add(key); vecs.append(t.vs[0])
W = np.stack(vecs)                        # u x delta
V = np.ascontiguousarray(W.T)
print("W", W.shape)
d = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="device"); o = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="oracle")
C1, _ = d.mp_gemm(np.ascontiguousarray(W), V, path=1); Co, _ = o.mp_gemm(np.ascontiguousarray(W), V)
with mpmath.workprec(600):
    w = wire.from_wire(W, prec); a = wire.from_wire(C1, prec); c = wire.from_wire(Co, prec)
    ex = [[int(mpmath.log(abs(v), 2)) if v != 0 else -9999 for v in row] for row in w]
    print("exponent range of W rows (min,max over k):", [(min(r), max(r)) for r in ex[:3]], "...", [(min(r), max(r)) for r in ex[-2:]])
    worst_rel, worst_norm = 0, 0
    sc = max(abs(v) for v in c.reshape(-1))
    for i in range(c.shape[0]):
        for j in range(c.shape[1]):
            e = abs(a[i, j] - c[i, j])
            if c[i, j] != 0: worst_rel = max(worst_rel, e / abs(c[i, j]))
            worst_norm = max(worst_norm, e / sc)
    print("W*W^T device vs oracle: worst componentwise rel err 2^%.1f, worst normwise 2^%.1f" % (float(mpmath.log(worst_rel, 2)), float(mpmath.log(worst_norm, 2))))
    cm = [[int(mpmath.log(abs(v), 2)) for v in row] for row in c]
    print("log2|C| corners:", cm[0][0], cm[0][-1], cm[-1][-1])

import sys
sys.path.insert(0, ".")
from fractions import Fraction
import mpmath
import clrs_b200
from clrs_b200 import workloads, Solver, wire
for (d, prec) in [(15, 300), (23, 256), (23, 300), (31, 256)]:
    sdp = workloads.sphere_packing(8, d, [Fraction(1, 2), Fraction(1, 2)], prec=prec)
    S = Solver(sdp, lib="device"); O = Solver(sdp, lib="oracle")
    try:
        a = S.iterate(); msg = f"ok mu={a.mu:.3e} alpha={a.alpha_d:.4f},{a.alpha_p:.4f}"
    except Exception as e:
        msg = "FAIL " + str(e)[:60]
    b = O.iterate()
    print(f"d={d} prec={prec} N={sdp.N}: device {msg} | oracle alpha={b.alpha_d:.4f},{b.alpha_p:.4f}", flush=True)
    # compare S of the big cluster and LinvB, Q between device and oracle
    with mpmath.workprec(400):
        for what, j in (("S", 1), ("LinvB", 1), ("Q", 0)):
            x = wire.from_wire(S.debug_get(what, j, 0), prec); y = wire.from_wire(O.debug_get(what, j, 0), prec)
            sc = max(abs(v) for v in y); err = max(abs(p - q) for p, q in zip(x, y)) / sc
            print("   ", what, "rel err 2^%.1f" % float(mpmath.log(err, 2) if err else -999), flush=True)

"""Ad-hoc GPU timing of a few workloads (not a test)."""
import sys, time
from fractions import Fraction
sys.path.insert(0, ".")
import clrs_b200
import oracle.binding  # noqa: F401  (registers lib="oracle")
from clrs_b200 import workloads, solvesdp, PHASES
import numpy as np

def run(name, sdp, iters=None, **kw):
    t = time.time()
    r = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30, maxiterations=iters or 500, **kw)
    dt = time.time() - t
    ph = np.array([h["phase_ms"] for h in r.history[1:]] or [[0.0] * 17])
    print(f"{name}: {sdp.describe()}\n   {r} {getattr(r, 'failure', '')}\n   wall {dt:.2f}s  {r.iterations/dt:.1f} it/s  phases(ms/it): " +
          " ".join(f"{PHASES[i]}={ph[:, i].mean():.2f}" for i in range(12)), flush=True)

which = sys.argv[1:] or ["poly", "del", "sp", "mc40"]
if "poly" in which: run("polyopt20", workloads.polyopt_random(20))
if "del" in which: run("delsarte16", workloads.delsarte(8, 16, Fraction(1, 2)))
if "sp" in which: run("sphere(2,15)", workloads.sphere_packing(8, 15, [Fraction(1, 2), Fraction(1, 2)]))
if "mc40" in which: run("maxcut40", workloads.maxcut(workloads.laplacian_random(40)))
if "mc100" in which: run("maxcut100", workloads.maxcut(workloads.laplacian_random(100)), iters=5)
if "mc300" in which: run("maxcut300", workloads.maxcut(workloads.laplacian_random(300)), iters=3)
if "tp6" in which: run("threepoint(4,6,6)", workloads.three_point_bound(4, Fraction(1, 6), 6, 6), omega_p=10 ** 3, omega_d=10 ** 3)
if "tp10" in which: run("threepoint(4,10,10)", workloads.three_point_bound(4, Fraction(1, 6), 10, 10), omega_p=10 ** 3, omega_d=10 ** 3)
if "sp31" in which: run("sphere(2,31) prec300", workloads.sphere_packing(8, 31, [Fraction(1, 2), Fraction(1, 2)], prec=300))
if "sp431" in which: run("sphere(4,31) prec300", workloads.sphere_packing(8, 31, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4), Fraction(1)], prec=300), iters=12)
if "sp512" in which:
    for d in (15, 23, 31):
        run(f"sphere(2,{d}) prec512", workloads.sphere_packing(8, d, [Fraction(1, 2), Fraction(1, 2)], prec=512))
if "sp4512" in which: run("sphere(4,31) prec512", workloads.sphere_packing(8, 31, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4), Fraction(1)], prec=512), iters=12)
if "gemm512" in which:
    from clrs_b200 import Solver
    import mpmath
    rng = np.random.default_rng(1)
    for prec in (300, 512):
        with mpmath.workprec(prec + 40):
            M, K, N = 70, 200, 50
            A = [mpmath.mpf(float(x)) * mpmath.mpf(2) ** int(e) / 3 for x, e in zip(rng.standard_normal(M * K), rng.integers(-30, 30, M * K))]
            B = [mpmath.mpf(float(x)) / 7 for x in rng.standard_normal(K * N)]
            dev = clrs_b200.mp_gemm(A, B, M, K, N, prec=prec, lib="device"); ora = clrs_b200.mp_gemm(A, B, M, K, N, prec=prec, lib="oracle")
            err = max(abs(a - b) for a, b in zip(dev, ora)) / max(abs(b) for b in ora)
            print(f"gemm prec {prec}: rel diff 2^{float(mpmath.log(err, 2)) if err else -9999:.1f}")
if "sp4" in which: run("sphere(4,23) prec512", workloads.sphere_packing(8, 23, [Fraction(1, 2), Fraction(1, 2), Fraction(3, 4), Fraction(1)], prec=512), iters=6)

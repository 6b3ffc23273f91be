"""Ad-hoc GPU timing of a few workloads (not a test)."""
import sys, time
from fractions import Fraction
sys.path.insert(0, ".")
import clrs_b200
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

"""Device history on the ill-conditioned sphere-packing instance (not a test)."""
import sys
sys.path.insert(0, ".")
from fractions import Fraction
import clrs_b200
from clrs_b200 import workloads, solvesdp, Solver
d = int(sys.argv[1]) if len(sys.argv) > 1 else 31
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sdp = workloads.sphere_packing(8, d, [Fraction(1, 2), Fraction(1, 2)], prec=prec)
if len(sys.argv) > 3:           # profile mode: a few iterations only
    S = Solver(sdp, lib="device")
    for _ in range(int(sys.argv[3])): info = S.iterate()
    print("ms/iter", S.last_iteration_ms()); sys.exit(0)
r = solvesdp(sdp, lib="device", duality_gap_threshold=1e-30, maxiterations=140)
print(r, getattr(r, "failure", ""))
for h in r.history[::6]: print(h["iter"], "mu %.2e gap %.2e P %.2e p %.2e d %.2e a %.4f %.4f" % (h["mu"], h["gap"], h["err_P"], h["err_p"], h["err_d"], h["alpha_p"], h["alpha_d"]))

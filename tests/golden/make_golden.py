"""Generate the golden end-to-end values under tests/golden/ with the CPU oracle.

Run from the repository root:  python tests/golden/make_golden.py [name ...]
The oracle (oracle/clrs_oracle.cpp) restates the reference's iteration on MPFR; its
zero-skipping mode is bit-identical to the plain dense path (tests/test_oracle.py) and is
what makes the n = 300 MAX-CUT instance affordable on a CPU.
"""
import json
import os
import sys
import time
from fractions import Fraction

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mpmath
import clrs_b200
import oracle.binding  # noqa: F401  (registers lib="oracle")
from clrs_b200 import workloads, solvesdp

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "maxcut300_seed0": lambda: workloads.maxcut(workloads.laplacian_random(300, 0.5, 0)),
    "maxcut130_seed1": lambda: workloads.maxcut(workloads.laplacian_random(130, 0.5, 1)),
    "polyopt20_seed0": lambda: workloads.polyopt_random(20, 0),
    "delsarte_8_16": lambda: workloads.delsarte(8, 16, Fraction(1, 2)),
    "delsarte_8_32": lambda: workloads.delsarte(8, 32, Fraction(1, 2)),                               # BASELINE config 3 at full size
    "sphere_2_31_prec512": lambda: workloads.sphere_packing(8, 31, [Fraction(1, 2), Fraction(1, 2)], prec=512),   # config 5, examples/SpherePacking.jl:13 default precision
    "threepoint_4_10_10": lambda: workloads.three_point_bound(4, Fraction(1, 6), 10, 10),             # config 4 (examples/ThreePointBound.jl)
}
KW = {"threepoint_4_10_10": dict(omega_p=10 ** 3, omega_d=10 ** 3)}
# The two largest shapes of SURVEY.md §8(d) are beyond a full oracle solve in this container: one oracle iteration of the three-point
# bound at d = 14 (P = 894, 61 blocks up to n = 225) takes ~9 minutes on 8 cores, one of sphere packing (8,40) at 512 bit (N = 2953)
# ~25 minutes.  For d = 14 the golden holds the FIRST TWO iterations (objectives of the iterate at full precision and the printed row);
# the device must reproduce them (tests/test_gpu_parity.py).
FIRST = {"threepoint_4_14_14_first2": (lambda: workloads.three_point_bound(4, Fraction(1, 6), 14, 14), 2, dict(omega_p=10 ** 3, omega_d=10 ** 3))}

for name in [n for n in sys.argv[1:] if n in FIRST]:
    make, its, kw = FIRST[name]
    sdp = make()
    t = time.time()
    r = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30, oracle_skip_zeros=True, maxiterations=its, **kw)
    out = {"name": name, "describe": sdp.describe(), "iterations": r.iterations, "history": r.history,
           "d_obj": mpmath.nstr(r.d_obj, 70), "p_obj": mpmath.nstr(r.p_obj, 70), "gap": mpmath.nstr(r.gap, 20),
           "options": {"prec": sdp.prec, "duality_gap_threshold": 1e-30, "maxiterations": its, **{k: str(v) for k, v in kw.items()}}, "oracle_seconds": round(time.time() - t, 1)}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: v for k, v in out.items() if k != "history"}, flush=True)

for name in ([n for n in sys.argv[1:] if n in CASES] or ([] if sys.argv[1:] else list(CASES))):
    sdp = CASES[name]()
    t = time.time()
    r = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30, oracle_skip_zeros=True, **KW.get(name, {}))
    out = {"name": name, "describe": sdp.describe(), "status": r.status, "iterations": r.iterations,
           "d_obj": mpmath.nstr(r.d_obj, 70), "p_obj": mpmath.nstr(r.p_obj, 70), "gap": mpmath.nstr(r.gap, 20),
           "options": {"prec": sdp.prec, "duality_gap_threshold": 1e-30, **{k: str(v) for k, v in KW.get(name, {}).items()}}, "oracle_seconds": round(time.time() - t, 1)}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out, flush=True)

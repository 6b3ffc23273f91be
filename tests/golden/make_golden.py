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
from clrs_b200 import workloads, solvesdp

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "maxcut300_seed0": lambda: workloads.maxcut(workloads.laplacian_random(300, 0.5, 0)),
    "maxcut130_seed1": lambda: workloads.maxcut(workloads.laplacian_random(130, 0.5, 1)),
    "polyopt20_seed0": lambda: workloads.polyopt_random(20, 0),
    "delsarte_8_16": lambda: workloads.delsarte(8, 16, Fraction(1, 2)),
}

for name in (sys.argv[1:] or list(CASES)):
    sdp = CASES[name]()
    t = time.time()
    r = solvesdp(sdp, lib="oracle", duality_gap_threshold=1e-30, oracle_skip_zeros=True)
    out = {"name": name, "describe": sdp.describe(), "status": r.status, "iterations": r.iterations,
           "d_obj": mpmath.nstr(r.d_obj, 70), "p_obj": mpmath.nstr(r.p_obj, 70), "gap": mpmath.nstr(r.gap, 20),
           "options": {"prec": 256, "duality_gap_threshold": 1e-30}, "oracle_seconds": round(time.time() - t, 1)}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out, flush=True)

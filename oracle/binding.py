"""Registers the MPFR oracle (oracle/libclrs_oracle.so, symbols clrs_oracle_*) as the "oracle" library of the host
mirror, so that tests, __graft_entry__.smoke() and the CPU legs of bench.py can run the same Python driver against it:

    import oracle.binding            # noqa: F401
    solvesdp(sdp, lib="oracle", ...)

TEST INFRASTRUCTURE ONLY.  The product package (clusteredlowranksolver.jl_b200) never imports this module and cannot
reach the oracle without it: `Solver(lib="oracle")` raises "unknown library" unless this registration has run.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import clrs_b200  # noqa: E402

ORACLE_LIB = os.path.join(_HERE, "libclrs_oracle.so")
clrs_b200.register_backend("oracle", ORACLE_LIB, "clrs_oracle_")

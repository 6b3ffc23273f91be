"""Latency of the warp-cooperative arithmetic (kernel tuning, not a test): CLRS_WOPS_BENCH=1 python tools/gpu_wops.py"""
import os, sys
sys.path.insert(0, ".")
os.environ["CLRS_WOPS_BENCH"] = "1"
import ctypes as C
import clrs_b200
from clrs_b200 import workloads, Solver
for prec in (256, 512):
    S = Solver(workloads.maxcut(workloads.laplacian_cycle(3), prec=prec), lib="device")
    fn = S.lib.clrs_debug_selftest; fn.restype = C.c_int
    print("selftest mismatches", fn(S.h), flush=True)
    S.close()

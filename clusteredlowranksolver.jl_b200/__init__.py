"""clrs_b200 — B200-native hot path of ClusteredLowRankSolver.jl's interior-point solver.

The product is the C-ABI shared library csrc/libclrs_b200.so (hand-written
sm_100a CUDA; include/clrs_b200.h).  This package is the thin host mirror of the
reference's `solvesdp` interface used by tests and bench.py in place of the
Julia shim, plus generators for the BASELINE.json workloads.
"""
from . import wire
from .sdp import ClusteredSDP, Cluster, PSDBlock, LowRankTerm
from .api import (Solver, SolverFailure, solvesdp, IterInfo, Options, load_library, PHASES, DEVICE_LIB, register_backend,
                  nccl_unique_id, partition_clusters, plan_shards)

__all__ = ["wire", "ClusteredSDP", "Cluster", "PSDBlock", "LowRankTerm", "Solver", "SolverFailure", "solvesdp",
           "IterInfo", "Options", "load_library", "PHASES", "DEVICE_LIB", "register_backend", "nccl_unique_id", "partition_clusters", "plan_shards"]

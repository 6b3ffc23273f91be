"""ctypes binding of the C ABI (include/clrs_b200.h) and the host-side mirror of
`solvesdp(sdp::ClusteredLowRankSDP; kwargs...)` (src/solver.jl:100-744).

Julia is not installed in this image, so the reference-facing host code is this
Python mirror: same keyword names, same status classification
(src/solver.jl:727-741), same error codes (0 ok, 1 SolverFailure, 2 maxiter,
3 complementary gap, 4 short step).  The loop returns to the host after every
iteration exactly like the Julia shim in INTEGRATION.md, so verbose printing,
saving callbacks and maxiterations stay on the host.

`Solver(lib="device")` binds libclrs_b200.so (CUDA, no CPU fallback: creating
it without an sm_100 GPU raises).  The package itself binds ONLY that library.  Test infrastructure can register
a second library with the same entry points under another name (`register_backend`); `oracle/binding.py` does that
for the MPFR oracle, and only tests/, `__graft_entry__.smoke()` and the CPU legs of bench.py import it.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from fractions import Fraction
from typing import Optional

import numpy as np
import mpmath

from . import wire
from .sdp import ClusteredSDP

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
DEVICE_LIB = os.path.join(_HERE, "csrc", "libclrs_b200.so")

PHASES = ["decomp", "predictor", "corrector", "alpha", "Xinv", "R", "residuals", "schur", "cholS",
          "LinvB", "Q", "cholQ", "Z", "rhs_x", "solve", "dX", "dY"]


class Options(C.Structure):
    _fields_ = [("prec", C.c_int32), ("matmul_prec", C.c_int32),
                ("beta_infeasible", C.c_double), ("beta_feasible", C.c_double), ("gamma", C.c_double),
                ("omega_p", C.c_double), ("omega_d", C.c_double),
                ("duality_gap_threshold", C.c_double), ("dual_error_threshold", C.c_double),
                ("primal_error_threshold", C.c_double), ("max_complementary_gap", C.c_double),
                ("step_length_threshold", C.c_double),
                ("need_dual_feasible", C.c_int32), ("need_primal_feasible", C.c_int32),
                ("safe_step", C.c_int32), ("correctoronly", C.c_int32),
                ("device", C.c_int32), ("gemm_path", C.c_int32), ("sparse_schur", C.c_int32)]


class IterInfo(C.Structure):
    _fields_ = [("iter", C.c_int32), ("stop", C.c_int32), ("pd_feasible", C.c_int32), ("reserved", C.c_int32),
                ("mu", C.c_double), ("d_obj", C.c_double), ("p_obj", C.c_double), ("gap", C.c_double),
                ("err_P", C.c_double), ("err_p", C.c_double), ("err_d", C.c_double),
                ("alpha_d", C.c_double), ("alpha_p", C.c_double), ("beta_c", C.c_double),
                ("d_obj_new", C.c_double), ("p_obj_new", C.c_double), ("gap_new", C.c_double),
                ("phase_ms", C.c_double * 17)]


_OPT_IDS = {"beta_infeasible": 0, "beta_feasible": 1, "gamma": 2, "omega_p": 3, "omega_d": 4,
            "duality_gap_threshold": 5, "dual_error_threshold": 6, "primal_error_threshold": 7,
            "max_complementary_gap": 8, "step_length_threshold": 9}

_DEFAULTS = dict(beta_infeasible=Fraction(3, 10), beta_feasible=Fraction(1, 10), gamma=Fraction(9, 10),
                 omega_p=10 ** 10, omega_d=10 ** 10, duality_gap_threshold=1e-15,
                 dual_error_threshold=1e-30, primal_error_threshold=1e-30,
                 max_complementary_gap=10 ** 100, step_length_threshold=1e-7)


class SolverFailure(RuntimeError):
    """Mirror of the reference's SolverFailure (src/solver.jl:9-11)."""


_libs = {}
_BACKENDS = {"device": (DEVICE_LIB, "clrs_", C.RTLD_GLOBAL)}      # name -> (shared library, symbol prefix, dlopen mode)


def register_backend(name: str, path: str, prefix: str) -> None:
    """Make `Solver(lib=name)` bind another shared library exporting the ABI of include/clrs_b200.h under `prefix`.
    Used by oracle/binding.py (test infrastructure); the product never calls it."""
    _BACKENDS[name] = (path, prefix, C.RTLD_LOCAL)


def load_library(kind: str) -> C.CDLL:
    """Load a registered library; fail loudly if it is unknown or missing (there is no fallback between libraries)."""
    if kind in _libs:
        return _libs[kind]
    if kind not in _BACKENDS:
        raise RuntimeError(f"unknown library {kind!r}: this package binds only the CUDA library ('device'); the CPU oracle is "
                           f"test infrastructure and has to be registered explicitly (import oracle.binding)")
    path, _, mode = _BACKENDS[kind]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path, mode=mode)
    _libs[kind] = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    return a.ctypes.data_as(C.c_void_p), a


class Solver:
    """One solver handle (device or oracle) holding one uploaded SDP."""

    def __init__(self, sdp: ClusteredSDP, lib: str = "device", device: int = 0, gemm_path: int = 0,
                 matmul_prec: int = 0, oracle_skip_zeros: bool = False, comm=None, sparse_schur: bool = False, **kwargs):
        """comm = (rank, nranks, unique_id_bytes): shard the clusters of the SDP over `nranks` handles
        (one per GPU / process); every rank uploads the same SDP and calls iterate() in lock step.
        sparse_schur: let dense blocks with sparse constraint matrices form S from the nonzero entries (clrs_options.sparse_schur;
        off by default: the reference's dense path exploits no sparsity, src/solver.jl:1088)."""
        self.kind = lib
        self.lib = load_library(lib)
        self.pre = _BACKENDS[lib][1]
        self.sdp = sdp
        self.prec = sdp.prec
        self.W = wire.limbs_for(self.prec)
        self.dtype = wire.wire_dtype(self.prec)
        opts = dict(_DEFAULTS)
        flags = dict(need_dual_feasible=False, need_primal_feasible=False, safe_step=True, correctoronly=False)
        for k, v in kwargs.items():
            if k in opts:
                opts[k] = v
            elif k in flags:
                flags[k] = bool(v)
            else:
                raise TypeError(f"unknown solver option {k!r}")
        o = Options()
        o.prec = self.prec
        o.matmul_prec = matmul_prec
        for k, v in opts.items():
            setattr(o, k, float(v))
        for k, v in flags.items():
            setattr(o, k, int(v))
        o.device = device
        o.gemm_path = gemm_path
        o.sparse_schur = int(bool(sparse_schur))
        self.h = C.c_void_p()
        self._call("create", C.byref(o), C.byref(self.h), check_handle=False)
        # full-precision overrides: the reference converts these with Arb(v, prec=prec) (src/solver.jl:138-139)
        with mpmath.workprec(self.prec + 64):
            for k, v in opts.items():
                if isinstance(v, float):
                    continue      # a Float64 kwarg is converted exactly by the library
                val = mpmath.mpf(v.numerator) / v.denominator if isinstance(v, Fraction) else mpmath.mpf(v)
                w = wire.to_wire(val, self.prec)
                self._call("set_option_num", self.h, C.c_int(_OPT_IDS[k]), w.ctypes.data_as(C.c_void_p))
        self.rank, self.nranks = (comm[0], comm[1]) if comm else (0, 1)
        if comm and lib == "device" and comm[1] > 1:
            uid = (C.c_char * 128).from_buffer_copy(bytes(comm[2]))
            self._call("comm_init", self.h, C.c_int32(comm[0]), C.c_int32(comm[1]), uid)
        if lib != "device" and oracle_skip_zeros:
            fn = self._fn("set_dense_skip_zeros")
            fn.restype = None
            fn(self.h, C.c_int32(1))
        self._upload(sdp)
        self.finished = False

    # -- plumbing ---------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, self.pre + name)

    def _call(self, name, *args, check_handle=True):
        fn = self._fn(name)
        fn.restype = C.c_int
        rc = fn(*args)
        if rc != 0:
            msg = ""
            if self.h:
                le = self._fn("last_error")
                le.restype = C.c_char_p
                msg = (le(self.h) or b"").decode()
            if 10 <= rc <= 14:
                raise SolverFailure(f"[{rc}] {msg}")
            raise RuntimeError(f"{self.pre}{name} failed with code {rc}: {msg}")
        return rc

    def _upload(self, sdp: ClusteredSDP):
        vp = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
        keep = []

        def P(a):
            a = np.ascontiguousarray(a)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        self._call("set_free", self.h, C.c_int32(sdp.N), P(sdp.b), P(sdp.constant), C.c_int32(int(sdp.maximize)))
        for j, cl in enumerate(sdp.clusters):
            self._call("add_cluster", self.h, C.c_int32(j), C.c_int32(cl.P), P(cl.B), P(cl.c))
            for l, blk in enumerate(cl.blocks):
                self._call("add_block", self.h, C.c_int32(j), C.c_int32(l), C.c_int32(blk.m), C.c_int32(blk.delta),
                           C.c_int32(int(blk.high_rank)), P(blk.C))
                for p, A in blk.dense.items():
                    self._call("add_dense_term", self.h, C.c_int32(j), C.c_int32(l), C.c_int32(p), P(A))
                for p, (rows, cols, vals, mirror) in blk.sparse.items():
                    r32, c32 = np.ascontiguousarray(rows, dtype=np.int32), np.ascontiguousarray(cols, dtype=np.int32)
                    keep.extend([r32, c32])
                    self._call("add_sparse_term", self.h, C.c_int32(j), C.c_int32(l), C.c_int32(p), C.c_int32(len(r32)),
                               r32.ctypes.data_as(C.c_void_p), c32.ctypes.data_as(C.c_void_p), P(vals), C.c_int32(int(mirror)))
                for t in blk.lowrank:
                    self._call("add_lowrank_term", self.h, C.c_int32(j), C.c_int32(l), C.c_int32(t.r), C.c_int32(t.s),
                               C.c_int32(t.p), C.c_int32(int(t.lam.shape[0])), P(t.lam), P(t.vs), P(t.ws))
                keep.clear()
        self._call("finalize", self.h)

    def close(self):
        if getattr(self, "h", None):
            fn = self._fn("destroy")
            fn.restype = None
            fn(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the hot path -----------------------------------------------------
    def iterate(self) -> IterInfo:
        info = IterInfo()
        self._call("iterate", self.h, C.byref(info))
        return info

    def objectives(self):
        """(d_obj, p_obj, gap) of the current iterate as mpf (src/solver.jl:626-628)."""
        out = wire.wire_zeros((3,), self.prec)
        base = out.ctypes.data
        sz = out.dtype.itemsize
        self._call("get_objectives", self.h, C.c_void_p(base), C.c_void_p(base + sz), C.c_void_p(base + 2 * sz))
        v = wire.from_wire(out, self.prec)
        return v[0], v[1], v[2]

    # -- state ------------------------------------------------------------
    def matrix_count(self) -> int:
        fn = self._fn("state_matrix_count")
        fn.restype = C.c_int64
        return int(fn(self.h))

    def state_buffers(self, matrices: bool = True, alloc=None):
        """Caller-owned wire buffers for (x, X, y, Y).  `alloc(nbytes) -> uint8 ndarray` lets the caller supply the memory
        (bench.py passes pinned host memory so that clrs_set_state / clrs_get_state copy by DMA)."""
        sdp = self.sdp
        cnt = self.matrix_count()

        def mk(n):
            if alloc is None:
                return wire.wire_zeros((n,), self.prec)
            raw = alloc(n * self.dtype.itemsize)
            raw[:] = 0
            return raw.view(self.dtype)
        return mk(sdp.num_constraints), (mk(cnt) if matrices else None), mk(max(sdp.N, 1)), (mk(cnt) if matrices else None)

    def get_state(self, matrices: bool = True, out=None):
        """Download (x, X, y, Y).  out = buffers from state_buffers() to write into (y has max(N, 1) records)."""
        sdp = self.sdp
        x, X, y, Y = out if out is not None else self.state_buffers(matrices)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        self._call("get_state", self.h, vp(x), vp(X), vp(y), vp(Y))
        return x, X, y[:sdp.N], Y

    def set_state(self, x=None, X=None, y=None, Y=None):
        vp = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p) if a is not None else None
        self._call("set_state", self.h, vp(x), vp(X), vp(y), vp(Y))

    def debug_get(self, what: str, j: int = 0, l: int = 0, capacity: int = 0) -> np.ndarray:
        fn = self._fn("debug_get")
        fn.restype = C.c_int64
        cap = capacity or 1
        buf = wire.wire_zeros((cap,), self.prec)
        n = fn(self.h, what.encode(), C.c_int32(j), C.c_int32(l), buf.ctypes.data_as(C.c_void_p), C.c_int64(cap))
        if n < 0 and -n > cap:
            buf = wire.wire_zeros((-n,), self.prec)
            n = fn(self.h, what.encode(), C.c_int32(j), C.c_int32(l), buf.ctypes.data_as(C.c_void_p), C.c_int64(-n))
        if n < 0:
            raise KeyError(what)
        return buf[:n]

    def owned_state_bytes(self) -> int:
        """Bytes of (x, X, y, Y) this handle reads in set_state / writes in get_state: the clusters it owns plus y."""
        n = self.sdp.N
        for j, cl in enumerate(self.sdp.clusters):
            if self.kind == "device" and self.nranks > 1:
                mine = [b for l, b in enumerate(cl.blocks) if self.block_owner(j, l) == self.rank]
                if not mine and self.cluster_owner(j) != self.rank:
                    continue
                n += cl.P + 2 * sum(b.n * b.n for b in mine)
                continue
            n += cl.P + 2 * sum(b.n * b.n for b in cl.blocks)
        return n * self.dtype.itemsize

    def cluster_owner(self, j: int) -> int:
        fn = self._fn("cluster_owner"); fn.restype = C.c_int
        return int(fn(self.h, C.c_int32(j)))

    def block_owner(self, j: int, l: int) -> int:
        """Rank that holds block l of cluster j (differs from cluster_owner(j) only inside a split cluster)."""
        fn = self._fn("block_owner"); fn.restype = C.c_int
        return int(fn(self.h, C.c_int32(j), C.c_int32(l)))

    # -- measurement hooks (device library only) ------------------------------
    def profile(self, enable: bool):
        fn = self._fn("profile"); fn.restype = None; fn(self.h, C.c_int32(int(enable)))

    def profile_get(self):
        out = (C.c_double * 10)()
        fn = self._fn("profile_get"); fn.restype = None; fn(self.h, out)
        names = ("dp4a", "tc_small", "tc_large")
        d = {"kernel_launches": int(out[9])}
        for i, nm in enumerate(names):
            d[nm] = {"ms": out[3 * i], "mp_flops": out[3 * i + 1], "launches": int(out[3 * i + 2])}
        return d

    def use_graph(self, enable: bool):
        """CUDA-graph replay of the iteration on/off (device library; on by default)."""
        fn = self._fn("use_graph"); fn.restype = None; fn(self.h, C.c_int32(int(enable)))

    def last_iteration_ms(self) -> float:
        fn = self._fn("last_iteration_ms"); fn.restype = C.c_double
        return float(fn(self.h))

    # -- standalone kernels -------------------------------------------------
    def mp_gemm(self, A: np.ndarray, B: np.ndarray, path: int = 0):
        M, K = A.shape
        K2, N = B.shape
        assert K == K2
        Cw = wire.wire_zeros((M, N), self.prec)
        A = np.ascontiguousarray(A)
        B = np.ascontiguousarray(B)
        if self.kind == "device":
            ms = C.c_double(0)
            self._call("mp_gemm", self.h, C.c_int32(M), C.c_int32(N), C.c_int32(K), A.ctypes.data_as(C.c_void_p),
                       B.ctypes.data_as(C.c_void_p), Cw.ctypes.data_as(C.c_void_p), C.c_int32(path), C.byref(ms))
            return Cw, ms.value
        self._call("mp_gemm", self.h, C.c_int32(M), C.c_int32(N), C.c_int32(K), A.ctypes.data_as(C.c_void_p),
                   B.ctypes.data_as(C.c_void_p), Cw.ctypes.data_as(C.c_void_p))
        return Cw, 0.0

    def mp_cholesky(self, A: np.ndarray):
        n = A.shape[0]
        L = wire.wire_zeros((n, n), self.prec)
        A = np.ascontiguousarray(A)
        self._call("mp_cholesky", self.h, C.c_int32(n), A.ctypes.data_as(C.c_void_p), L.ctypes.data_as(C.c_void_p))
        return L

    def mp_qr_pivot(self, A: np.ndarray):
        """Column-pivoted QR (clrs_mp_qr_pivot): returns (R wire (min(m,n), n) in the pivoted column order, perm int32 (n,))."""
        m, n = A.shape
        R = wire.wire_zeros((min(m, n), n), self.prec)
        perm = np.zeros(n, dtype=np.int32)
        A = np.ascontiguousarray(A)
        self._call("mp_qr_pivot", self.h, C.c_int32(m), C.c_int32(n), A.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p),
                   perm.ctypes.data_as(C.c_void_p))
        return R, perm


STATUS = ["Optimal", "NearOptimal", "Feasible", "PrimalFeasible", "DualFeasible", "NotConverged"]


class SolveResult:
    def __init__(self):
        self.status = "NotConverged"
        self.error_code = 0
        self.iterations = 0
        self.time = 0.0
        self.d_obj = self.p_obj = self.gap = None
        self.history = []
        self.solver = None

    def __repr__(self):
        return (f"SolveResult(status={self.status}, error_code={self.error_code}, iterations={self.iterations}, "
                f"d_obj={mpmath.nstr(self.d_obj, 30)}, p_obj={mpmath.nstr(self.p_obj, 30)}, gap={mpmath.nstr(self.gap, 5)})")


def solvesdp(sdp: ClusteredSDP, lib: str = "device", maxiterations: int = 500, verbose: bool = False,
             callback=None, keep_solver: bool = False, device: int = 0, gemm_path: int = 0,
             oracle_skip_zeros: bool = False, comm=None, **kwargs) -> SolveResult:
    """Mirror of solvesdp(sdp; kwargs...) (src/solver.jl:100-744) above the C ABI."""
    res = SolveResult()
    S = Solver(sdp, lib=lib, device=device, gemm_path=gemm_path, oracle_skip_zeros=oracle_skip_zeros, comm=comm, **kwargs)
    gap_thr = float(kwargs.get("duality_gap_threshold", 1e-15))
    derr_thr = float(kwargs.get("dual_error_threshold", 1e-30))
    perr_thr = float(kwargs.get("primal_error_threshold", 1e-30))
    t0 = time.time()
    if verbose:
        print("%5s %8s %11s %11s %11s %10s %10s %10s %10s %10s %10s %10s" % (
            "iter", "time(s)", "mu", "D-obj", "P-obj", "gap", "D-error", "d-error", "p-error", "a_d", "a_p", "beta"))
    pd_feas = False
    last = None
    try:
        while True:
            if res.iterations + 1 > maxiterations:      # src/solver.jl:362-366
                res.error_code = 2
                break
            info = S.iterate()
            if info.stop in (1, 2, 3):                   # terminate() fired at the loop top
                pd_feas = bool(info.pd_feasible)
                break
            if info.stop == 4:
                res.error_code = 3
                break
            last = info
            pd_feas = bool(info.pd_feasible)
            if info.stop == 5:
                res.error_code = 4
                break
            res.iterations += 1
            res.history.append({k: getattr(info, k) for k, _ in IterInfo._fields_ if k not in ("phase_ms", "reserved")}
                               | {"phase_ms": list(info.phase_ms)})
            if verbose:
                print("%5d %8.1f %11.3e %11.3e %11.3e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e" % (
                    info.iter, time.time() - t0, info.mu, info.d_obj, info.p_obj, info.gap, info.err_P, info.err_p,
                    info.err_d, info.alpha_d, info.alpha_p, info.beta_c))
            if callback is not None:
                callback(S, info)
    except SolverFailure as e:                           # src/solver.jl:594-623
        if verbose:
            print("SolverFailure:", e)
        res.error_code = 1
        res.failure = str(e)
    res.time = time.time() - t0
    res.d_obj, res.p_obj, res.gap = S.objectives()
    derr = max(last.err_P, last.err_p) if last is not None else float("inf")
    perr = last.err_d if last is not None else float("inf")
    gap = float(res.gap)
    if pd_feas and gap < gap_thr:                        # src/solver.jl:727-741
        res.status = "Optimal"
    elif (pd_feas and gap < 1e-8) or (derr < 1e-15 and perr < 1e-15 and gap < 1e-8):
        res.status = "NearOptimal"
    elif pd_feas:
        res.status = "Feasible"
    elif perr < perr_thr:
        res.status = "PrimalFeasible"
    elif derr < derr_thr:
        res.status = "DualFeasible"
    if keep_solver:
        res.solver = S
    else:
        S.close()
    return res


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId for Solver(comm=...); create on rank 0 and broadcast to the other ranks."""
    lib = load_library("device")
    buf = (C.c_char * 128)()
    fn = lib.clrs_comm_unique_id
    fn.restype = C.c_int
    if fn(buf) != 0:
        raise RuntimeError("clrs_comm_unique_id failed (libnccl.so.2 not found?)")
    return bytes(buf)


def partition_clusters(weights, nranks: int):
    """The library's cluster -> rank partitioner (host only; usable without a GPU)."""
    lib = load_library("device")
    w = (C.c_double * len(weights))(*[float(v) for v in weights])
    out = (C.c_int32 * len(weights))()
    fn = lib.clrs_partition_clusters
    fn.restype = C.c_int
    if fn(C.c_int32(len(weights)), w, C.c_int32(nranks), out) != 0:
        raise ValueError("clrs_partition_clusters failed")
    return list(out)


def plan_shards(sdp: ClusteredSDP, nranks: int, split_mode: int = -1, big_cluster: int = 512):
    """The library's shard plan for `sdp` on `nranks` ranks (host only): (cluster_owner, split, block_owner[j][l]).
    Same weights as clrs_finalize: P_j^3 per cluster, 15 n^3 per block ((2 #A_p + 15) n^3 for a dense block)."""
    lib = load_library("device")
    J = len(sdp.clusters)
    nb = [len(c.blocks) for c in sdp.clusters]
    bw = [float(b.n) ** 3 * (2.0 * (len(b.dense) + len(b.sparse)) + 15 if b.high_rank else 15) for c in sdp.clusters for b in c.blocks]
    p3 = (C.c_double * J)(*[float(c.P) ** 3 for c in sdp.clusters])
    big = (C.c_int32 * J)(*[int(nranks > 1 and sdp.N > 0 and big_cluster > 0 and c.P >= big_cluster) for c in sdp.clusters])
    own, spl, bown = (C.c_int32 * J)(), (C.c_int32 * J)(), (C.c_int32 * max(1, len(bw)))()
    fn = lib.clrs_plan_shards
    fn.restype = C.c_int
    if fn(C.c_int32(J), p3, (C.c_int32 * J)(*nb), (C.c_double * max(1, len(bw)))(*bw), big, C.c_int32(nranks), C.c_int32(split_mode), own, spl, bown) != 0:
        raise ValueError("clrs_plan_shards failed")
    out, o = [], 0
    for n in nb:
        out.append(list(bown[o:o + n])); o += n
    return list(own), [bool(v) for v in spl], out

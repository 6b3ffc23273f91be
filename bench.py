#!/usr/bin/env python
"""bench.py — IPM iterations/sec of the B200 hot path (BASELINE.json metric).

One "step" is one predictor-corrector IPM iteration (src/solver.jl:362-592) on the resident SDP.

Workloads
  N = 1   BASELINE.json configs[1]: Goemans-Williamson MAX-CUT relaxation, Laplacian of G(300, 0.5)
          (numpy default_rng(0)), every constraint matrix E_pp stored dense (the reference's dense
          Schur path), prec = 256 bit.  One cluster with one block: it does not shard.
  N > 1   BASELINE.json configs[4]: sphere packing n = 8, d = 31, four radii (SURVEY.md §8(d) size
          (4,31): J = 16 clusters, P = 1294, N = 641 free variables), prec = 512 bit (the example's default; at 256 bit
          the reference's own algorithm fails on this shape: "Q was not decomposed correctly"), clusters sharded
          over the ranks, strong scaling.  The line also carries the 1-GPU rate of the
          SAME workload measured in the same run on rank 0 (`strong_scaling`), because the N = 1 line of
          this script is the MAX-CUT headline.

Keys
  value     iterations/s from the device time of K consecutive clrs_iterate calls (CUDA events on the
            library's stream, SDP and iterate resident in HBM); every call is asserted to be a real
            iteration (info.stop == 0).  The solver runs with duality_gap_threshold = 1e-30 and the iterate
            is put back to a snapshot (iteration 3 of the solve) every 24 iterations, between timed calls,
            so no timed, e2e or profiled call can fall on a converged (no-op) iteration.
  e2e       the same K iterations driven through the C ABI with HOST buffers every step: clrs_set_state
            (H2D of x, X, y, Y from pinned host memory) + clrs_iterate + clrs_get_state (D2H), wall clock
  roofline  the dominant GEMM class: canonical int8 ops 2*M*N*K*528 / CUDA-event time of
            every launch, against 2 x the measured bf16 peak (burst: the launches are timed one by one)
  cpu_baseline  the MPFR oracle (a port of the reference, not the reference) on the host cores, bounded sample
  configs   (N = 1) iterations/s, ms/step and the 17 phase timers of the other BASELINE configs

--impl reference runs only the CPU arm on the same workload as the B200 arm at that N.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ipm_iterations_per_sec"
UNIT = "iterations/s"
RESTORE_EVERY = 24          # iterations after which the iterate is put back to the snapshot (config 2 converges in 55)
SNAP_ITER = 3               # the snapshot is the iterate after this many iterations from the default start

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from `ncu --set full` captures of this
# build (profiles/): keyed by (workload, class).  None = not captured for this build.
NCU_TRAFFIC = {
    # tc::k_gemm_tc grid (2,148,1) on one 63-constraint chunk (18900 x 300 x 300): 668.8 MB read + 176.2 MB written per launch
    # (profiles/r02_tc_kernel_ncu_full.txt, final build; algorithmic: 212 MB of left-operand planes in, 224 MB of byte planes + carry plane out)
    ("maxcut", "tc_large"): 845.0e6,
}


def workload(n, kind="maxcut"):
    import clrs_b200  # noqa: F401
    from clrs_b200 import workloads
    from fractions import Fraction as F
    if kind == "sphere":        # BASELINE.json configs[4]: many clusters coupled through the free variables; shards by cluster
        # SURVEY.md §8(d) size (4,31): J=16, N=641, P=1294.  prec = 512 (the default of examples/SpherePacking.jl:13): at 256 bit the
        # reference's own algorithm fails on this shape in the first iteration ("Q was not decomposed correctly", oracle and device alike)
        return workloads.sphere_packing(8, n if n != 300 else 31, [F(1, 2), F(1, 2), F(3, 4), F(1)], prec=512)
    if kind == "sphere2":       # (2,31): J=7, N=193, P=389
        return workloads.sphere_packing(8, n if n != 300 else 31, [F(1, 2), F(1, 2)], prec=512)
    if kind == "sphere8":       # (8,40): J=46, N=2953, P=5948 — the largest shape of SURVEY.md §8(d) (device only: one oracle iteration is ~1e11 512-bit MACs)
        return workloads.sphere_packing(8, n if n != 300 else 40, [F(1, 2), F(5, 8), F(3, 4), F(7, 8), F(1), F(9, 8), F(5, 4), F(11, 8)], prec=512)
    if kind == "threepoint14":  # configs[3] at d2 = d3 = 14: P=894, 61 blocks up to n=225
        return workloads.three_point_bound(4, F(1, 6), 14, 14, prec=256)
    if kind == "threepoint":    # configs[3] at the size of examples/ThreePointBound.jl (one cluster: the dense F_k blocks are shared)
        return workloads.three_point_bound(4, F(1, 6), n if n != 300 else 10, n if n != 300 else 10, prec=256)
    if kind == "delsarte":      # configs[2]
        return workloads.delsarte(8, n if n != 300 else 16, F(1, 2), prec=256)
    if kind == "polyopt":       # configs[0]
        return workloads.polyopt_random(n if n != 300 else 20, 0, prec=256)
    return workloads.maxcut(workloads.laplacian_random(n, 0.5, 0), prec=256)


CONFIG_INDEX = {"polyopt": 0, "maxcut": 1, "delsarte": 2, "threepoint": 3, "threepoint14": 3, "sphere": 4, "sphere2": 4, "sphere8": 4}


def config(n, n_gpus, kind="maxcut", sdp=None):
    if kind != "maxcut":
        return {"workload": f"{sdp.describe() if sdp is not None else kind} (BASELINE.json configs[{CONFIG_INDEX[kind]}])",
                "step": "one predictor-corrector IPM iteration", "l2": "L2 flushed by the iteration's own temporaries only; latency-bound small blocks",
                "parallelism": ("clusters sharded over ranks (LPT on P^3 + sum n^3), NCCL on the free-variable coupling (Q, u, p) and scalars" if kind.startswith("sphere") else
                                "one cluster: its PSD blocks sharded over ranks (LPT on n^3), S_j and <A_*,.> all-reduced, factor and solves replicated" if kind.startswith("threepoint") else "replicas only") if n_gpus > 1 else "single GPU"}
    return {"workload": f"GW MAX-CUT relaxation, G({n},0.5) numpy default_rng(0), dense constraint path "
                        f"(BASELINE.json configs[1]), prec=256, J=1 P={n} one dense block {n}x{n}, N=0",
            "step": "one predictor-corrector IPM iteration", "l2": "working set (~7 GB of slices/temporaries) exceeds the 126 MB L2",
            "parallelism": "replicas only" if n_gpus > 1 else "single GPU"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self):
        self.rows = []          # nvidia-smi rows (fallback)
        self.proc = None
        self.samples = []       # (sm_mhz, reason bitmask) from NVML
        self.stop_flag = False
        self.thread = None
        self.nvml = None
        self.max_mhz = None

    def start(self, device=0):
        # NVML in a polling thread (a sample every ~5 ms, so even a 100 ms timed region is covered); nvidia-smi -lms as fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and vis.split(",")[device].isdigit() else device
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml

            def poll():
                while not self.stop_flag:
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, device=0):
        if self.nvml is not None:
            self.stop_flag = True
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            mask = 0
            for _, m in self.samples:
                mask |= m
            return {"sm_mhz": statistics.median(x for x, _ in self.samples), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, v in names.items() if mask & v), "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 9 and r[0] == str(device)]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows), "source": "nvidia-smi"}


def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed to run on (torchrun's OMP_NUM_THREADS=1 is ignored on purpose)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm(n, steps, warmup, target_seconds=12.0, kind="maxcut", budget_seconds=100.0):
    """The oracle on the host cores, bounded sample.  Returns (it/s, OpenMP threads actually used, description)."""
    import ctypes as C
    from clrs_b200 import Solver
    import oracle.binding  # noqa: F401  (the one place bench.py touches oracle/: the CPU baseline / reference arm)
    sdp = workload(n, kind)
    S = Solver(sdp, lib="oracle", oracle_skip_zeros=True, duality_gap_threshold=1e-30)
    lib = S.lib
    lib.clrs_oracle_set_sample_limit.restype = None
    lib.clrs_oracle_get_sample_times.restype = None
    lib.clrs_oracle_set_threads.restype = None
    lib.clrs_oracle_get_threads.restype = C.c_int32
    lib.clrs_oracle_set_threads(C.c_int32(host_threads()))      # explicit: the environment may carry OMP_NUM_THREADS=1 (torchrun)
    threads = int(lib.clrs_oracle_get_threads())
    limit = 1
    times = []
    desc = ""
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        if times and time.perf_counter() - t_begin > budget_seconds:      # bounded: the arm must end within minutes whatever K is
            desc += f"; stopped after {len(times)} timed steps ({budget_seconds:.0f} s budget)"
            break
        lib.clrs_oracle_set_sample_limit(S.h, C.c_int32(limit))
        t0 = time.perf_counter()
        info = S.iterate()
        wall = time.perf_counter() - t0
        if info.stop != 0:
            raise RuntimeError(f"CPU arm: iteration {it} did not run (stop = {info.stop})")
        out = (C.c_double * 3)()
        lib.clrs_oracle_get_sample_times(S.h, out)
        t_plain, t_skip, np_ = out[0], out[1], int(out[2])
        if np_ == 0:            # no dense Schur rows (low-rank workloads): the iteration ran in full
            if it >= warmup:
                times.append(wall)
            desc = f"full iterations of {sdp.describe()}"
            continue
        est = (wall - t_plain - t_skip) + t_plain * np_ / max(limit, 1)
        if it >= warmup:
            times.append(est)
        desc = (f"per step: {limit} of {np_} rows of the dense Schur path (X^-1 A_p Y and its {np_} inner products) at full "
                f"cost, scaled x{np_}/{limit}; the rest of the iteration in full ({wall - t_plain - t_skip:.1f} s)")
        per_row = t_plain / max(limit, 1)
        limit = max(1, min(np_, int(target_seconds / max(per_row, 1e-3))))
    S.close()
    its = len(times) / sum(times)
    return its, threads, desc


class Runner:
    """A Solver plus the snapshot discipline: every call of step() is a real iteration (asserted), and the iterate goes
    back to the snapshot every RESTORE_EVERY iterations, outside the device-timed part of a step."""

    def __init__(self, S, alloc=None):
        self.S = S
        for _ in range(SNAP_ITER):
            self._iterate()
        self.snap = tuple(None if a is None else a.copy() for a in S.get_state(out=S.state_buffers()))
        self.since = 0
        self.alloc = alloc

    def _iterate(self):
        info = self.S.iterate()
        if info.stop != 0:
            raise RuntimeError(f"bench: clrs_iterate returned stop = {info.stop} (a timed call must be a real iteration)")
        return info

    def restore(self):
        self.S.set_state(*self.snap)
        self.since = 0

    def step(self):
        if self.since >= RESTORE_EVERY:
            self.restore()
        info = self._iterate()
        self.since += 1
        return info


def measure(S, K, W, sync, alloc, e2e=True):
    """(device ms total, wall s, launches, e2e seconds, bytes per e2e step, last info) of K steps after W warm-ups."""
    R = Runner(S, alloc)
    for _ in range(W):
        R.step()
    R.restore()
    S.profile(False)            # resets the launch counter
    sync()
    t0 = time.perf_counter()
    dev_ms = 0.0
    info = None
    for _ in range(K):
        info = R.step()
        dev_ms += S.last_iteration_ms()
    sync()
    wall = time.perf_counter() - t0
    launches = S.profile_get()["kernel_launches"]
    if not e2e:
        return R, dev_ms, wall, launches, None, 0, info
    # ---- end to end through the C ABI with host buffers (pinned) every step ----
    bufs = S.state_buffers(alloc=alloc)

    def load_snapshot():
        for dst, src in zip(bufs, R.snap):
            dst[:len(src)] = src          # (y has max(N, 1) records in the buffer, N in the snapshot)
    load_snapshot()
    x, X, y, Y = bufs
    nbytes = S.owned_state_bytes()
    sync()
    t0 = time.perf_counter()
    for k in range(K):
        if k and k % RESTORE_EVERY == 0:
            load_snapshot()
        S.set_state(x, X, y, Y)
        info2 = S.iterate()
        if info2.stop != 0:
            raise RuntimeError(f"bench e2e: clrs_iterate returned stop = {info2.stop}")
        S.get_state(out=bufs)
    sync()
    e2e_s = time.perf_counter() - t0
    R.restore()
    return R, dev_ms, wall, launches, e2e_s, nbytes, info


def main():
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (build commands, NCCL's version
    # banner, library chatter) is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", dest="n", type=int, default=300, help="workload size parameter (300 = the BASELINE size of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-time-to-gap", action="store_true", help="skip the full solve to gap 1e-30 (for runs under ncu)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table (configs 1, 3, 4, 5 at one GPU)")
    ap.add_argument("--gemm-path", type=int, default=0)
    ap.add_argument("--workload", default=None, choices=["maxcut", "sphere", "sphere2", "sphere8", "threepoint", "threepoint14", "delsarte", "polyopt"],
                    help="default: maxcut (BASELINE configs[1], the metric's config) at one GPU; sphere (configs[4], sharded by cluster) at N > 1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    kind = args.workload or ("maxcut" if n_gpus == 1 else "sphere")

    if args.impl == "reference":
        if rank != 0:
            return
        its, threads, desc = cpu_arm(args.n, max(1, args.steps), max(0, min(args.warmup, 1)), kind=kind)
        sdp = workload(args.n, kind) if kind != "maxcut" else None
        print(file=real_stdout, flush=True, *[json.dumps({"impl": "reference", "metric": METRIC, "value": its, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "strong" if ((kind.startswith("sphere") or kind.startswith("threepoint")) and n_gpus > 1) else "weak", "vs_baseline": None,
                          "dtype": "mpfr (cpu)", "data": "synthetic", "config": config(args.n, args.gpus, kind, sdp),
                          "cpu_baseline": {"value": its, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc, "host_cpus": os.cpu_count()},
                          "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})])
        return

    import torch
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if dist is not None:
        dist.barrier()
    import clrs_b200
    from clrs_b200 import Solver
    sdp = workload(args.n, kind)
    sharded = world > 1 and kind in ("sphere", "sphere2", "sphere8", "threepoint", "threepoint14")      # clusters / blocks of a split cluster over the ranks
    comm = None
    if sharded:                 # one communicator over the ranks; the id travels through torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(clrs_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        comm = (rank, world, bytes(uid.cpu().tolist()))
    GAP = 1e-30
    S = Solver(sdp, lib="device", device=local_rank, gemm_path=args.gemm_path, comm=comm, duality_gap_threshold=GAP)
    reps = 1 if sharded else world
    W, K = max(args.warmup, 3), args.steps
    comm_info = None
    if sharded:                 # what the library's own plan gave every rank (stderr: one line per rank; JSON: the plan as rank 0 sees it)
        try:
            owners = [S.cluster_owner(j) for j in range(len(sdp.clusters))]
            blocks = [[S.block_owner(j, l) for l in range(len(c.blocks))] for j, c in enumerate(sdp.clusters)]
            mine_c = [j for j, o in enumerate(owners) if o == rank]
            mine_b = sum(1 for bl in blocks for o in bl if o == rank)
            print(f"[rank {rank}/{world}] cuda:{local_rank} leads clusters {mine_c}, holds {mine_b} of {sum(len(b) for b in blocks)} PSD blocks", file=sys.stderr, flush=True)
            comm_info = {"nranks": world, "backend": "NCCL (ncclCommInitRank inside libclrs_b200, id broadcast with torch.distributed)",
                         "clusters_led_per_rank": [sum(1 for o in owners if o == r) for r in range(world)],
                         "blocks_held_per_rank": [sum(1 for bl in blocks for o in bl if o == r) for r in range(world)],
                         "split_clusters": [j for j, bl in enumerate(blocks) if len(set(bl)) > 1]}
        except Exception as e:      # reporting only
            comm_info = {"nranks": world, "error": str(e)}

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def pinned(nbytes):
        return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, pin_memory=True).numpy()

    def allmax(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler()
    if rank == 0:
        sampler.start(local_rank)
    R, dev_ms, wall, launches, e2e_s, nbytes, last_info = measure(S, K, W, sync, pinned)
    clocks = sampler.stop(local_rank) if rank == 0 else None
    dev_ms_max = allmax(dev_ms)
    value = reps * K / (dev_ms_max / 1e3)
    e2e_value = reps * K / allmax(e2e_s)
    nbytes_all = nbytes
    if dist is not None:
        t = torch.tensor([float(nbytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nbytes_all = int(t.item())

    # ---- roofline of the dominant kernel: per-launch CUDA-event timing of every GEMM ----
    NP = min(K, 2)
    S.profile(True)
    for _ in range(NP):
        R.step()
    prof = S.profile_get()
    S.profile(False)
    R.restore()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_burst = peaks.get("bf16_tflops") or 1400.0
    bf16_sus = peaks.get("bf16_tflops_sustained") or bf16_burst
    peak_src = ("2 x bf16_tflops (burst) of MEASURED_PEAKS.json: int8 dense = 2x the bf16 rate; burst because every launch is timed alone between CUDA events"
                if peaks else "2 x 1.4 PFLOP/s fallback of B200_PROFILING.md")
    classes = {"dp4a": "k_gemm_dp4a (CUDA-core int8 path, small shapes)",
               "tc_small": "tc::k_gemm_tc + recombination, products with < 1e6 outputs (block-level n x n products, split-K Schur dots)",
               "tc_large": "tc::k_gemm_tc + k_tc_recombine, products with >= 1e6 outputs (the two 90000 x 300 x 300 products of the dense Schur path, in wave-sized chunks of constraints)"}

    def rl(c):
        ms, mpf, nl = prof[c]["ms"], prof[c]["mp_flops"], prof[c]["launches"]
        ach = (mpf * 528.0 / (ms / 1e3)) / 1e12 if ms > 0 else 0.0
        return {"kernel": classes[c], "achieved": ach, "frac": ach / (2 * bf16_burst), "frac_of_sustained_peak": ach / (2 * bf16_sus), "launches": nl,
                "avg_launch_ms": ms / max(1, nl), "share_of_step": ms / max(1e-9, NP * dev_ms_max / K)}
    dom = max(classes, key=lambda c: prof[c]["ms"])
    r = rl(dom)
    if r["launches"] <= 0:
        raise RuntimeError("bench: the profiled iterations launched no GEMM (roofline would be empty)")
    traffic = NCU_TRAFFIC.get((kind, dom))
    roofline = {"bound": "tensor", "kernel": r["kernel"], "achieved": r["achieved"], "peak": 2 * bf16_burst,
                "unit": "TOP/s (int8, canonical 2*M*N*K*528 per 256-bit GEMM)", "frac": r["frac"], "frac_of_sustained_peak": r["frac_of_sustained_peak"],
                "traffic": traffic, "peak_source": peak_src, "launches": r["launches"], "avg_launch_ms": r["avg_launch_ms"], "gemm_share_of_step": r["share_of_step"],
                "other_gemm_classes": {c: rl(c) for c in classes if c != dom}}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K,
           "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": f"int8 slices of {sdp.prec}-bit mantissas (exact int32 accumulation)",
           "data": "synthetic", "config": config(args.n, world, kind, sdp), "wall_ms_per_step": 1e3 * wall / K,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes_all, "d2h_bytes_per_step": nbytes_all,
                   "path": "clrs_set_state + clrs_iterate + clrs_get_state with pinned host wire buffers"},
           "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "comm": comm_info,
           "phase_ms": dict(zip(clrs_b200.PHASES, [round(v, 4) for v in last_info.phase_ms])),
           "solver": {"duality_gap_threshold": GAP, "snapshot_iteration": SNAP_ITER, "restore_every": RESTORE_EVERY,
                      "cuda_graph": os.environ.get("CLRS_GRAPH", "1") != "0"}}
    S.close()

    # ---- strong-scaling base: the same workload on ONE GPU, measured on rank 0 in the same run ----
    if sharded:
        if rank == 0:
            S1 = Solver(sdp, lib="device", device=local_rank, gemm_path=args.gemm_path, duality_gap_threshold=GAP)
            _, dms1, _, _, _, _, _ = measure(S1, K, W, torch.cuda.synchronize, None, e2e=False)
            S1.close()
            v1 = K / (dms1 / 1e3)
            out["strong_scaling"] = {"value_1gpu": v1, "ms_per_step_1gpu": dms1 / K, "speedup": value / v1, "n_gpus": world,
                                     "note": "same SDP, same build, unsharded handle on rank 0's GPU while the other ranks wait"}
        dist.barrier()

    # ---- second sharded workload in the same run: BASELINE configs[3] (three-point bound, ONE cluster) sharded by PSD blocks ----
    if sharded and args.workload is None and not args.no_configs:
        try:
            sdp2 = workload(300, "threepoint")
            uid2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid2 = torch.tensor(list(clrs_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(uid2, 0)
            S2 = Solver(sdp2, lib="device", device=local_rank, comm=(rank, world, bytes(uid2.cpu().tolist())), duality_gap_threshold=GAP)
            owners = sorted({S2.block_owner(0, l) for l in range(len(sdp2.clusters[0].blocks))})
            _, dms2, _, _, _, _, info2 = measure(S2, 4, 3, sync, None, e2e=False)
            S2.close()
            dms2 = allmax(dms2)
            entry = {"workload": sdp2.describe(), "config_index": 3, "value": 4 / (dms2 / 1e3), "ms_per_step": dms2 / 4, "n_gpus": world, "scaling": "strong",
                     "ranks_holding_blocks": owners, "parallelism": config(300, world, "threepoint", sdp2)["parallelism"],
                     "phase_ms": dict(zip(clrs_b200.PHASES, [round(v, 4) for v in info2.phase_ms]))}
            if rank == 0:
                S3 = Solver(sdp2, lib="device", device=local_rank, duality_gap_threshold=GAP)
                _, dms3, _, _, _, _, _ = measure(S3, 4, 3, torch.cuda.synchronize, None, e2e=False)
                S3.close()
                entry["value_1gpu"] = 4 / (dms3 / 1e3); entry["ms_per_step_1gpu"] = dms3 / 4; entry["speedup"] = (dms3 / 4) / (dms2 / 4)
            out["sharded_configs"] = {"threepoint": entry}
        except Exception as e:          # a side table must never take the headline down
            out["sharded_configs"] = {"threepoint": {"error": str(e)}}
        dist.barrier()

    if rank == 0 and world == 1:
        # ---- the other BASELINE configs at one GPU: it/s, ms/step, phase timers ----
        if not args.no_configs:
            from clrs_b200.api import PHASES
            table = {}
            for name in ("polyopt", "delsarte", "threepoint", "sphere2", "sphere"):
                if name == kind:
                    continue
                try:
                    s2 = workload(300, name)
                    S2 = Solver(s2, lib="device", device=local_rank, duality_gap_threshold=GAP)
                    _, dms, _, nl, _, _, info = measure(S2, 4, 3, torch.cuda.synchronize, None, e2e=False)
                    table[name] = {"workload": s2.describe(), "config_index": CONFIG_INDEX[name], "value": 4 / (dms / 1e3), "ms_per_step": dms / 4,
                                   "gpu_launches_per_step": nl / 4, "phase_ms": dict(zip(PHASES, [round(v, 4) for v in info.phase_ms]))}
                    S2.close()
                except Exception as e:      # a side table must never take the headline down
                    table[name] = {"error": str(e)}
            # opt-in fast path (SURVEY.md §8(f)2,4; NOT the graded dense path): the same MAX-CUT SDP read from SDPA-sparse text,
            # uploaded as triplets, Schur complement from the nonzero entries of the A_p (X^-1 o Y)
            if kind == "maxcut":
                try:
                    from clrs_b200 import sdpa, workloads as wl
                    t0 = time.perf_counter()
                    s2 = sdpa.sdpa_sparse_to_sdp(sdpa.maxcut_sdpa_text(wl.laplacian_random(args.n, 0.5, 0)), name=f"maxcut(n={args.n}) from SDPA-sparse text")
                    S2 = Solver(s2, lib="device", device=local_rank, duality_gap_threshold=GAP, sparse_schur=True)
                    torch.cuda.synchronize()
                    setup_s = time.perf_counter() - t0
                    _, dms, _, nl, _, _, info = measure(S2, 4, 3, torch.cuda.synchronize, None, e2e=False)
                    table["maxcut_sparse_schur"] = {"workload": s2.describe(), "config_index": 1, "value": 4 / (dms / 1e3), "ms_per_step": dms / 4,
                                                    "gpu_launches_per_step": nl / 4, "setup_seconds": setup_s,
                                                    "phase_ms": dict(zip(PHASES, [round(v, 4) for v in info.phase_ms])),
                                                    "note": "clrs_options.sparse_schur = 1 + clrs_add_sparse_term: opt-in shortcut, same optimum as the dense path (tests/test_gpu_sparse.py); the headline above is the dense GEMM path"}
                    S2.close()
                except Exception as e:
                    table["maxcut_sparse_schur"] = {"error": str(e)}
            out["configs"] = table
        # ---- time to duality gap 1e-30 (the second half of BASELINE.json's metric): a full solve from the default start ----
        if kind == "maxcut" and not args.no_time_to_gap:
            try:
                from clrs_b200 import solvesdp
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                full = solvesdp(sdp, lib="device", device=local_rank, gemm_path=args.gemm_path, duality_gap_threshold=1e-30)
                torch.cuda.synchronize()
                out["time_to_gap_1e-30"] = {"seconds": time.perf_counter() - t0, "iterations": full.iterations, "status": full.status,
                                            "note": "wall clock of solvesdp(...) through the C ABI incl. upload of the SDP and the per-iteration host round trip"}
            except Exception as e:
                out["time_to_gap_1e-30"] = {"seconds": None, "note": f"failed: {e}"}
        if not args.no_cpu_baseline:
            try:
                its, threads, desc = cpu_arm(args.n, 1, 1, kind=kind)
                out["cpu_baseline"] = {"value": its, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc, "host_cpus": os.cpu_count()}
            except Exception as e:      # the baseline must never take the GPU number down
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(out), file=real_stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — IPM iterations/sec of the B200 hot path on the BASELINE workload.

Workload (config.workload): BASELINE.json configs[1] — Goemans-Williamson MAX-CUT
relaxation, Laplacian of G(300, 0.5) (numpy default_rng(0)), every constraint matrix
E_pp stored dense (the reference's dense Schur path), prec = 256 bit.  One "step" is
one predictor-corrector IPM iteration (src/solver.jl:362-592) on the resident SDP.

  value   iterations/s from the device time of K consecutive clrs_iterate calls
          (CUDA events on the library's stream, problem and iterate resident in HBM)
  e2e     the same K iterations driven through the C ABI with HOST buffers every step:
          clrs_set_state (H2D of x,X,y,Y) + clrs_iterate + clrs_get_state (D2H)
  roofline  the tcgen05 slice-pair GEMM launches, canonical int8 op count
          2*M*N*K*528 / CUDA-event time, against 2x the measured bf16 peak
  cpu_baseline  the MPFR oracle (a port of the reference, not the reference) on the host cores,
          bounded sample: a few rows of the dense Schur path at full cost, the rest of the
          iteration measured in full, scaled to one iteration

--impl reference runs only that CPU arm.  N > 1 (torchrun): the workload is one cluster with
one block, it does not shard: every rank solves an independent replica.
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


def workload(n, kind="maxcut"):
    import clrs_b200
    from clrs_b200 import workloads
    from fractions import Fraction as F
    if kind == "sphere":        # BASELINE.json configs[4]: many clusters coupled through the free variables; shards by cluster
        return workloads.sphere_packing(8, n if n != 300 else 31, [F(1, 2), F(1, 2), F(3, 4), F(1)], prec=512)   # SURVEY.md §8(d) size (4,31): J=16, N=641, P=1294
    if kind == "threepoint":    # configs[3] at the size of examples/ThreePointBound.jl (one cluster: the dense F_k blocks are shared)
        return workloads.three_point_bound(4, F(1, 6), n if n != 300 else 10, n if n != 300 else 10, prec=256)
    return workloads.maxcut(workloads.laplacian_random(n, 0.5, 0), prec=256)


def config(n, n_gpus, kind="maxcut", sdp=None):
    if kind != "maxcut":
        return {"workload": f"{sdp.describe() if sdp is not None else kind} (BASELINE.json configs[{4 if kind == 'sphere' else 3}])",
                "step": "one predictor-corrector IPM iteration", "l2": "L2 flushed by the iteration's own temporaries only; latency-bound small blocks",
                "parallelism": ("clusters sharded over ranks (LPT), NCCL all-gather of Q, u, p and scalars" if kind == "sphere" else "replicas only") if n_gpus > 1 else "single GPU"}
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


def cpu_arm(n, steps, warmup, target_seconds=12.0, kind="maxcut"):
    """The oracle on the host cores, bounded sample.  Returns (it/s, description, seconds per step list)."""
    import ctypes as C
    from clrs_b200 import Solver
    import oracle.binding  # noqa: F401  (the one place bench.py touches oracle/: the CPU baseline / reference arm)
    sdp = workload(n, kind)
    S = Solver(sdp, lib="oracle", oracle_skip_zeros=True)
    lib = S.lib
    lib.clrs_oracle_set_sample_limit.restype = None
    lib.clrs_oracle_get_sample_times.restype = None
    cores = os.cpu_count() or 1
    limit = 1
    times = []
    desc = ""
    for it in range(warmup + steps):
        lib.clrs_oracle_set_sample_limit(S.h, C.c_int32(limit))
        t0 = time.perf_counter()
        S.iterate()
        wall = time.perf_counter() - t0
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
    return its, cores, desc


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
    ap.add_argument("--size", dest="n", type=int, default=300, help="graph size (300 = the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-time-to-gap", action="store_true", help="skip the full solve to gap 1e-30 (for runs under ncu)")
    ap.add_argument("--gemm-path", type=int, default=0)
    ap.add_argument("--workload", default="maxcut", choices=["maxcut", "sphere", "threepoint"],
                    help="maxcut = BASELINE configs[1] (default, the metric's config); sphere = configs[4], sharded by cluster when N > 1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        its, cores, desc = cpu_arm(args.n, max(1, args.steps), max(0, min(args.warmup, 1)), kind=args.workload)
        print(file=real_stdout, flush=True, *[json.dumps({"impl": "reference", "metric": METRIC, "value": its, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "mpfr (cpu)", "data": "synthetic", "config": config(args.n, args.gpus, args.workload, workload(args.n, args.workload) if args.workload != "maxcut" else None),
                          "cpu_baseline": {"value": its, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
                          "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})])
        return

    import torch
    import numpy as np
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
    from clrs_b200 import Solver, wire
    sdp = workload(args.n, args.workload)
    sharded = world > 1 and args.workload == "sphere"
    comm = None
    if sharded:                 # one communicator over the ranks; the id travels through torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(clrs_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        comm = (rank, world, bytes(uid.cpu().tolist()))
    S = Solver(sdp, lib="device", device=local_rank, gemm_path=args.gemm_path, comm=comm)
    reps = 1 if sharded else world
    W, K = max(args.warmup, 3), args.steps

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(W):
        S.iterate()
    # ---- device-resident throughput ----
    sampler = ClockSampler()
    if rank == 0:
        sampler.start(local_rank)
    S.profile(False)            # resets the launch counter
    sync()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(K):
        S.iterate()
        dev_ms += S.last_iteration_ms()
    sync()
    wall = time.perf_counter() - t0
    launches = S.profile_get()["kernel_launches"]
    clocks = sampler.stop(local_rank) if rank == 0 else None
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = reps * K / (dev_ms_max / 1e3)

    # ---- end to end through the C ABI with host buffers every step ----
    x, X, y, Y = S.get_state()
    h2d = sum(a.nbytes for a in (x, X, y, Y) if a is not None)
    sync()
    t0 = time.perf_counter()
    for _ in range(K):
        S.set_state(x, X, y, Y)
        S.iterate()
        x, X, y, Y = S.get_state()
    sync()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = reps * K / float(te.item())

    # ---- roofline of the dominant kernel: per-launch CUDA-event timing of every GEMM ----
    S.profile(True)
    for _ in range(min(K, 2)):
        S.iterate()
    prof = S.profile_get()
    S.profile(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 dense = 2x bf16 rate)" if peaks else "2 x 1.4 PFLOP/s fallback of B200_PROFILING.md"
    classes = {"dp4a": "k_gemm_dp4a (CUDA-core int8 path, small shapes)",
               "tc_small": "tc::k_gemm_tc + k_tc_recombine, products with < 1e6 outputs (block-level n x n products, split-K Schur dots)",
               "tc_large": "tc::k_gemm_tc + k_tc_recombine, the two 90000 x 300 x 300 products of the dense Schur path (in chunks of 63 constraints = 3 waves of CTAs)"}
    def rl(c):
        ms, mpf, nl = prof[c]["ms"], prof[c]["mp_flops"], prof[c]["launches"]
        ach = (mpf * 528.0 / (ms / 1e3)) / 1e12 if ms > 0 else 0.0
        return {"kernel": classes[c], "achieved": ach, "frac": ach / (2 * bf16), "launches": nl, "avg_launch_ms": ms / max(1, nl),
                "share_of_step": ms / max(1e-9, min(K, 2) * dev_ms_max / K)}
    dom = max(classes, key=lambda c: prof[c]["ms"])
    r = rl(dom)
    roofline = {"bound": "tensor", "kernel": r["kernel"], "achieved": r["achieved"], "peak": 2 * bf16,
                "unit": "TOP/s (int8, canonical 2*M*N*K*528 per 256-bit GEMM)", "frac": r["frac"],
                "traffic": (2.93e9 * (2.0 * min(K, 2)) / max(1, r["launches"])) if (dom == "tc_large" and args.workload == "maxcut" and args.n == 300) else None,
                "traffic_note": "per launch: dram__bytes_read.sum + dram__bytes_write.sum of ONE 90000x300x300 launch (2.93e9; ncu --set full, profiles/r01_tc_kernel_ncu_full.txt; algorithmic 2.08e9) scaled by rows per launch - the two 90000-row products of an iteration run as wave-sized chunks of constraints, and traffic is proportional to rows (A slices in, byte planes out)",
                "peak_source": peak_src, "launches": r["launches"], "avg_launch_ms": r["avg_launch_ms"], "gemm_share_of_step": r["share_of_step"],
                "other_gemm_classes": {c: rl(c) for c in classes if c != dom}}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K,
           "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": f"int8 slices of {sdp.prec}-bit mantissas (exact int32 accumulation)",
           "data": "synthetic", "config": config(args.n, world, args.workload, sdp), "wall_ms_per_step": 1e3 * wall / K,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d,
                   "path": "clrs_set_state + clrs_iterate + clrs_get_state with host wire buffers"},
           "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
    # ---- time to duality gap 1e-30 (the second half of BASELINE.json's metric): a full solve from the default start ----
    if rank == 0 and world == 1 and args.workload == "maxcut" and not args.no_time_to_gap:
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
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            its, cores, desc = cpu_arm(args.n, 1, 1, kind=args.workload)
            out["cpu_baseline"] = {"value": its, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        except Exception as e:      # the baseline must never take the GPU number down
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    S.close()
    if rank == 0:
        print(json.dumps(out), file=real_stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

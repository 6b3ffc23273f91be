"""Kernel-only benchmark of the multi-precision GEMM paths (not a test)."""
import ctypes as C, sys, json
sys.path.insert(0, ".")
import clrs_b200
from clrs_b200 import workloads, Solver
S = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="device")
fn = S.lib.clrs_bench_gemm; fn.restype = C.c_int
PATH = int(sys.argv.pop(1)[1:]) if len(sys.argv) > 1 and sys.argv[1].startswith("p") else 2
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(90000, 300, 300), (16384, 128, 512), (4096, 256, 256), (300, 300, 3584)]
for (M, N, K) in shapes:
    out = (C.c_double * 3)()
    rc = fn(S.h, M, N, K, 3, PATH, out)
    ops = 2.0 * M * N * K * 528
    print(json.dumps({"path": PATH, "M": M, "N": N, "K": K, "rc": rc, "split_ms": round(out[0], 3), "gemm_ms": round(out[1], 3), "kernel_ms": round(out[2], 3),
                      "canonical_int8_TOPS_kernel": round(ops / (out[2] / 1e3) / 1e12, 1) if out[2] else None,
                      "canonical_int8_TOPS_gemm": round(ops / (out[1] / 1e3) / 1e12, 1)}), flush=True)

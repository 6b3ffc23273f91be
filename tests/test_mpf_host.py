"""The device arithmetic (csrc/mpf.cuh, csrc/i8split.cuh) compiled for the host and checked
against mpmath / the MPFR oracle.  The kernels compile exactly these sources."""
import ctypes as C
import os
import random

import mpmath
import numpy as np
import pytest

import clrs_b200
from clrs_b200 import wire, workloads, Solver

PREC = 256


@pytest.fixture(scope="module")
def hc():
    return C.CDLL(os.path.join(os.path.dirname(clrs_b200.DEVICE_LIB), "libclrs_hostcheck.so"))


def rnd(rng):
    if rng.random() < 0.05:
        return mpmath.mpf(0)
    m = mpmath.mpf(rng.getrandbits(300)) / 2 ** 300 + mpmath.mpf(1) / 7
    e = rng.randint(-300, 300) if rng.random() < 0.7 else rng.randint(-3, 3)
    return rng.choice([1, -1]) * m * mpmath.mpf(2) ** e


def test_add_sub_mul_div_within_one_ulp(hc):
    rng = random.Random(1)
    with mpmath.workprec(400):
        for it in range(3000):
            a, b = rnd(rng), rnd(rng)
            if it % 5 == 0:
                b = a * (1 + mpmath.mpf(2) ** -rng.randint(1, 250)) * rng.choice([1, -1])     # cancellation
            wa, wb = wire.to_wire(a, PREC), wire.to_wire(b, PREC)
            ra, rb = wire.from_wire(wa, PREC), wire.from_wire(wb, PREC)
            for op, ex in ((0, ra + rb), (1, ra - rb), (2, ra * rb), (3, None)):
                if op == 3:
                    if rb == 0:
                        continue
                    ex = ra / rb
                r = wire.wire_zeros((), PREC)
                hc.hc_binop(op, wa.ctypes.data_as(C.c_void_p), wb.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
                got = wire.from_wire(r, PREC)
                if ex == 0:
                    assert got == 0
                else:
                    assert abs(got - ex) <= abs(ex) * mpmath.mpf(2) ** -254, (op, a, b)


def test_recip_sqrt_rsqrt(hc):
    rng = random.Random(2)
    with mpmath.workprec(400):
        for _ in range(1000):
            a = abs(rnd(rng))
            if a == 0:
                continue
            wa = wire.to_wire(a, PREC); ra = wire.from_wire(wa, PREC)
            for op, ex in ((0, 1 / ra), (1, mpmath.sqrt(ra)), (2, 1 / mpmath.sqrt(ra))):
                r = wire.wire_zeros((), PREC)
                hc.hc_unop(op, wa.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
                assert abs(wire.from_wire(r, PREC) - ex) <= abs(ex) * mpmath.mpf(2) ** -253


def test_double_conversion_and_compare(hc):
    hc.hc_to_double.restype = C.c_double
    for v in (0.0, 1.0, -3.5, 1e-300, 2.0 ** 70, -1e10):
        r = wire.wire_zeros((), PREC)
        hc.hc_from_double(C.c_double(v), r.ctypes.data_as(C.c_void_p))
        assert hc.hc_to_double(r.ctypes.data_as(C.c_void_p)) == v
    with mpmath.workprec(300):
        a, b = wire.to_wire(mpmath.mpf(1) / 3, PREC), wire.to_wire(mpmath.mpf(1) / 3 + mpmath.mpf(2) ** -250, PREC)
    assert hc.hc_cmp(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == -1
    assert hc.hc_cmp(b.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)) == 1
    assert hc.hc_cmp(a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)) == 0


@pytest.mark.parametrize("M,N,K,spread", [(3, 4, 5, 0), (4, 3, 17, 40), (2, 2, 33, 300), (5, 5, 8, 2), (1, 1, 1, 0)])
def test_int8_slice_gemm_matches_oracle_normwise(hc, M, N, K, spread):
    """split -> exact int32 slice-pair sums -> recombine: error <= K 2^-250 rowmax colmax (empty/zero rows included)."""
    rng = random.Random(M + N + K)
    with mpmath.workprec(600):
        A = [[rnd(rng) if spread else mpmath.mpf(rng.randint(-5, 5)) for _ in range(K)] for _ in range(M)]
        B = [[rnd(rng) if spread else mpmath.mpf(rng.randint(-5, 5)) for _ in range(N)] for _ in range(K)]
        wa, wb = wire.to_wire(A, PREC), wire.to_wire(B, PREC)
        wc = wire.wire_zeros((M, N), PREC)
        hc.hc_gemm(M, N, K, wa.ctypes.data_as(C.c_void_p), wb.ctypes.data_as(C.c_void_p), wc.ctypes.data_as(C.c_void_p))
        o = Solver(workloads.maxcut(workloads.laplacian_cycle(3)), lib="oracle")
        wo, _ = o.mp_gemm(wa, wb)
        a, b, c, co = (wire.from_wire(v, PREC) for v in (wa, wb, wc, wo))
        for i in range(M):
            for j in range(N):
                scale = max(abs(v) for v in a[i, :]) * max(abs(v) for v in b[:, j])
                assert abs(c[i, j] - co[i, j]) <= scale * K * mpmath.mpf(2) ** -250
        o.close()


def test_wire_round_trip():
    with mpmath.workprec(300):
        vals = [mpmath.mpf(1) / 3, -mpmath.pi * 10 ** 40, 0, 5, mpmath.mpf(2) ** -700]
        w = wire.to_wire(vals, PREC)
        back = wire.from_wire(w, PREC)
        for v, b in zip(vals, back):
            with mpmath.workprec(PREC):
                assert b == +mpmath.mpf(v)
    assert wire.wire_dtype(256).itemsize == 16 + 8 * 4 and wire.wire_dtype(300).itemsize == 16 + 8 * 5


def test_wire_records_are_top_aligned_for_any_limb_count():
    """A wire record carries ceil(prec/64) 64-bit limbs; the device number has 8, 10 or 16 32-bit limbs.  Shorter records
    fill the upper limbs, longer ones are truncated, and the value survives (exactly when nothing is cut)."""
    import ctypes as C
    import mpmath
    from clrs_b200 import wire
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clusteredlowranksolver.jl_b200", "csrc", "libclrs_hostcheck.so"))
    with mpmath.workprec(600):
        v = -mpmath.mpf(2) ** 77 / 3
        for prec_in, prec_out in ((128, 128), (128, 256), (256, 128), (256, 256), (384, 256), (256, 512)):
            a = wire.to_wire([v], prec_in)
            out = wire.wire_zeros((1,), prec_out)
            assert lib.hc_wire_convert(C.c_int((prec_in + 63) // 64), a.ctypes.data_as(C.c_void_p), C.c_int((prec_out + 63) // 64), out.ctypes.data_as(C.c_void_p)) == -1
            back = wire.from_wire(out, prec_out)[0]
            kept = min(prec_in, prec_out, 256)
            assert abs(back - v) <= abs(v) * mpmath.mpf(2) ** (-(kept - 2)), (prec_in, prec_out)

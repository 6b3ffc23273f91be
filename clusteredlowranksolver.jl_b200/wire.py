"""Wire-number conversions (include/clrs_b200.h "Wire number format").

One multi-precision value = a little-endian record of 16 + 8*W bytes,
W = ceil(prec/64): int64 exp, int32 sign (-1/0/+1), int32 reserved, uint64
limb[W] (limb[W-1] most significant, top bit set) — the (sign, exp, d) triple
of an MPFR / Julia BigFloat at `prec` bits, which is what the reference hands
out at src/solver.jl:747-750.  Host-side values are mpmath `mpf`.
"""
from __future__ import annotations

import numpy as np
import mpmath
from mpmath.libmp import normalize, MPZ

__all__ = ["wire_dtype", "limbs_for", "to_wire", "from_wire", "wire_zeros", "wire_eye_scaled"]


def limbs_for(prec: int) -> int:
    return (int(prec) + 63) // 64


def wire_dtype(prec: int) -> np.dtype:
    return np.dtype([("exp", "<i8"), ("sign", "<i4"), ("pad", "<i4"), ("limb", "<u8", (limbs_for(prec),))])


def wire_zeros(shape, prec: int) -> np.ndarray:
    return np.zeros(shape, dtype=wire_dtype(prec))


def _raw(v, prec: int):
    """(sign, man, exp, bc) of v rounded to `prec` bits (round to nearest)."""
    if isinstance(v, mpmath.mpf):
        s, man, e, bc = v._mpf_
    else:
        with mpmath.workprec(prec + 64):
            s, man, e, bc = mpmath.mpf(v)._mpf_
    return normalize(s, MPZ(man), e, bc, prec, "n")


def _store(rec, v, prec: int, W: int):
    s, man, e, bc = _raw(v, prec)
    if man == 0:
        rec["sign"] = 0
        return
    top = int(man) << (64 * W - bc)
    rec["exp"] = e + bc
    rec["sign"] = -1 if s else 1
    limbs = rec["limb"]
    for k in range(W):
        limbs[k] = (top >> (64 * k)) & 0xFFFFFFFFFFFFFFFF


def to_wire(values, prec: int) -> np.ndarray:
    """Convert a scalar / nested sequence / mpmath matrix to a wire array of the same shape."""
    W = limbs_for(prec)
    if isinstance(values, mpmath.matrix):
        out = wire_zeros((values.rows, values.cols), prec)
        for i in range(values.rows):
            for j in range(values.cols):
                _store(out[i, j], values[i, j], prec, W)
        return out
    if isinstance(values, np.ndarray) and values.dtype == wire_dtype(prec):
        return values
    arr = np.asarray(values, dtype=object)
    out = wire_zeros(arr.shape, prec)
    if arr.shape == ():
        _store(out[()], arr[()], prec, W)
        return out
    flat_out = out.reshape(-1)
    for idx, v in enumerate(arr.reshape(-1)):
        if v == 0:
            continue
        _store(flat_out[idx], v, prec, W)
    return out


def from_wire(arr: np.ndarray, prec: int):
    """Wire array -> object ndarray of mpf (same shape); 0-d input gives an mpf."""
    W = limbs_for(prec)
    a = np.asarray(arr)
    flat = a.reshape(-1)
    out = np.empty(flat.shape, dtype=object)
    with mpmath.workprec(64 * W):
        for i in range(flat.shape[0]):
            rec = flat[i]
            sg = int(rec["sign"])
            if sg == 0:
                out[i] = mpmath.mpf(0)
                continue
            man = 0
            limbs = rec["limb"]
            for k in range(W):
                man |= int(limbs[k]) << (64 * k)
            out[i] = mpmath.mpf((1 if sg < 0 else 0, man, int(rec["exp"]) - 64 * W, man.bit_length()))
    if a.shape == ():
        return out[0]
    return out.reshape(a.shape)


def wire_eye_scaled(n: int, value, prec: int) -> np.ndarray:
    out = wire_zeros((n, n), prec)
    one = to_wire(value, prec)
    for i in range(n):
        out[i, i] = one[()]
    return out

// lanes.cuh — cross-rank sums of multi-limb numbers with NCCL's own reductions.
//
// NCCL has no multi-limb type.  A sum over ranks is done in fixed point:
//   (1) all ranks agree on one exponent per entry, E = max over ranks (ncclMax on int32; LANE_EXP_NONE marks a zero);
//   (2) every rank writes its value relative to E into NL + 1 lanes of 32 payload bits (one guard lane below the
//       mantissa), each lane a signed int64;
//   (3) ncclSum on int64 adds the lanes exactly (|lane| < 2^32: headroom for 2^31 ranks) and in any order, so every
//       rank ends up with the same bits;
//   (4) the carries are resolved and the sum is normalised back into a multi-limb number.
// Bits below the guard lane are truncated, as in a floating-point sum of the aligned operands.
// These two functions are the whole arithmetic of the scheme; they are __host__ __device__ so that the CPU tests run
// the protocol with gloo all-reduces (tests/test_multirank_cpu.py).
#pragma once
#include "mpf.cuh"

static const int32_t LANE_EXP_NONE = -(1 << 30);
template <int NL> HD int32_t mp_lane_exp(const mpn<NL>& a) { return a.sign ? a.exp : LANE_EXP_NONE; }

// lanes[k], k = 0 (guard lane, least significant) .. NL: a relative to exponent e >= a.exp
template <int NL> HD void mp_to_lanes(const mpn<NL>& a, int32_t e, long long (&lanes)[NL + 1]) {
  constexpr int LN = NL + 1;
  uint32_t w[LN];
#pragma unroll
  for (int k = 0; k < LN; k++) w[k] = (k == 0) ? 0u : a.l[(k > 0) ? k - 1 : 0];
  const long long d = (long long)e - a.exp;
  if (a.sign == 0 || e == LANE_EXP_NONE || d >= 32 * LN) {
#pragma unroll
    for (int k = 0; k < LN; k++) w[k] = 0u;
  } else { limbs_shr_words<LN>(w, (int)(d >> 5)); limbs_shr_bits<LN>(w, (int)(d & 31)); }
#pragma unroll
  for (int k = 0; k < LN; k++) lanes[k] = a.sign < 0 ? -(long long)w[k] : (long long)w[k];
}
// the summed lanes (any values that fit int64) back to a number
template <int NL> HD void mp_from_lanes(mpn<NL>& r, const long long (&lanes)[NL + 1], int32_t e) {
  constexpr int LN = NL + 1, NW = LN + 2;
  mp_zero(r);
  if (e == LANE_EXP_NONE) return;
  uint32_t w[NW]; long long carry = 0;
#pragma unroll
  for (int k = 0; k < LN; k++) { const long long t = lanes[k] + carry; w[k] = (uint32_t)(t & 0xffffffffll); carry = t >> 32; }
  w[LN] = (uint32_t)(carry & 0xffffffffll); w[LN + 1] = (uint32_t)((carry >> 32) & 0xffffffffll);
  const bool neg = carry < 0;
  if (neg) { uint32_t c = 1;
#pragma unroll
    for (int k = 0; k < NW; k++) { const uint64_t t = (uint64_t)(~w[k]) + c; w[k] = (uint32_t)t; c = (uint32_t)(t >> 32); } }
  const int sh = limbs_normalize<NW>(w);
  if (sh < 0) return;
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = w[NW - NL + k];
  r.exp = e + 64 - sh; r.sign = neg ? -1 : 1;
}

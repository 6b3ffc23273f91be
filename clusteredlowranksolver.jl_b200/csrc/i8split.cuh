// i8split.cuh — Ozaki-style split of multi-limb numbers into int8 slices and the
// exact recombination of the int32 slice-pair sums (north star subsystem 1).
//
// A vector (a row of the left operand / a column of the right operand) shares
// one exponent E = max exponent of its entries.  Each entry is turned into the
// fixed-point integer  I = trunc(a * 2^(B - E)),  B = 8*NS - 2,  |I| < 2^B, and
// written in balanced radix 256:  I = sum_t d_t 256^(NS-1-t),  d_t in [-128,127],
// |d_0| <= 64.  Digit t (t = 0 most significant) has weight 2^(-8t-6) 2^E.
//
// For C = A*B the slice-pair sums  D_s = sum_{t+u=s} sum_k d_t(a_ik) d_u(b_kj)
// are exact in int32 (|d d'| <= 2^14), and
//     c_ij = 2^(E_i + F_j - 12) * sum_{s=0}^{NS-1} D_s 2^(-8 s)
// up to the truncated pairs s >= NS:  error <= K (NS+1) 2^(-8 NS + 2) relative to
// rowmax*colmax, i.e. K * 2^-256.. for NS = 4*NL + 3 (35 slices at 256 bit).
#pragma once
#include "mpf.cuh"

template <int NL> struct I8Cfg {
  static constexpr int NS = 4 * NL + 3;                   // slices per number
  static constexpr int NSP = (NS + 3) & ~3;               // padded to a multiple of 4 words
  static constexpr int NW = (8 * NS + 31) / 32;           // words holding the fixed-point integer
  static constexpr int LSH = 8 * NS - 2 - 32 * NL;        // = 22: left shift placing the mantissa at 2^B
  static constexpr int NPAIRS = NS * (NS + 1) / 2;
};
static const int32_t I8_EXP_NONE = -(1 << 29);           // exponent of an all-zero vector

// digits of one entry relative to the vector exponent E; dig[t], t = 0 most significant
template <int NL> HD void i8_split(const mpn<NL>& a, int32_t E, int8_t (&dig)[I8Cfg<NL>::NS]) {
  constexpr int NS = I8Cfg<NL>::NS, NW = I8Cfg<NL>::NW, LSH = I8Cfg<NL>::LSH;
  uint32_t w[NW];
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = (i < NL) ? a.l[(i < NL) ? i : 0] : 0u;
  uint32_t d = (uint32_t)(E - a.exp);
  if (a.sign == 0 || d >= (uint32_t)(8 * NS)) {
#pragma unroll
    for (int t = 0; t < NS; t++) dig[t] = 0;
    return;
  }
  limbs_shl_bits<NW>(w, LSH);                             // LSH < 32
  limbs_shr_words<NW>(w, (int)(d >> 5)); limbs_shr_bits<NW>(w, (int)(d & 31));
  if (a.sign < 0) {                                       // two's complement
    uint32_t c = 1;
#pragma unroll
    for (int i = 0; i < NW; i++) { uint64_t t = (uint64_t)(~w[i]) + c; w[i] = (uint32_t)t; c = (uint32_t)(t >> 32); }
  }
  int carry = 0;
#pragma unroll
  for (int t = 0; t < NS; t++) {                          // t counts from the least significant byte
    int u = (int)((w[t / 4] >> (8 * (t % 4))) & 255u);
    int v;
    if (t < NS - 1) { v = u + carry; carry = v >= 128; v -= carry << 8; }
    else v = (int)(int8_t)u + carry;
    dig[NS - 1 - t] = (int8_t)v;
  }
}

// Exact recombination.  dg[s], s = 1..NS-1, are the carry-normalised low digits
// in [0,255]; `top` is the (signed) integer part in units of slice-pair sum 0.
// Esum = E_i + F_j.  Result truncated to NL limbs.
template <int NL> HD void i8_recombine(mpn<NL>& r, int64_t top, const uint32_t (&dg)[I8Cfg<NL>::NS], int32_t Esum) {
  constexpr int NS = I8Cfg<NL>::NS, NB = NS - 1, NW = (8 * NB + 64 + 31) / 32, WI = (8 * NB) / 32, SH = (8 * NB) % 32;
  uint32_t w[NW];
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = 0;
#pragma unroll
  for (int s = 1; s < NS; s++) { const int b = NS - 1 - s; w[b / 4] |= dg[s] << (8 * (b % 4)); }
#pragma unroll
  for (int j = 0; j < NW - WI; j++) {
    uint32_t piece;
    if (j == 0) piece = (uint32_t)((uint64_t)top << SH);
    else { const int shr = 32 * j - SH; piece = (uint32_t)(top >> (shr > 63 ? 63 : shr)); }
    w[WI + j] |= piece;
  }
  const bool neg = top < 0;
  if (neg) {
    uint32_t c = 1;
#pragma unroll
    for (int i = 0; i < NW; i++) { uint64_t t = (uint64_t)(~w[i]) + c; w[i] = (uint32_t)t; c = (uint32_t)(t >> 32); }
  }
  int sh = limbs_normalize<NW>(w);
  if (sh < 0) { mp_zero(r); return; }
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = w[NW - NL + i];
  r.exp = 32 * NW - sh - 8 * (NS - 1) - 12 + Esum;
  r.sign = neg ? -1 : 1;
}

// In-register carry normalisation of the slice-pair sums: afterwards acc[s] in
// [0,255] for s >= 1, acc[0] = 0 and `top` has absorbed everything above.
template <int NS> HD void i8_carry_normalize(int32_t (&acc)[NS], int64_t& top) {
#pragma unroll
  for (int s = NS - 1; s >= 1; s--) { int32_t c = acc[s] >> 8; acc[s] -= c << 8; acc[s - 1] += c; }
  top += acc[0]; acc[0] = 0;
}

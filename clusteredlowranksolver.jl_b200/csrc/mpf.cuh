// mpf.cuh — fixed multi-limb binary floating point for the device.
//
// Replaces the arb/arf midpoint arithmetic the reference reaches through
// Arblib.jl: a number is a sign, a 32-bit binary
// exponent and an NL x 32-bit integer mantissa (NL = 8 for the solver's default
// 256 bits), normalised so the top mantissa bit is set.  40 bytes at 256 bit —
// the per-number figure of SURVEY.md §8(d).
//
//   value = sign * (sum_i l[i] 2^(32 i)) / 2^(32 NL) * 2^exp ,   sign in {-1,0,+1}
//
// Every routine is __host__ __device__ so the identical code is unit-tested on
// the CPU against mpmath (tests/test_mpf_host.py) and runs inside the kernels.
// All register arrays are indexed with compile-time constants only (variable
// shifts are log-shifters), so nothing spills to local memory.
// Rounding: truncation (error < 1 ulp per operation; Arb's approx_* kernels
// give no last-bit guarantee either).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

template <int NL>
struct alignas(8) mpn {
  uint32_t l[NL];
  int32_t exp;
  int32_t sign;
};

HD int mp_clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
// low 32 bits of (hi:lo) >> b, b in [0,31]
HD uint32_t mp_fshr(uint32_t lo, uint32_t hi, int b) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, b);
#else
  return b ? (lo >> b) | (hi << (32 - b)) : lo;
#endif
}
// high 32 bits of (hi:lo) << b, b in [0,31]
HD uint32_t mp_fshl(uint32_t lo, uint32_t hi, int b) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, b);
#else
  return b ? (hi << b) | (lo >> (32 - b)) : hi;
#endif
}

// ---- raw limb-array helpers (N compile-time) -------------------------------
template <int N> HD void limbs_shr_words(uint32_t (&x)[N], int s) {   // x >>= 32*s, 0 <= s
#pragma unroll
  for (int st = 1; st < 2 * N; st <<= 1)
    if (s & st) {
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = (i + st < N) ? x[(i + st < N) ? i + st : 0] : 0u;
    }
}
template <int N> HD void limbs_shl_words(uint32_t (&x)[N], int s) {   // x <<= 32*s, 0 <= s
#pragma unroll
  for (int st = 1; st < 2 * N; st <<= 1)
    if (s & st) {
#pragma unroll
      for (int i = N - 1; i >= 0; i--) x[i] = (i - st >= 0) ? x[(i - st >= 0) ? i - st : 0] : 0u;
    }
}
template <int N> HD void limbs_shr_bits(uint32_t (&x)[N], int b) {    // b in [0,31]
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = mp_fshr(x[i], (i + 1 < N) ? x[(i + 1 < N) ? i + 1 : 0] : 0u, b);
}
template <int N> HD void limbs_shl_bits(uint32_t (&x)[N], int b) {    // b in [0,31]
#pragma unroll
  for (int i = N - 1; i >= 0; i--) x[i] = mp_fshl((i > 0) ? x[(i > 0) ? i - 1 : 0] : 0u, x[i], b);
}
template <int N> HD uint32_t limbs_add(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N]) {
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { uint64_t t = (uint64_t)a[i] + b[i] + c; r[i] = (uint32_t)t; c = (uint32_t)(t >> 32); }
  return c;
}
template <int N> HD uint32_t limbs_sub(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N]) {
  uint32_t brw = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { uint64_t t = (uint64_t)a[i] - b[i] - brw; r[i] = (uint32_t)t; brw = (uint32_t)(t >> 63); }
  return brw;
}
template <int N> HD int limbs_cmp(const uint32_t (&a)[N], const uint32_t (&b)[N]) {
  int r = 0;
#pragma unroll
  for (int i = 0; i < N; i++) r = (a[i] > b[i]) ? 1 : ((a[i] < b[i]) ? -1 : r);   // higher limbs decide last
  return r;
}
// Normalise a magnitude held in x (N words) so its top bit is set; returns the left
// shift applied (in bits), or -1 when x == 0.
template <int N> HD int limbs_normalize(uint32_t (&x)[N]) {
  int lzw = 0; bool seen = false;
#pragma unroll
  for (int i = N - 1; i >= 0; i--) { if (!seen) { if (x[i] == 0) lzw++; else seen = true; } }
  if (!seen) return -1;
  limbs_shl_words<N>(x, lzw);
  int b = mp_clz32(x[N - 1]);
  limbs_shl_bits<N>(x, b);
  return 32 * lzw + b;
}

// ---- basic value helpers --------------------------------------------------
template <int NL> HD void mp_zero(mpn<NL>& r) {
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = 0;
  r.exp = 0; r.sign = 0;
}
template <int NL> HD void mp_set_i32(mpn<NL>& r, int32_t v) {
  mp_zero(r); if (v == 0) return;
  uint32_t m = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v; int z = mp_clz32(m);
  r.l[NL - 1] = m << z; r.exp = 32 - z; r.sign = v < 0 ? -1 : 1;
}
// exact power of two as a double, -1022 <= k <= 1023 (the seeds of the Newton iterations scale by these instead of calling ldexp:
// the library calls dominated the ~2200 cycles the seed of one reciprocal square root cost, measured with k_bench_wops)
HD double mp_pow2(int k) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((1023 + k) << 20, 0);
#else
  return ldexp(1.0, k);
#endif
}
template <int NL> HD void mp_from_double(mpn<NL>& r, double d) {
  mp_zero(r); if (d == 0.0 || !(d == d)) return;
  uint64_t bits; memcpy(&bits, &d, 8);
  const int ef = (int)((bits >> 52) & 0x7ff);
  if (ef != 0 && ef != 0x7ff) {                           // normal number: the 53-bit significand sits at the top of 64 bits
    const uint64_t mant = ((bits & 0xFFFFFFFFFFFFFull) | (1ull << 52)) << 11;
    r.l[NL - 1] = (uint32_t)(mant >> 32); r.l[NL - 2] = (uint32_t)mant; r.exp = ef - 1022; r.sign = d < 0 ? -1 : 1; return; }
  int e; double m = frexp(fabs(d), &e);                  // m in [0.5,1)
  uint64_t mant = (uint64_t)ldexp(m, 64);                // exact: 53 significant bits
  r.l[NL - 1] = (uint32_t)(mant >> 32); r.l[NL - 2] = (uint32_t)mant; r.exp = e; r.sign = d < 0 ? -1 : 1;
}
template <int NL> HD double mp_to_double(const mpn<NL>& a) {
  if (a.sign == 0) return 0.0;
  uint64_t top = ((uint64_t)a.l[NL - 1] << 32) | a.l[NL - 2];
  int e = a.exp; if (e > 2000) e = 2000; if (e < -2000) e = -2000;
  double v = ldexp((double)top, e - 64);
  return a.sign < 0 ? -v : v;
}
template <int NL> HD int mp_cmp_abs(const mpn<NL>& a, const mpn<NL>& b) {
  if (a.sign == 0 || b.sign == 0) return (a.sign != 0) - (b.sign != 0);
  if (a.exp != b.exp) return a.exp > b.exp ? 1 : -1;
  return limbs_cmp<NL>(a.l, b.l);
}
template <int NL> HD int mp_cmp(const mpn<NL>& a, const mpn<NL>& b) {
  if (a.sign != b.sign) return a.sign > b.sign ? 1 : -1;
  if (a.sign == 0) return 0;
  int c = mp_cmp_abs(a, b); return a.sign > 0 ? c : -c;
}

// ---- multiplication -------------------------------------------------------
template <int NL> HD void mp_mul(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) {
  if (a.sign == 0 || b.sign == 0) { mp_zero(r); return; }
  uint32_t p[2 * NL];
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < 2 * NL - 1; k++) {
    uint64_t lo = carry, hi = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
      if (k - i >= 0 && k - i < NL) { uint64_t t = (uint64_t)a.l[i] * b.l[k - i]; lo += (uint32_t)t; hi += t >> 32; }
    }
    p[k] = (uint32_t)lo; carry = (lo >> 32) + hi;
  }
  p[2 * NL - 1] = (uint32_t)carry;
  int32_t e = a.exp + b.exp; int32_t s = a.sign * b.sign;
  const bool top = (p[2 * NL - 1] >> 31) != 0;           // product of two [1/2,1) mantissas is in [1/4,1)
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = top ? p[NL + i] : mp_fshl(p[NL + i - 1], p[NL + i], 1);
  r.exp = top ? e : e - 1; r.sign = s;
}

// ---- addition / subtraction ----------------------------------------------
template <int NL> HD void mp_add_signed(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b, int bsign_mul) {
  const int bs = b.sign * bsign_mul;
  if (bs == 0) { r = a; return; }
  if (a.sign == 0) { r = b; r.sign = bs; return; }
  // order by magnitude (exponent, then mantissa)
  bool swap = false;
  if (a.exp < b.exp) swap = true; else if (a.exp == b.exp && limbs_cmp<NL>(a.l, b.l) < 0) swap = true;
  const mpn<NL>& hi = swap ? b : a; const mpn<NL>& lo = swap ? a : b;
  const int his = swap ? bs : a.sign, los = swap ? a.sign : bs;
  const uint32_t d = (uint32_t)(hi.exp - lo.exp);
  if (d >= (uint32_t)(32 * NL + 32)) { r = hi; r.sign = his; return; }
  uint32_t x[NL + 1], y[NL + 1];                          // one guard limb below
  x[0] = 0; y[0] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) { x[i + 1] = hi.l[i]; y[i + 1] = lo.l[i]; }
  limbs_shr_words<NL + 1>(y, (int)(d >> 5)); limbs_shr_bits<NL + 1>(y, (int)(d & 31));
  int32_t e = hi.exp;
  if (his == los) {
    uint32_t c = limbs_add<NL + 1>(x, x, y);
    if (c) { limbs_shr_bits<NL + 1>(x, 1); x[NL] |= 0x80000000u; e += 1; }
  } else {
    limbs_sub<NL + 1>(x, x, y);
    int sh = limbs_normalize<NL + 1>(x);
    if (sh < 0) { mp_zero(r); return; }
    e -= sh;
  }
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = x[i + 1];
  r.exp = e; r.sign = his;
}
template <int NL> HD void mp_add(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_add_signed(t, a, b, 1); r = t; }
template <int NL> HD void mp_sub(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_add_signed(t, a, b, -1); r = t; }
template <int NL> HD void mp_neg(mpn<NL>& r) { r.sign = -r.sign; }
// r += a*b ;  r -= a*b
template <int NL> HD void mp_addmul(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_mul(t, a, b); mp_add(r, r, t); }
template <int NL> HD void mp_submul(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) { mpn<NL> t; mp_mul(t, a, b); mp_sub(r, r, t); }

// ---- reciprocal, division, (inverse) square root: Newton from a double-double seed ----
// The seed is one Newton step carried out in double-double arithmetic on the top 96 mantissa bits (~95 correct bits
// for a dozen double operations), which saves one full-precision Newton step (two or three multi-limb products on the
// sequential pivot chain of the Cholesky panels).  The double operations are written with explicit roundings
// (__dmul_rn / __fma_rn on the device, no contraction on the host) so host and device produce the same seed.
#ifdef __CUDA_ARCH__
#define CLRS_DMUL(a, b) __dmul_rn((a), (b))
#define CLRS_DADD(a, b) __dadd_rn((a), (b))
#define CLRS_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define CLRS_DMUL(a, b) ((a) * (b))
#define CLRS_DADD(a, b) ((a) + (b))
#define CLRS_DFMA(a, b, c) fma((a), (b), (c))
#endif
template <int NL> HD constexpr int mp_newton_steps() { return 32 * NL <= 160 ? 1 : (32 * NL <= 352 ? 2 : (32 * NL <= 704 ? 3 : 4)); }
// top 96 bits of a mantissa given by its three leading limbs, times 2^sc, as an exact sum hi + lo of two doubles
HD void mp_mant_dd(uint32_t A, uint32_t B, uint32_t C, int sc, double& hi, double& lo) {
  hi = (double)A * mp_pow2(sc - 32) + (double)(B >> 11) * mp_pow2(sc - 53);      // 32 + 21 = 53 bits: exact
  lo = (double)(B & 2047u) * mp_pow2(sc - 64) + (double)C * mp_pow2(sc - 96);    // 11 + 32 = 43 bits: exact
}
// y0 + c = m^(-1/2) (1 + O(2^-95)) for m = mh + ml in [1/4, 1)
HD void dd_rsqrt_seed(double mh, double ml, double& y0, double& c) {
#if defined(__CUDA_ARCH__)
  y0 = rsqrt(mh);                                                          // (any ~52-bit start: the correction below brings y0 + c to ~2^-95)
#else
  y0 = 1.0 / sqrt(mh);
#endif
  const double ph = CLRS_DMUL(y0, y0), pl = CLRS_DFMA(y0, y0, -ph);          // y0^2 = ph + pl
  const double th = CLRS_DMUL(mh, ph), tl0 = CLRS_DFMA(mh, ph, -th);         // mh ph = th + tl0
  const double tl = CLRS_DADD(tl0, CLRS_DADD(CLRS_DMUL(mh, pl), CLRS_DMUL(ml, ph)));
  const double e = CLRS_DADD(CLRS_DADD(1.0, -th), -tl);                      // 1 - m y0^2 (1 - th is exact)
  c = CLRS_DMUL(CLRS_DMUL(0.5, y0), e);
}
// x0 + c = 1/m (1 + O(2^-95)) for m = mh + ml in [1/2, 1)
HD void dd_recip_seed(double mh, double ml, double& x0, double& c) {
  x0 = 1.0 / mh;
  const double th = CLRS_DMUL(mh, x0), tl0 = CLRS_DFMA(mh, x0, -th);
  const double tl = CLRS_DADD(tl0, CLRS_DMUL(ml, x0));
  const double e = CLRS_DADD(CLRS_DADD(1.0, -th), -tl);
  c = CLRS_DMUL(x0, e);
}

template <int NL> HD void mp_recip(mpn<NL>& r, const mpn<NL>& a) {
  if (a.sign == 0) { mp_zero(r); return; }
  mpn<NL> m = a; m.exp = 0; m.sign = 1;                   // mantissa in [1/2,1)
  double mh, ml, x0, c; mp_mant_dd(a.l[NL - 1], a.l[NL - 2], NL >= 3 ? a.l[NL >= 3 ? NL - 3 : 0] : 0u, 0, mh, ml); dd_recip_seed(mh, ml, x0, c);
  mpn<NL> x, t, two; mp_from_double(x, x0); mp_from_double(t, c); mp_add(x, x, t); mp_set_i32(two, 2);
#pragma unroll 1
  for (int it = 0; it < mp_newton_steps<NL>(); it++) { mp_mul(t, m, x); mp_sub(t, two, t); mp_mul(x, x, t); }
  x.exp -= a.exp; x.sign = a.sign; r = x;
}
template <int NL> HD void mp_div(mpn<NL>& r, const mpn<NL>& a, const mpn<NL>& b) {
  mpn<NL> x, q, t; mp_recip(x, b); mp_mul(q, a, x);
  mp_mul(t, q, b); mp_sub(t, a, t); mp_mul(t, t, x); mp_add(r, q, t);   // one correction step: q += (a - q b) x
}
// r = a^(-1/2) for a > 0
template <int NL> HD void mp_rsqrt(mpn<NL>& r, const mpn<NL>& a) {
  mpn<NL> m = a; const int odd = a.exp & 1; m.exp = -odd; m.sign = 1;   // m in [1/4,1), a = m 2^(exp+odd), exponent even
  double mh, ml, y0, c; mp_mant_dd(a.l[NL - 1], a.l[NL - 2], NL >= 3 ? a.l[NL >= 3 ? NL - 3 : 0] : 0u, -odd, mh, ml); dd_rsqrt_seed(mh, ml, y0, c);
  mpn<NL> y, t, three; mp_from_double(y, y0); mp_from_double(t, c); mp_add(y, y, t); mp_set_i32(three, 3);
#pragma unroll 1
  for (int it = 0; it < mp_newton_steps<NL>(); it++) { mp_mul(t, y, y); mp_mul(t, t, m); mp_sub(t, three, t); mp_mul(y, y, t); y.exp -= 1; }
  y.exp -= (a.exp + odd) / 2; r = y;
}
// r = sqrt(a), rinv = 1/sqrt(a) for a > 0 (one correction step on the root)
template <int NL> HD void mp_sqrt_rsqrt(mpn<NL>& r, mpn<NL>& rinv, const mpn<NL>& a) {
  mpn<NL> y, s, t; mp_rsqrt(y, a); mp_mul(s, a, y);
  mp_mul(t, s, s); mp_sub(t, a, t); mp_mul(t, t, y); t.exp -= 1; mp_add(s, s, t);   // s += (a - s^2) y / 2
  r = s; rinv = y;
}

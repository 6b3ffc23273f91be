// mpw.cuh — warp-cooperative multi-limb arithmetic (north star subsystem 2: "multi-limb panel kernels with
// warp-level carry propagation").
//
// A number is spread over a warp: lane t holds mantissa limb (t mod NL), sign and exponent are warp-uniform.
// The product of two such numbers is formed column-wise — lane k (0..2NL-1) accumulates the partial-product
// column sum_{i+j=k} a_i b_j with operands fetched by shuffles — and the carries are resolved across lanes in
// O(1): neighbour shuffles for the multi-word column sums, then one ballot-based generate/propagate step (an
// integer addition plays carry-lookahead adder).  The result has the same bits as the single-thread mp_mul of
// mpf.cuh (top NL limbs of the exact product, truncated), which is how it is tested (clrs_debug_selftest).  It
// serves the sequential pivot chain of the diagonal-block Cholesky, where a single thread's multiply is the
// critical path.  NL = 8 (256 bit: 16 columns, the two half-warps mirror each other) and NL = 16 (512 bit: the 32
// columns fill the warp).
#pragma once
#include "mpf.cuh"

struct wnum { uint32_t limb; int32_t exp; int32_t sign; };     // limb (lane mod NL) of a warp-distributed number

template <int NL> struct WCfg {
  static_assert(NL == 8 || NL == 16, "warp-cooperative arithmetic: 8 or 16 limbs");
  static constexpr int W = 2 * NL;                               // product columns = shuffle segment width
  static constexpr uint32_t WMASK = (W == 32) ? 0xFFFFFFFFu : ((1u << (W & 31)) - 1u);
  static constexpr uint32_t LMASK = (1u << NL) - 1u;             // lanes holding the NL limbs
  static constexpr uint32_t AMASK = (1u << (NL + 1)) - 1u;       // lanes of the (NL+1)-word addition array
};

template <int NL> __device__ __forceinline__ wnum w_from(const mpn<NL>& a) {       // replicated -> distributed (a identical in all lanes)
  const int t = threadIdx.x & (NL - 1); wnum r; uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) if (t == i) v = a.l[i];
  r.limb = v; r.exp = a.exp; r.sign = a.sign; return r;
}
template <int NL> __device__ __forceinline__ mpn<NL> w_to(const wnum& a) {         // distributed -> replicated in every lane
  mpn<NL> r;
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = __shfl_sync(0xffffffffu, a.limb, i);
  r.exp = a.exp; r.sign = a.sign; return r;
}
template <int NL> __device__ __forceinline__ wnum w_load(const mpn<NL>* p) {       // from (shared) memory: each lane reads its limb
  wnum r; r.limb = p->l[threadIdx.x & (NL - 1)]; r.exp = p->exp; r.sign = p->sign; return r;
}
template <int NL> __device__ __forceinline__ void w_store(mpn<NL>* p, const wnum& a) {
  const int lane = threadIdx.x & 31;
  if (lane < NL) p->l[lane] = a.limb;
  if (lane == 0) { p->exp = a.exp; p->sign = a.sign; }
}

// r = a * b, all 32 lanes of the warp must call it
template <int NL> __device__ __forceinline__ wnum w_mul(const wnum& a, const wnum& b) {
  constexpr int W = WCfg<NL>::W;
  wnum r;
  if (a.sign == 0 || b.sign == 0) { r.limb = 0; r.exp = 0; r.sign = 0; return r; }   // warp-uniform
  const int k = threadIdx.x & (W - 1);
  uint64_t lo = 0; uint32_t hi = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const uint32_t ai = __shfl_sync(0xffffffffu, a.limb, i);
    const int j = k - i;
    const uint32_t bj = __shfl_sync(0xffffffffu, b.limb, j & (NL - 1));
    if (j >= 0 && j < NL) { const uint64_t p = (uint64_t)ai * bj; lo += p; hi += (lo < p) ? 1u : 0u; }
  }
  // column sum S_k = hi:lo = x + 2^32 y + 2^64 z ; limb k of the product = x_k + y_{k-1} + z_{k-2} + carries
  const uint32_t x = (uint32_t)lo, y = (uint32_t)(lo >> 32), z = hi;
  uint32_t y1 = __shfl_up_sync(0xffffffffu, y, 1, W), z2 = __shfl_up_sync(0xffffffffu, z, 2, W);
  if (k < 1) y1 = 0; if (k < 2) z2 = 0;
  const uint64_t t = (uint64_t)x + y1 + z2;
  const uint32_t u = (uint32_t)t; uint32_t c = (uint32_t)(t >> 32);                 // c in {0,1,2}
  uint32_t c1 = __shfl_up_sync(0xffffffffu, c, 1, W); if (k < 1) c1 = 0;
  const uint64_t v64 = (uint64_t)u + c1;
  const uint32_t v = (uint32_t)v64; const uint32_t g = (uint32_t)(v64 >> 32);        // g in {0,1}
  // carry-lookahead over the W lanes: cin_{k+1} = g_k | (p_k & cin_k), evaluated by one integer addition
  // (the carry out of the top column cannot occur: the product of two NL-limb numbers has 2 NL limbs)
  const uint32_t sh = (threadIdx.x & 31) & ~(W - 1);                                 // which segment of the warp
  const uint32_t G = (__ballot_sync(0xffffffffu, g != 0) >> sh) & WCfg<NL>::WMASK;
  const uint32_t P = (__ballot_sync(0xffffffffu, v == 0xFFFFFFFFu) >> sh) & WCfg<NL>::WMASK;
  const uint32_t X = G | P, Y = G;
  const uint32_t cin = (X + Y) ^ X ^ Y;
  const uint32_t limb = v + ((cin >> k) & 1u);                                       // limb k of the 2NL-limb product
  // top NL limbs, normalised: the product of two [1/2,1) mantissas is in [1/4,1)
  const uint32_t top = __shfl_sync(0xffffffffu, limb, W - 1, W);
  const uint32_t below = __shfl_up_sync(0xffffffffu, limb, 1, W);
  const bool norm = (top >> 31) != 0;
  const uint32_t shifted = norm ? limb : ((limb << 1) | (below >> 31));
  r.limb = __shfl_sync(0xffffffffu, shifted, NL + (threadIdx.x & (NL - 1)), W);
  r.exp = a.exp + b.exp - (norm ? 0 : 1); r.sign = a.sign * b.sign;
  return r;
}
// r = a + bsgn*b with one limb per lane: alignment by shuffles, carry / borrow resolved by one ballot-based
// lookahead step, renormalisation by ballot + clz.  Same truncation as mp_add_signed (one guard limb), so the
// bits equal the single-thread result.  Positions p = 0..NL of the (NL+1)-word working array live in lanes 0..NL
// (p = 0 is the guard limb); the other lanes carry zeros.
template <int NL> __device__ __forceinline__ wnum w_addsub(const wnum& a, const wnum& b, int bsgn) {
  constexpr uint32_t LMASK = WCfg<NL>::LMASK, AMASK = WCfg<NL>::AMASK;
  const int bs = b.sign * bsgn;
  if (bs == 0) return a;
  if (a.sign == 0) { wnum r = b; r.sign = bs; return r; }
  const int lane = threadIdx.x & 31;
  // order by magnitude (exponent, then mantissa)
  bool swap = a.exp < b.exp;
  if (a.exp == b.exp) {
    const uint32_t gt = __ballot_sync(0xffffffffu, a.limb > b.limb) & LMASK, lt = __ballot_sync(0xffffffffu, a.limb < b.limb) & LMASK;
    swap = lt > gt;
  }
  const uint32_t hl = swap ? b.limb : a.limb, ll = swap ? a.limb : b.limb;
  const int hexp = swap ? b.exp : a.exp, lexp = swap ? a.exp : b.exp, his = swap ? bs : a.sign, los = swap ? a.sign : bs;
  const uint32_t d = (uint32_t)(hexp - lexp);
  wnum r;
  if (d >= 32u * NL + 32u) { r.limb = hl; r.exp = hexp; r.sign = his; return r; }
  const int p = lane;                                             // position in the (NL+1)-word array
  uint32_t x = __shfl_sync(0xffffffffu, hl, (p - 1) & 31); if (p < 1 || p > NL) x = 0;
  const int s = (int)(d >> 5), bb = (int)(d & 31);
  const int q0 = p + s, q1 = p + s + 1;                           // words of the low operand feeding position p
  uint32_t y0 = __shfl_sync(0xffffffffu, ll, (q0 - 1) & 31); if (q0 < 1 || q0 > NL) y0 = 0;
  uint32_t y1 = __shfl_sync(0xffffffffu, ll, (q1 - 1) & 31); if (q1 < 1 || q1 > NL) y1 = 0;
  uint32_t y = __funnelshift_r(y0, y1, bb); if (p > NL) y = 0;
  int32_t e = hexp; uint32_t v;
  if (his == los) {
    const uint32_t t = x + y; const bool g = t < x, pr = t == 0xFFFFFFFFu;
    const uint32_t G = __ballot_sync(0xffffffffu, g) & AMASK, P = __ballot_sync(0xffffffffu, pr) & AMASK;
    const uint32_t X = G | P, cin = (X + G) ^ X ^ G;
    v = t + ((cin >> p) & 1u);
    if ((cin >> (NL + 1)) & 1u) {                                 // carry out of the top limb: shift right one bit
      uint32_t up = __shfl_down_sync(0xffffffffu, v, 1); if (p >= NL) up = 1u;      // the carry becomes the new top bit
      v = (v >> 1) | (up << 31); e += 1;
    }
  } else {
    const uint32_t t = x - y; const bool g = x < y, pr = x == y;
    const uint32_t G = __ballot_sync(0xffffffffu, g) & AMASK, P = __ballot_sync(0xffffffffu, pr) & AMASK;
    const uint32_t X = G | P, bin = (X + G) ^ X ^ G;
    v = t - ((bin >> p) & 1u); if (p > NL) v = 0;
    const uint32_t nz = __ballot_sync(0xffffffffu, v != 0) & AMASK;
    if (nz == 0) { r.limb = 0; r.exp = 0; r.sign = 0; return r; }
    const int tp = 31 - __clz((int)nz), lzw = NL - tp;
    const uint32_t topw = __shfl_sync(0xffffffffu, v, tp);
    const int lz = __clz((int)topw);
    const int src = p - lzw;                                      // word shift, then bit shift (funnel with the word below)
    uint32_t w1 = __shfl_sync(0xffffffffu, v, src & 31); if (src < 0 || src > NL) w1 = 0;
    uint32_t w0 = __shfl_sync(0xffffffffu, v, (src - 1) & 31); if (src - 1 < 0 || src - 1 > NL) w0 = 0;
    v = __funnelshift_l(w0, w1, lz);
    e -= 32 * lzw + lz;
  }
  r.limb = __shfl_sync(0xffffffffu, v, (lane & (NL - 1)) + 1);
  r.exp = e; r.sign = his;
  return r;
}
template <int NL> __device__ __forceinline__ wnum w_sub(const wnum& a, const wnum& b) { return w_addsub<NL>(a, b, -1); }
template <int NL> __device__ __forceinline__ wnum w_add(const wnum& a, const wnum& b) { return w_addsub<NL>(a, b, 1); }
// r = a^(-1/2), a > 0: Newton from a double seed, multiplications warp-cooperative
template <int NL> __device__ __forceinline__ wnum w_rsqrt(const wnum& a) {
  wnum m = a; const int odd = a.exp & 1; m.exp = -odd; m.sign = 1;
  const uint32_t A = __shfl_sync(0xffffffffu, a.limb, NL - 1), B = __shfl_sync(0xffffffffu, a.limb, NL - 2), C = __shfl_sync(0xffffffffu, a.limb, NL - 3);
  double mh, ml, y0d, cd; mp_mant_dd(A, B, C, -odd, mh, ml); dd_rsqrt_seed(mh, ml, y0d, cd);     // same seed as mp_rsqrt
  mpn<NL> y0, c0, three; mp_from_double(y0, y0d); mp_from_double(c0, cd); mp_set_i32(three, 3);
  wnum y = w_add<NL>(w_from<NL>(y0), w_from<NL>(c0)); const wnum w3 = w_from<NL>(three);
#pragma unroll 1
  for (int it = 0; it < mp_newton_steps<NL>(); it++) { wnum t = w_mul<NL>(y, y); t = w_mul<NL>(t, m); t = w_sub<NL>(w3, t); y = w_mul<NL>(y, t); y.exp -= 1; }
  y.exp -= (a.exp + odd) / 2;
  return y;
}

// a double as a warp-distributed number (same value as w_from(mp_from_double(d)) for normal d, without the limb loop)
template <int NL> __device__ __forceinline__ wnum w_from_double(double d) {
  wnum r; r.limb = 0; r.exp = 0; r.sign = 0;
  if (d == 0.0 || !(d == d)) return r;
  const uint64_t bits = (uint64_t)__double_as_longlong(d); const int ef = (int)((bits >> 52) & 0x7ff);
  const uint64_t mant = ((bits & 0xFFFFFFFFFFFFFull) | (1ull << 52)) << 11;      // (subnormals and infinities do not occur: seeds of mantissas in [1/4, 1))
  const int t = threadIdx.x & (NL - 1);
  r.limb = t == NL - 1 ? (uint32_t)(mant >> 32) : (t == NL - 2 ? (uint32_t)mant : 0u); r.exp = ef - 1022; r.sign = d < 0 ? -1 : 1; return r;
}
// ---- out-of-line variants ------------------------------------------------------------------------------------------------
// A sequential chain of warp-cooperative operations executed ONCE per step (the pivot chain of the diagonal-block Cholesky)
// is bound by instruction fetch when every operation is inlined: ~2500 straight-line instructions per column, no reuse,
// 25 cycles per instruction measured (clock64 timeline, profiles/r02_potrf_timeline.txt: 62 k cycles per pivot).  Called
// as functions the same operations run from ~2 KB of code that stays in the instruction cache.  Same bits as the inline forms.
template <int NL> __device__ __noinline__ wnum w_mul_c(wnum a, wnum b) { return w_mul<NL>(a, b); }
template <int NL> __device__ __noinline__ wnum w_addsub_c(wnum a, wnum b, int bsgn) { return w_addsub<NL>(a, b, bsgn); }
template <int NL> __device__ __noinline__ wnum w_rsqrt_c(wnum a) {
  wnum m = a; const int odd = a.exp & 1; m.exp = -odd; m.sign = 1;
  const uint32_t A = __shfl_sync(0xffffffffu, a.limb, NL - 1), B = __shfl_sync(0xffffffffu, a.limb, NL - 2), C = __shfl_sync(0xffffffffu, a.limb, NL - 3);
  double mh, ml, y0d, cd; mp_mant_dd(A, B, C, -odd, mh, ml); dd_rsqrt_seed(mh, ml, y0d, cd);     // same seed as mp_rsqrt
  wnum y = w_addsub_c<NL>(w_from_double<NL>(y0d), w_from_double<NL>(cd), 1); const wnum w3 = w_from_double<NL>(3.0);
#pragma unroll 1
  for (int it = 0; it < mp_newton_steps<NL>(); it++) { wnum t = w_mul_c<NL>(y, y); t = w_mul_c<NL>(t, m); t = w_addsub_c<NL>(w3, t, -1); y = w_mul_c<NL>(y, t); y.exp -= 1; }
  y.exp -= (a.exp + odd) / 2;
  return y;
}

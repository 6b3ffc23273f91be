// hostcheck.cpp — CPU build of the device arithmetic (mpf.cuh, i8split.cuh) for
// the `-m "not gpu"` tests: the SAME source the kernels compile is exercised
// here against mpmath / the oracle.  Test infrastructure only; the product
// library never links this file.
#include <vector>
#include "wire_host.h"
#include "i8split.cuh"
#include "lanes.cuh"
typedef mpn<8> N8;
static const size_t WS = 16 + 32;
extern "C" {
int hc_binop(int op, const void* a, const void* b, void* r) {
  N8 x, y, z; wire_to_mpn(x, a); wire_to_mpn(y, b);
  switch (op) { case 0: mp_add(z, x, y); break; case 1: mp_sub(z, x, y); break; case 2: mp_mul(z, x, y); break; case 3: mp_div(z, x, y); break; default: return 1; }
  mpn_to_wire(r, z); return 0;
}
int hc_unop(int op, const void* a, void* r) {
  N8 x, z, w; wire_to_mpn(x, a);
  switch (op) { case 0: mp_recip(z, x); break; case 1: mp_sqrt_rsqrt(z, w, x); break; case 2: mp_rsqrt(z, x); break; default: return 1; }
  mpn_to_wire(r, z); return 0;
}
// wire record with W 64-bit limbs -> 8-limb device number -> wire record with W2 limbs (top-aligned mantissas)
int hc_wire_convert(int W, const void* in, int W2, void* out) { N8 x; wire_to_mpn<8>(x, in, W); mpn_to_wire<8>(out, x, W2); return x.sign; }
double hc_to_double(const void* a) { N8 x; wire_to_mpn(x, a); return mp_to_double(x); }
void hc_from_double(double d, void* r) { N8 x; mp_from_double(x, d); mpn_to_wire(r, x); }
int hc_cmp(const void* a, const void* b) { N8 x, y; wire_to_mpn(x, a); wire_to_mpn(y, b); return mp_cmp(x, y); }
// the three local steps of the cross-rank sum (lanes.cuh); the reductions between them are the caller's (gloo in the CPU tests)
void hc_lane_exp(int n, const void* v, int32_t* E) { for (int i = 0; i < n; i++) { N8 x; wire_to_mpn(x, (const char*)v + i * WS); E[i] = mp_lane_exp(x); } }
void hc_to_lanes(int n, const void* v, const int32_t* E, long long* lanes) {       // lanes[i][9]
  for (int i = 0; i < n; i++) { N8 x; wire_to_mpn(x, (const char*)v + i * WS); long long t[9]; mp_to_lanes<8>(x, E[i], t); for (int k = 0; k < 9; k++) lanes[(size_t)i * 9 + k] = t[k]; } }
void hc_from_lanes(int n, const long long* lanes, const int32_t* E, void* out) {
  for (int i = 0; i < n; i++) { long long t[9]; for (int k = 0; k < 9; k++) t[k] = lanes[(size_t)i * 9 + k]; N8 r; mp_from_lanes<8>(r, t, E[i]); mpn_to_wire((char*)out + i * WS, r); } }
// C = A*B through split -> exact integer slice-pair sums -> recombine (what the int8 GEMM kernels compute)
int hc_gemm(int M, int N, int K, const void* A, const void* B, void* C) {
  constexpr int NS = I8Cfg<8>::NS;
  std::vector<N8> a((size_t)M * K), b((size_t)K * N);
  for (size_t i = 0; i < a.size(); i++) wire_to_mpn(a[i], (const char*)A + i * WS);
  for (size_t i = 0; i < b.size(); i++) wire_to_mpn(b[i], (const char*)B + i * WS);
  std::vector<int32_t> E(M, I8_EXP_NONE), F(N, I8_EXP_NONE);
  for (int i = 0; i < M; i++) for (int k = 0; k < K; k++) if (a[(size_t)i * K + k].sign && a[(size_t)i * K + k].exp > E[i]) E[i] = a[(size_t)i * K + k].exp;
  for (int j = 0; j < N; j++) for (int k = 0; k < K; k++) if (b[(size_t)k * N + j].sign && b[(size_t)k * N + j].exp > F[j]) F[j] = b[(size_t)k * N + j].exp;
  std::vector<int8_t> da((size_t)M * K * NS), db((size_t)K * N * NS);
  for (int i = 0; i < M; i++) for (int k = 0; k < K; k++) { int8_t dg[NS]; i8_split<8>(a[(size_t)i * K + k], E[i], dg); memcpy(&da[((size_t)i * K + k) * NS], dg, NS); }
  for (int j = 0; j < N; j++) for (int k = 0; k < K; k++) { int8_t dg[NS]; i8_split<8>(b[(size_t)k * N + j], F[j], dg); memcpy(&db[((size_t)j * K + k) * NS], dg, NS); }
  for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
    int32_t acc[NS]; for (int s = 0; s < NS; s++) acc[s] = 0; int64_t top = 0; int since = 0;
    for (int k = 0; k < K; k++) {
      const int8_t* x = &da[((size_t)i * K + k) * NS]; const int8_t* y = &db[((size_t)j * K + k) * NS];
      for (int s = 0; s < NS; s++) for (int t = 0; t <= s; t++) acc[s] += (int)x[t] * (int)y[s - t];
      if (++since == 2048) { i8_carry_normalize<NS>(acc, top); since = 0; }
    }
    i8_carry_normalize<NS>(acc, top);
    uint32_t dg[NS]; for (int s = 0; s < NS; s++) dg[s] = (uint32_t)acc[s];
    N8 r;
    if (E[i] == I8_EXP_NONE || F[j] == I8_EXP_NONE) mp_zero(r); else i8_recombine<8>(r, top, dg, E[i] + F[j]);
    mpn_to_wire((char*)C + ((size_t)i * N + j) * WS, r);
  }
  return 0;
}
}

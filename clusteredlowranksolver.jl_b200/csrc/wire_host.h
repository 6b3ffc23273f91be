// wire_host.h — host conversion between wire numbers (include/clrs_b200.h) and mpn<NL>.
#pragma once
#include <cstring>
#include "mpf.cuh"

// W = number of 64-bit limbs of the wire record (ceil(prec/64)); the mantissa is aligned at the TOP (the most significant
// wire limb has its top bit set), so a record with fewer limbs than the device number fills the upper limbs and one with
// more is truncated.  The two-argument forms are for W = NL/2 (prec = 32 NL).
template <int NL> inline void wire_to_mpn(mpn<NL>& r, const void* src, int W) {
  const char* s = (const char*)src; int64_t ex; int32_t sg; memcpy(&ex, s, 8); memcpy(&sg, s + 8, 4);
  for (int i = 0; i < NL; i++) r.l[i] = 0;
  if (sg == 0) { r.exp = 0; r.sign = 0; return; }
  const int nw = 2 * W;                                     // 32-bit limbs in the record, least significant first
  if (nw <= NL) memcpy(r.l + (NL - nw), s + 16, 4 * (size_t)nw); else memcpy(r.l, s + 16 + 4 * (size_t)(nw - NL), 4 * (size_t)NL);
  if (ex > (1 << 28)) ex = (1 << 28); if (ex < -(1 << 28)) ex = -(1 << 28);
  r.exp = (int32_t)ex; r.sign = sg < 0 ? -1 : 1;
}
template <int NL> inline void mpn_to_wire(void* dst, const mpn<NL>& a, int W) {
  char* s = (char*)dst; memset(s, 0, 16 + 8 * (size_t)W);
  if (a.sign == 0) return;
  const int nw = 2 * W;
  int64_t ex = a.exp; int32_t sg = a.sign; memcpy(s, &ex, 8); memcpy(s + 8, &sg, 4);
  if (nw <= NL) memcpy(s + 16, a.l + (NL - nw), 4 * (size_t)nw); else memcpy(s + 16 + 4 * (size_t)(nw - NL), a.l, 4 * (size_t)NL);
}
template <int NL> inline void wire_to_mpn(mpn<NL>& r, const void* src) { wire_to_mpn<NL>(r, src, NL / 2); }
template <int NL> inline void mpn_to_wire(void* dst, const mpn<NL>& a) { mpn_to_wire<NL>(dst, a, NL / 2); }

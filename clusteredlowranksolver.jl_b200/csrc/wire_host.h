// wire_host.h — host conversion between wire numbers (include/clrs_b200.h) and mpn<NL>.
#pragma once
#include <cstring>
#include "mpf.cuh"

template <int NL> inline void wire_to_mpn(mpn<NL>& r, const void* src) {
  const char* s = (const char*)src; int64_t ex; int32_t sg; memcpy(&ex, s, 8); memcpy(&sg, s + 8, 4);
  if (sg == 0) { for (int i = 0; i < NL; i++) r.l[i] = 0; r.exp = 0; r.sign = 0; return; }
  memcpy(r.l, s + 16, 4 * NL);                             // little-endian: uint64 limb k = l[2k] | l[2k+1] << 32
  if (ex > (1 << 28)) ex = (1 << 28); if (ex < -(1 << 28)) ex = -(1 << 28);
  r.exp = (int32_t)ex; r.sign = sg < 0 ? -1 : 1;
}
template <int NL> inline void mpn_to_wire(void* dst, const mpn<NL>& a) {
  char* s = (char*)dst; memset(s, 0, 16 + 4 * NL);
  if (a.sign == 0) return;
  int64_t ex = a.exp; int32_t sg = a.sign; memcpy(s, &ex, 8); memcpy(s + 8, &sg, 4); memcpy(s + 16, a.l, 4 * NL);
}

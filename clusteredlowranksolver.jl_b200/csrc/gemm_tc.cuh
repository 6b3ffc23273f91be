// gemm_tc.cuh — the int8 slice-pair GEMM on the 5th-generation tensor cores
// (north star subsystem 1): TMA -> shared memory -> tcgen05.mma kind::i8 -> int32
// accumulators in TMEM -> carry-propagating epilogue.
//
// Operands are "tc panels": for every slice t an int8 matrix [nvec][Kp] (K-major,
// Kp = K rounded up to 32), i.e. one 3-D tensor {Kp, nvec, NS} described to TMA.
// An output tile is 128 (rows of the left panel) x BN (rows of the right panel).
// The slice-pair sums D_s = sum_{t+u=s} A_t B_u^T are produced `group` diagonals at
// a time, least significant first (group = 4 with BN <= 128, or 3 with BN <= 160: both
// fill the 512 TMEM columns; the wider tile makes the MMA tensor-bound instead of
// shared-memory-bound and cuts N = 300 into 160 + 144 instead of 128 + 128 + 48):
//
//   group g = diagonals d0..d1 (d0 = group*g): `group` int32 accumulators of BN columns in
//   TMEM.  For every 128-byte K chunk the slices stream through two shared-memory
//   rings:  A_0..A_d1  and  B_d1..B_0 ; A_i meets the window B_{d0-i..d1-i}, so each
//   operand chunk is loaded once per group and used by up to four MMAs.
//   Epilogue (4 warps, one TMEM lane = one output row per thread): low byte of
//   each diagonal (plus carry) goes to a byte plane in HBM, the carry moves up;
//   the carry out of the group is written back into TMEM as the initial value of
//   the next group's least significant accumulator, so carries never leave the SM.
//
// The result is NS byte planes + one int32 plane per output: the exact integer
// sum_s D_s 256^(NS-1-s) in two's complement radix 256.  k_tc_recombine turns it
// into multi-limb numbers (same i8_recombine as the CUDA-core path: the two paths
// produce identical bits).
//
// Products with few output tiles (the n x n block products of the iteration: 6-9 tiles on
// 148 SMs) run in "dsplit" mode: blockIdx.z enumerates (K chunk, diagonal group), every CTA
// produces the raw int32 sums of ONE group over ONE K range and adds them into int32
// planes with red.global.add (exact, order-independent); k_tc_recombine_raw resolves
// the carries.  27 CTAs x 150 us become ~220 CTAs x 20 us.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "mpf.cuh"
#include "i8split.cuh"

namespace tc {

constexpr int BM = 128;            // rows per tile = TMEM lanes
constexpr int KC = 128;            // bytes of K per ring slot (one 128B swizzle atom row)
constexpr int BNMAX = 160;         // widest column tile (3 accumulators x 160 columns = 480 of the 512 TMEM columns)
constexpr int SLOT_BYTES = BM * KC;      // 16 KiB: one left-operand chunk
constexpr int SLOTB_BYTES = BNMAX * KC;  // 20 KiB: one right-operand chunk
constexpr int NA = 4, NB = 8;      // ring slots for left / right operand chunks (powers of two)
constexpr int NMMA = 4;            // MMA-issuing warps: warp 1+k owns window offset k (one accumulator per step)
constexpr int NTHREADS = 288;      // warp 0: TMA, warps 1-4: MMA, warps 5-8: epilogue
constexpr int SMEM_BYTES = NA * SLOT_BYTES + NB * SLOTB_BYTES + 1024 /*align*/ + 512 /*barriers*/;   // 230912 <= 232448

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// empty asm that "rewrites" 16 registers: orders their later uses after the preceding asm volatile (the tcgen05.wait::ld)
__device__ __forceinline__ void reg_fence16(uint32_t (&v)[16]) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                    "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, rows 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct Args {
  int M, N, Kp;            // Kp: padded K (multiple of 32) of this launch's K range
  int k0;                  // first K byte of the range (multiple of 128)
  int BN;                  // tile width, multiple of 16, <= 128 (group 4) or <= 160 (group 3)
  int group;               // diagonals per group: 4 or 3
  int dsplit;              // != 0: blockIdx.z = kz * ngroups + group index; raw int32 sums are added into oraw
  int32_t* oraw;           // dsplit: [NS][M][Npitch] int32, zeroed by the caller
  int a_bvec, b_bvec;      // panel rows per batch step (0: shared)
  int NS;                  // slices
  int Npitch, batch;
  uint8_t* obytes;         // [NS][batch][M][Npitch]
  int32_t* otop;           // [batch][M][Npitch]
  int lower_only;
  int kz_stride;           // > 0: blockIdx.z selects the K range [z*kz_stride, ...) instead of a batch entry (split-K, partial results per z)
  int Kp_total;
  long long* dbg;          // optional timeline of CTA (0,0,0): clock64 stamps (kernel tuning only)
  int epi;                 // epilogue variant: 0 = one TMEM load + wait per accumulator and column chunk; 1 = the loads of all accumulators of a chunk in flight together, one wait (CLRS_TC_EPI)
};

__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* ringA = ring;
  unsigned char* ringB = ring + NA * SLOT_BYTES;
  uint64_t* full_a = (uint64_t*)(ring + NA * SLOT_BYTES + NB * SLOTB_BYTES);
  uint64_t* full_b = full_a + NA;
  uint64_t* empty_a = full_b + NB;
  uint64_t* empty_b = empty_a + NA;
  uint64_t* tmem_full = empty_b + NB;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * a.BN, m0 = blockIdx.y * BM;
  if (a.lower_only && n0 > m0 + BM - 1) return;
  const int NS = a.NS, BN = a.BN, GROUP = a.group;
  const int bn = min(BN, ((a.N - n0) + 15) & ~15);            // the last column tile may be narrower: MMAs and epilogue cover bn columns (TMA still fills BN rows, zeros beyond N)
  const int ngroups = (NS + GROUP - 1) / GROUP;
  // dsplit: this CTA owns ONE diagonal group (the heaviest groups, i.e. the largest g, get the lowest blockIdx.z so they start first)
  const int bz = a.dsplit ? (int)blockIdx.z / ngroups : (int)blockIdx.z;
  const int g_hi = a.dsplit ? ngroups - 1 - (int)blockIdx.z % ngroups : ngroups - 1, g_lo = a.dsplit ? g_hi : 0;
  const int kbase = a.kz_stride > 0 ? a.k0 + bz * a.kz_stride : a.k0;
  const int Kp = a.kz_stride > 0 ? min(a.kz_stride, a.Kp_total - bz * a.kz_stride) : a.Kp;
  const int nkc = (Kp + KC - 1) / KC;
  const int brow_z = a.kz_stride > 0 ? 0 : bz;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; s++) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], NMMA); }
    for (int s = 0; s < NB; s++) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], NMMA); }
    mbar_init(tmem_full, NMMA); mbar_init(tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
      uint32_t ca = 0, cb = 0;                                  // running load counters of the two rings
      const uint32_t abytes = BM * KC, bbytes = (uint32_t)BN * KC;
      const int arow = brow_z * a.a_bvec + m0, brow = brow_z * a.b_bvec + n0;
      for (int g = g_hi; g >= g_lo; g--) {
        const int d0 = g * GROUP, d1 = min(d0 + GROUP - 1, NS - 1), w = d1 - d0;
        for (int kc = 0; kc < nkc; kc++) {
          const int kcoord = kbase + kc * KC;
          auto loadB = [&](int slice) {
            const uint32_t slot = cb & (NB - 1);
            mbar_wait(&empty_b[slot], ((cb >> 3) & 1) ^ 1);
            mbar_expect_tx(&full_b[slot], bbytes);
            tma_load_3d(ringB + slot * SLOTB_BYTES, &tmB, &full_b[slot], kcoord, brow, slice);
            cb++;
          };
          auto loadA = [&](int slice) {
            const uint32_t slot = ca & (NA - 1);
            mbar_wait(&empty_a[slot], ((ca >> 2) & 1) ^ 1);
            mbar_expect_tx(&full_a[slot], abytes);
            tma_load_3d(ringA + slot * SLOT_BYTES, &tmA, &full_a[slot], kcoord, arow, slice);
            ca++;
          };
          for (int sp = 0; sp < w; sp++) loadB(d1 - sp);           // B_{s'} holds slice d1 - s'
          for (int i = 0; i <= d1; i++) { if (i + w <= d1) loadB(d1 - (i + w)); loadA(i); }
        }
      }
    }
  } else if (warp <= NMMA) {
    // ===================== MMA issuers =====================
    // A single thread cannot issue tcgen05.mma fast enough for N <= 128 (profiles/: ~110 cycles per
    // instruction from one warp vs 56-64 cycles of tensor work), so the four accumulators of a step are
    // driven by four warps: warp 1+k issues the pair (A_i, B_{s'=i+k}).  Every warp walks the same
    // warp-uniform schedule, one elected lane issues; every warp commits every step so the barrier
    // counts are constant.
    static_assert(NA == 4 && NB == 8, "ring phase shifts below assume NA = 4, NB = 8");
    const int koff = warp - 1;
    const uint64_t descA0 = make_desc(smem_u32(ringA)), descB0 = make_desc(smem_u32(ringB));   // slot s: + s * (SLOT_BYTES >> 4)
    const bool leader = (lane == 0);
    const bool dbgon = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 1;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    uint32_t ca = 0, cbase = 0; uint32_t epi_parity = 0; bool first_group = true;
    for (int g = g_hi; g >= g_lo; g--) {
      const int d0 = g * GROUP, d1 = min(d0 + GROUP - 1, NS - 1), w = d1 - d0;
      const bool carry_in = (g != g_hi);
      if (dbgon && leader) a.dbg[(ngroups - 1 - g) * 8 + 0] = clock64();
      // the accumulators are free once the epilogue has drained the previous group
      if (!first_group) { mbar_wait(tmem_empty, epi_parity); epi_parity ^= 1; asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      first_group = false;
      if (dbgon && leader) a.dbg[(ngroups - 1 - g) * 8 + 1] = clock64();
      const uint32_t tcol = tmem_base + (uint32_t)((w - koff) * BN);         // this warp's accumulator (diagonal d1 - koff)
      for (int kc = 0; kc < nkc; kc++) {
        const int ninstr = min(KC / 32, (Kp - kc * KC) / 32);
        uint32_t cbw = cbase;                                           // next B load (of this K chunk) not yet waited for
        for (int i = 0; i <= d1; i++) {
          const int sp = i + koff;
          const bool active = koff <= w && sp <= d1;
          if (active) { while (cbw <= cbase + (uint32_t)sp) { mbar_wait(&full_b[cbw & (NB - 1)], (cbw >> 3) & 1); cbw++; } }
          const uint32_t slotA = ca & (NA - 1);
          // every warp waits for A_i even when it has no pair in this step: it keeps the warps within one
          // ring phase of each other, so the per-step commits below can never alias a later phase
          mbar_wait(&full_a[slotA], (ca >> 2) & 1);
          if (active) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dA = descA0 + (uint64_t)(slotA * (SLOT_BYTES >> 4));
            const uint64_t dB = descB0 + (uint64_t)(((cbase + (uint32_t)sp) & (NB - 1)) * (SLOTB_BYTES >> 4));
            const uint32_t acc0 = (kc == 0 && i == 0) ? ((carry_in && koff == 0) ? 1u : 0u) : 1u;   // koff == 0 is the least significant diagonal
            if (leader) {
              if (ninstr == 4) {
                umma_i8(tcol, dA, dB, idesc, acc0); umma_i8(tcol, dA + 2, dB + 2, idesc, 1u);
                umma_i8(tcol, dA + 4, dB + 4, idesc, 1u); umma_i8(tcol, dA + 6, dB + 6, idesc, 1u);
              } else {
                for (int kk = 0; kk < ninstr; kk++) umma_i8(tcol, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idesc, kk == 0 ? acc0 : 1u);
              }
            }
          }
          if (leader) { umma_commit(&empty_a[slotA]); umma_commit(&empty_b[(cbase + (uint32_t)i) & (NB - 1)]); }   // A_i and B_{s'=i} are done
          __syncwarp();
          ca++;
        }
        cbase += (uint32_t)(d1 + 1);
      }
      if (leader) umma_commit(tmem_full);
      if (dbgon && leader) a.dbg[(ngroups - 1 - g) * 8 + 2] = clock64();
      __syncwarp();
    }
  } else {
    // ===================== epilogue: 4 warps, one output row per thread =====================
    const int quad = warp & 3;                                   // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int m = m0 + row;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const size_t plane = (size_t)a.batch * a.M * a.Npitch;
    const size_t rowoff = ((size_t)bz * a.M + m) * a.Npitch;
    uint32_t full_parity = 0;
    for (int g = g_hi; g >= g_lo; g--) {
      const int d0 = g * GROUP, d1 = min(d0 + GROUP - 1, NS - 1), w = d1 - d0;
      const bool dbge = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 160;
      if (dbge) a.dbg[(ngroups - 1 - g) * 8 + 3] = clock64();
      mbar_wait(tmem_full, full_parity); full_parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbge) a.dbg[(ngroups - 1 - g) * 8 + 4] = clock64();
      if (a.dsplit) {
        // raw sums of this group and K range: added into the int32 planes (exact, so the order of the CTAs does not matter)
        for (int c0 = 0; c0 < bn; c0 += 16) {
          const bool store_ok = (m < a.M) && (n0 + c0 < a.Npitch);
          for (int acc = w; acc >= 0; acc--) {
            uint32_t v[16];
            tmem_ld16(tlane + (uint32_t)(acc * BN + c0), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (store_ok) {
              int32_t* dst = a.oraw + ((size_t)(d0 + acc) * a.M + m) * a.Npitch + n0 + c0;
#pragma unroll
              for (int q = 0; q < 16; q++) if ((int32_t)v[q] != 0) atomicAdd(dst + q, (int32_t)v[q]);
            }
          }
        }
        continue;
      }
      if (a.epi == 1) {
        // all accumulators of a 16-column chunk are loaded together and one tcgen05.wait::ld covers them (8-10 exposed TMEM round trips
        // per group instead of 30-32).  A double-buffered version of this loop (next chunk's loads in flight during the carry chain)
        // needed 168 registers with spills in this one-body kernel and slowed the MMA-issuing warps: 0.98 -> 1.16 ms per launch.
        for (int c0 = 0; c0 < bn; c0 += 16) {
          uint32_t v[4][16];
#pragma unroll
          for (int acc = 0; acc < 4; acc++) if (acc <= w) tmem_ld16(tlane + (uint32_t)(acc * BN + c0), v[acc]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int acc = 0; acc < 4; acc++) reg_fence16(v[acc]);
          int32_t carry[16];
#pragma unroll
          for (int q = 0; q < 16; q++) carry[q] = 0;
          const bool store_ok = (m < a.M) && (n0 + c0 < a.Npitch);
#pragma unroll
          for (int acc = 3; acc >= 0; acc--) {
            if (acc > w) continue;
            uint32_t packed[4] = {0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const int32_t t = (int32_t)v[acc][q] + carry[q];
              packed[q >> 2] |= (uint32_t)(t & 255) << (8 * (q & 3));
              carry[q] = t >> 8;
            }
            if (store_ok) *(uint4*)(a.obytes + (size_t)(d0 + acc) * plane + rowoff + n0 + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
          if (g > 0) {
            uint32_t cv[16];
#pragma unroll
            for (int q = 0; q < 16; q++) cv[q] = (uint32_t)carry[q];
            tmem_st16(tlane + (uint32_t)((GROUP - 1) * BN + c0), cv);
          } else if (store_ok) {
            int4* dst = (int4*)(a.otop + rowoff + n0 + c0);
            dst[0] = make_int4(carry[0], carry[1], carry[2], carry[3]); dst[1] = make_int4(carry[4], carry[5], carry[6], carry[7]);
            dst[2] = make_int4(carry[8], carry[9], carry[10], carry[11]); dst[3] = make_int4(carry[12], carry[13], carry[14], carry[15]);
          }
        }
      } else
      for (int c0 = 0; c0 < bn; c0 += 16) {
        int32_t carry[16];
#pragma unroll
        for (int q = 0; q < 16; q++) carry[q] = 0;
        const bool store_ok = (m < a.M) && (n0 + c0 < a.Npitch);
        for (int acc = w; acc >= 0; acc--) {
          uint32_t v[16];
          tmem_ld16(tlane + (uint32_t)(acc * BN + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          uint32_t packed[4] = {0, 0, 0, 0};
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int32_t t = (int32_t)v[q] + carry[q];
            packed[q >> 2] |= (uint32_t)(t & 255) << (8 * (q & 3));
            carry[q] = t >> 8;
          }
          if (store_ok) *(uint4*)(a.obytes + (size_t)(d0 + acc) * plane + rowoff + n0 + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        if (g > 0) {
          uint32_t cv[16];
#pragma unroll
          for (int q = 0; q < 16; q++) cv[q] = (uint32_t)carry[q];
          tmem_st16(tlane + (uint32_t)((GROUP - 1) * BN + c0), cv);   // initial value of the next group's least significant diagonal
        } else if (store_ok) {
          int4* dst = (int4*)(a.otop + rowoff + n0 + c0);
          dst[0] = make_int4(carry[0], carry[1], carry[2], carry[3]); dst[1] = make_int4(carry[4], carry[5], carry[6], carry[7]);
          dst[2] = make_int4(carry[8], carry[9], carry[10], carry[11]); dst[3] = make_int4(carry[12], carry[13], carry[14], carry[15]);
        }
      }
      if (dbge) a.dbg[(ngroups - 1 - g) * 8 + 5] = clock64();
      // the group is drained (and, for g > 0, the carry is in place): the MMA warps may write the accumulators again
      if (g > 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tmem_empty);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}


}  // namespace tc

// split into the tc panel layout: planes[t][vec][Kp] int8; one thread per (vec, 4 consecutive k)
template <int NL> __global__ void k_split_tc(VecView v, const int32_t* E, int Kp, int64_t nvec_pitch, uint8_t* planes, int kfast) {
  constexpr int NS = I8Cfg<NL>::NS;
  const int K4 = Kp / 4;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)v.nvec * K4) return;
  int vec, k4; if (kfast) { k4 = (int)(idx % K4); vec = (int)(idx / K4); } else { vec = (int)(idx % v.nvec); k4 = (int)(idx / v.nvec); }
  const mpn<NL>* p = (const mpn<NL>*)v.base + vec_off(v, vec);
  const int32_t e = E[vec];
  uint32_t w[NS];
#pragma unroll
  for (int t = 0; t < NS; t++) w[t] = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int k = 4 * k4 + q;
    if (k < v.K) {
      mpn<NL> a = p[(int64_t)k * v.sk]; int8_t dg[NS]; i8_split<NL>(a, e, dg);
#pragma unroll
      for (int t = 0; t < NS; t++) w[t] |= (uint32_t)(uint8_t)dg[t] << (8 * q);
    }
  }
#pragma unroll
  for (int t = 0; t < NS; t++) *(uint32_t*)(planes + ((int64_t)t * nvec_pitch + vec) * Kp + 4 * k4) = w[t];
}

// Split over the PACKED UPPER TRIANGLE of n x n matrices (one matrix per vector): k <-> (a <= b), row by row.  With sym != 0 the
// entry is M[a][b] + M[b][a] for a != b — so that <A, T> = sum_{a<=b} A[a][b] (T[a][b] + T[b][a] (a != b)) for a symmetric A runs
// over K = n (n + 1) / 2 instead of n^2 (the dense Schur inner products, src/solver.jl:1100-1103).  E must already allow for the
// extra bit of the sums (k_exp_add).
// E[i] += add for the vectors that are not all zero (the symmetrised sums can carry one bit beyond the largest entry; the
// recombination reads the same exponents the split used)
__global__ void k_exp_add(int n, int32_t* E, int add) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n && E[i] != I8_EXP_NONE) E[i] += add; }
template <int NL> __global__ void k_split_tc_tri(const mpn<NL>* base, int64_t mstride, int nvec, int n, int sym, const int32_t* E, int Kp, int64_t nvec_pitch, uint8_t* planes) {
  constexpr int NS = I8Cfg<NL>::NS;
  const int K4 = Kp / 4;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nvec * K4) return;
  const int k4 = (int)(idx % K4), vec = (int)(idx / K4);
  const mpn<NL>* M = base + (int64_t)vec * mstride;
  const int32_t e = E[vec];
  const int KS = n * (n + 1) / 2, k = 4 * k4;
  // row a of packed index k: the largest a with a (2n - a + 1) / 2 <= k
  int a = (int)(((2.0 * n + 1.0) - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * (double)k)) * 0.5);
  if (a < 0) a = 0; if (a > n - 1) a = n - 1;
  while (a + 1 < n && (int64_t)(a + 1) * (2 * n - a) / 2 <= k) a++;
  while (a > 0 && (int64_t)a * (2 * n - a + 1) / 2 > k) a--;
  int b = a + (k - (int)((int64_t)a * (2 * n - a + 1) / 2));
  uint32_t w[NS];
#pragma unroll
  for (int t = 0; t < NS; t++) w[t] = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if (k + q < KS) {
      mpn<NL> v = M[(int64_t)a * n + b];
      if (sym && a != b) { const mpn<NL> o = M[(int64_t)b * n + a]; mp_add(v, v, o); }
      int8_t dg[NS]; i8_split<NL>(v, e, dg);
#pragma unroll
      for (int t = 0; t < NS; t++) w[t] |= (uint32_t)(uint8_t)dg[t] << (8 * q);
      if (++b == n) { a++; b = a; }
    }
  }
#pragma unroll
  for (int t = 0; t < NS; t++) *(uint32_t*)(planes + ((int64_t)t * nvec_pitch + vec) * Kp + 4 * k4) = w[t];
}
// flag[0] |= 1 if some matrix of the batch is not symmetric (bit pattern comparison)
template <int NL> __global__ void k_check_symmetric(const mpn<NL>* base, int64_t count, int n, int* flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count * n * n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t mtx = i / ((int64_t)n * n); const int r = (int)((i / n) % n), c = (int)(i % n);
    if (r < c) { const mpn<NL> x = base[mtx * n * n + (int64_t)r * n + c], y = base[mtx * n * n + (int64_t)c * n + r];
      bool same = x.sign == y.sign && (x.sign == 0 || x.exp == y.exp);
      if (same && x.sign != 0) for (int q = 0; q < NL; q++) same = same && x.l[q] == y.l[q];
      if (!same) atomicOr(flag, 1); }
  }
}

// Same split for vectors whose entries are NOT contiguous (columns of a row-major matrix): consecutive
// threads take consecutive vectors so the 40-byte reads coalesce, and the digits go through a
// shared-memory transpose so each plane row receives full 32-byte segments.
template <int NL> struct SplitTCfg { static constexpr int VT = (I8Cfg<NL>::NS <= 35) ? 32 : 16; };   // vectors per CTA (static smem <= 48 KiB)
template <int NL> __global__ void __launch_bounds__(256) k_split_tc_t(VecView v, const int32_t* E, int Kp, int64_t nvec_pitch, uint8_t* planes) {
  constexpr int NS = I8Cfg<NL>::NS, VT = SplitTCfg<NL>::VT;
  __shared__ uint32_t sm[NS][VT][9];
  const int vl = threadIdx.x % VT, kl = threadIdx.x / VT;            // blockDim.x = VT * 8
  const int vec = blockIdx.x * VT + vl, k4 = blockIdx.y * 8 + kl;
  uint32_t w[NS];
#pragma unroll
  for (int t = 0; t < NS; t++) w[t] = 0;
  if (vec < v.nvec) {
    const mpn<NL>* p = (const mpn<NL>*)v.base + vec_off(v, vec);
    const int32_t e = E[vec];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = 4 * k4 + q;
      if (k < v.K) {
        mpn<NL> a = p[(int64_t)k * v.sk]; int8_t dg[NS]; i8_split<NL>(a, e, dg);
#pragma unroll
        for (int t = 0; t < NS; t++) w[t] |= (uint32_t)(uint8_t)dg[t] << (8 * q);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NS; t++) sm[t][vl][kl] = w[t];
  __syncthreads();
  for (int idx = threadIdx.x; idx < NS * VT; idx += VT * 8) {
    const int t = idx / VT, vv = idx % VT, vec2 = blockIdx.x * VT + vv;
    if (vec2 < v.nvec) {
      uint4* dst = (uint4*)(planes + ((int64_t)t * nvec_pitch + vec2) * Kp + 32 * blockIdx.y);
      dst[0] = make_uint4(sm[t][vv][0], sm[t][vv][1], sm[t][vv][2], sm[t][vv][3]);
      dst[1] = make_uint4(sm[t][vv][4], sm[t][vv][5], sm[t][vv][6], sm[t][vv][7]);
    }
  }
}

// byte planes + top -> multi-limb C (op with D), one thread per output
template <int NL, bool PARTIAL = false> __global__ void k_tc_recombine(int M, int N, int Npitch, int batch, const uint8_t* obytes, const int32_t* otop,
                                                 const int32_t* EA, int64_t a_bvec, const int32_t* EB, int64_t b_bvec,
                                                 mpn<NL>* C, int ldc, int64_t c_bs, const mpn<NL>* D, int ldd, int64_t d_bs, int mode, int lower_only, int nsum, int trans, int ns_used = I8Cfg<NL>::NS) {
  constexpr int NS = I8Cfg<NL>::NS;
  // trans: the product was formed with the operands swapped (small left operand on the N side); element (m,n) goes to C[n][m]
  // nsum > 1: the `batch` planes are split-K partial results of ONE product and are summed here
  // (measured alternatives that were slower or no faster than this one-output-per-thread form: four outputs per thread
  //  with 32-bit plane loads - register count 47 -> 168; results staged in shared memory for 128-byte stores - 0.7 -> 1.1 ms
  //  on 27e6 outputs: the kernel is bound by load/store instruction issue, not by sectors; profiles/r01_hbm_kernels_ncu.txt)
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nb = nsum > 1 ? 1 : batch;
  if (idx >= (int64_t)nb * M * N) return;
  const int n = (int)(idx % N), m = (int)((idx / N) % M), bz = (int)(idx / ((int64_t)N * M));
  if (lower_only && n > m) return;
  const size_t plane = (size_t)batch * M * Npitch;
  const int32_t ea = EA[(int64_t)bz * a_bvec + m], eb = EB[(int64_t)bz * b_bvec + n];
  mpn<NL> r; mp_zero(r);
  for (int z = 0; z < (nsum > 1 ? nsum : 1); z++) {
    const size_t off = ((size_t)(nsum > 1 ? z : bz) * M + m) * Npitch + n;
    uint32_t dg[NS];
    // PARTIAL (matmul_prec): diagonals beyond ns_used were not produced.  A separate instantiation: predicating the 35 byte loads of the
    // full-precision kernel doubled its time (114 -> 228 us per 5.7e6 outputs)
#pragma unroll
    for (int s = 0; s < NS; s++) { if constexpr (PARTIAL) dg[s] = s < ns_used ? obytes[(size_t)s * plane + off] : 0u; else dg[s] = obytes[(size_t)s * plane + off]; }
    const int64_t top = (int64_t)otop[off] * 256 + (int64_t)dg[0];
    mpn<NL> t;
    if (ea == I8_EXP_NONE || eb == I8_EXP_NONE) mp_zero(t); else i8_recombine<NL>(t, top, dg, ea + eb);
    if (z == 0) r = t; else mp_add(r, r, t);
  }
  const int cr = trans ? n : m, cc = trans ? m : n;
  if (mode == 1 || mode == 2) { mpn<NL> d = D[(int64_t)bz * d_bs + (int64_t)cr * ldd + cc]; if (mode == 1) mp_sub(r, d, r); else mp_add(r, d, r); }
  else if (mode == 3) r.sign = -r.sign;
  C[(int64_t)bz * c_bs + (int64_t)cr * ldc + cc] = r;
}

// dsplit products: raw int32 slice-pair sums [NS][M][Npitch] (all K ranges and diagonal groups already added by the
// product kernel) -> carry-normalised digits -> multi-limb C (op with D); one thread per output, plane reads coalesce along n
template <int NL> __global__ void k_tc_recombine_raw(int M, int N, int Npitch, const int32_t* oraw, const int32_t* EA, const int32_t* EB,
                                                     mpn<NL>* C, int ldc, const mpn<NL>* D, int ldd, int mode, int lower_only, int trans) {
  constexpr int NS = I8Cfg<NL>::NS;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  const int n = (int)(idx % N), m = (int)(idx / N);
  if (lower_only && n > m) return;
  const size_t plane = (size_t)M * Npitch, off = (size_t)m * Npitch + n;
  const int32_t ea = EA[m], eb = EB[n];
  mpn<NL> r;
  if (ea == I8_EXP_NONE || eb == I8_EXP_NONE) mp_zero(r);
  else {
    uint32_t dg[NS]; int64_t carry = 0;
#pragma unroll
    for (int s = NS - 1; s >= 1; s--) { const int64_t t = (int64_t)oraw[(size_t)s * plane + off] + carry; dg[s] = (uint32_t)(t & 255); carry = t >> 8; }
    dg[0] = 0;
    const int64_t top = (int64_t)oraw[off] + carry;
    i8_recombine<NL>(r, top, dg, ea + eb);
  }
  const int cr = trans ? n : m, cc = trans ? m : n;
  if (mode == 1 || mode == 2) { mpn<NL> d = D[(int64_t)cr * ldd + cc]; if (mode == 1) mp_sub(r, d, r); else mp_add(r, d, r); }
  else if (mode == 3) r.sign = -r.sign;
  C[(int64_t)cr * ldc + cc] = r;
}

// adler_core.cuh -- Adler-32 building blocks shared by adler32.cu and the fused per-block update in
// inflate.cu (reference src/zipc_deflate.ml:175-198).
#pragma once
#include <stdint.h>
#include "../../include/zipc_b200.h"

namespace zb {

// One step of the reference's chunk recurrence over per-chunk partial sums (A = sum b, B = sum (n-i) b):
//   s1' = rem(s1 + A), s2' = rem(s2 + n*s1 + B), int32 wrap then *signed* rem in REF_COMPAT mode.
__host__ __device__ inline void adler_fold_step(uint32_t &s1, uint32_t &s2, uint32_t n, uint32_t A, uint32_t B, int mode) {
  uint32_t t2 = s2 + n * s1 + B;  // all int32-wrapping in the reference
  uint32_t t1 = s1 + A;
  if (mode == ZIPC_ADLER_REF_COMPAT) {
    s1 = (uint32_t)((int32_t)t1 % 65521);
    s2 = (uint32_t)((int32_t)t2 % 65521);
  } else {  // RFC 1950: exact arithmetic (n*s1 + B + s2 can exceed 2^32)
    uint64_t w2 = (uint64_t)s2 + (uint64_t)n * s1 + B;
    s1 = t1 % 65521u;
    s2 = (uint32_t)(w2 % 65521u);
  }
}

#if defined(__CUDACC__)
template <bool COHERENT>
__device__ __forceinline__ uint4 adler_ld16(const uint4 *p) {
  uint4 v;
  if (COHERENT)  // data written earlier in the same kernel: go to L2
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <bool COHERENT>
__device__ __forceinline__ uint32_t adler_ld8(const uint8_t *p) {
  if (!COHERENT) return *p;
  unsigned int v;
  asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// (A, B) of one range [ptr, ptr+n), n <= 5552, valid in every lane after the butterfly
template <bool COHERENT>
__device__ __forceinline__ void adler_range_warp(const uint8_t *ptr, uint32_t n, int lane, uint32_t &A, uint32_t &B) {
  uint32_t a = 0, b = 0;
  uint32_t head = (uint32_t)((16 - ((uintptr_t)ptr & 15)) & 15);
  if (head > n) head = n;
  uint32_t nblk = (n - head) >> 4;
  uint32_t tail_off = head + (nblk << 4);
  if (lane == 0) {  // ragged ends, bytewise
    for (uint32_t i = 0; i < head; i++) { uint32_t v = adler_ld8<COHERENT>(ptr + i); a += v; b += (n - i) * v; }
    for (uint32_t i = tail_off; i < n; i++) { uint32_t v = adler_ld8<COHERENT>(ptr + i); a += v; b += (n - i) * v; }
  }
  const uint4 *p = reinterpret_cast<const uint4 *>(ptr + head);
  for (uint32_t j = lane; j < nblk; j += 32) {
    uint4 w = adler_ld16<COHERENT>(p + j);
    uint32_t sa = __dp4a(w.x, 0x01010101u, 0u);
    sa = __dp4a(w.y, 0x01010101u, sa);
    sa = __dp4a(w.z, 0x01010101u, sa);
    sa = __dp4a(w.w, 0x01010101u, sa);
    uint32_t sw = __dp4a(w.x, 0x0D0E0F10u, 0u);  // weights 16,15,14,13 for bytes 0..3
    sw = __dp4a(w.y, 0x090A0B0Cu, sw);
    sw = __dp4a(w.z, 0x05060708u, sw);
    sw = __dp4a(w.w, 0x01020304u, sw);
    uint32_t o = head + (j << 4);           // offset of the block in the range
    a += sa;
    b += (n - o - 16) * sa + sw;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  A = a; B = b;
}

// Adler_32.bytes_update over [ptr, ptr+len) exactly as the reference chunks it (first chunk = len mod
// 5552), by a whole warp; `state` is the packed (s2 << 16) + s1 value the reference threads between calls.
template <bool COHERENT>
__device__ __forceinline__ uint32_t adler_update_warp(uint32_t state, const uint8_t *ptr, uint64_t len, int mode, int lane) {
  uint32_t s1 = state & 0xFFFFu, s2 = state >> 16;  // zipc_deflate.ml:178
  uint64_t off = 0;
  uint32_t m = (uint32_t)(len % 5552);
  if (m == 0) m = 5552;  // an empty first round is the identity on a reduced state
  while (off < len) {
    uint32_t A, B;
    adler_range_warp<COHERENT>(ptr + off, m, lane, A, B);
    adler_fold_step(s1, s2, m, A, B, mode);
    off += m;
    m = 5552;
  }
  return (s2 << 16) + s1;  // :198
}
#endif

}  // namespace zb

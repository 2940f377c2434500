// adler32.cu -- Adler-32 at HBM bandwidth, in both the reference's flavour and RFC 1950's.
//
// Replaces Zipc_deflate.Adler_32.string / string_update (reference src/zipc_deflate.ml:166-206).
// The reference walks the buffer in 5552-byte chunks (first chunk = len mod 5552), keeps s1, s2
// as wrapping int32 and reduces them with the *signed* Int32.rem at each chunk end
// (zipc_deflate.ml:95,196).  That only differs from RFC 1950 when a chunk pushes s2 past 2^31.
//
// Device part (this file, bandwidth bound): one warp per chunk computes
//      A = sum b_i                  B = sum (n - i) * b_i          (i = 0..n-1, n <= 5552)
// with dp4a over coalesced 16-byte loads.  B < 2^32 always (255*5552*5553/2 = 3.93e9).
// Fold part (0.14 % of the bytes: 8 bytes per 5552): the chunk recurrence
//      s1' = rem(s1 + A)            s2' = rem(s2 + n*s1 + B)        (int32 wrap, then rem)
// is evaluated over the per-chunk (A, B) pairs; in REF_COMPAT mode it is order dependent (sign
// of the wrapped s2) and is therefore done as a scalar scan over the partials, in RFC1950 mode it
// is the same scan with unsigned arithmetic.
//
// Algorithmic bytes per launch: N.  Roofline: HBM.
#include "common.cuh"

namespace zb {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kChunk = 5552;

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// (A, B) of one range [ptr, ptr+n), n <= 5552, valid in lane 0
__device__ __forceinline__ void adler_range_warp(const uint8_t *ptr, uint32_t n, int lane, uint32_t &A, uint32_t &B) {
  uint32_t a = 0, b = 0;
  uint32_t head = (uint32_t)((16 - ((uintptr_t)ptr & 15)) & 15);
  if (head > n) head = n;
  uint32_t nblk = (n - head) >> 4;
  uint32_t tail_off = head + (nblk << 4);
  if (lane == 0) {  // ragged ends, bytewise
    for (uint32_t i = 0; i < head; i++) { uint32_t v = ptr[i]; a += v; b += (n - i) * v; }
    for (uint32_t i = tail_off; i < n; i++) { uint32_t v = ptr[i]; a += v; b += (n - i) * v; }
  }
  const uint4 *p = reinterpret_cast<const uint4 *>(ptr + head);
  for (uint32_t j = lane; j < nblk; j += 32) {
    uint4 w = ldg_stream(p + j);
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

// chunk c of a single buffer: c == 0 -> [0, first), else [first + (c-1)*5552, +5552)   (first may
// equal 5552 when len is a multiple of it; an empty first round is a no-op and is skipped)
__global__ void __launch_bounds__(kThreads)
adler_chunks_kernel(const uint8_t *__restrict__ src, uint32_t first, uint32_t nchunks, uint2 *__restrict__ ab) {
  const int lane = threadIdx.x & 31;
  for (uint32_t c = blockIdx.x * kWarps + (threadIdx.x >> 5); c < nchunks; c += gridDim.x * kWarps) {
    uint64_t off = c == 0 ? 0 : (uint64_t)first + (uint64_t)(c - 1) * kChunk;
    uint32_t n = c == 0 ? first : kChunk;
    uint32_t A, B;
    adler_range_warp(src + off, n, lane, A, B);
    if (lane == 0) ab[c] = make_uint2(A, B);
  }
}

// arbitrary ranges (ptr, len <= 5552)
__global__ void __launch_bounds__(kThreads)
adler_ranges_kernel(const AdlerSeg *__restrict__ segs, uint32_t nseg, uint2 *__restrict__ ab) {
  const int lane = threadIdx.x & 31;
  for (uint32_t c = blockIdx.x * kWarps + (threadIdx.x >> 5); c < nseg; c += gridDim.x * kWarps) {
    uint32_t A, B;
    adler_range_warp(segs[c].ptr, segs[c].len, lane, A, B);
    if (lane == 0) ab[c] = make_uint2(A, B);
  }
}

}  // namespace

// The chunk recurrence of zipc_deflate.ml:181-197 over per-chunk partial sums.
static inline void adler_fold_step(uint32_t &s1, uint32_t &s2, uint32_t n, uint32_t A, uint32_t B, int mode) {
  uint32_t t2 = s2 + n * s1 + B;  // all int32-wrapping in the reference
  uint32_t t1 = s1 + A;
  if (mode == ZIPC_ADLER_REF_COMPAT) {
    s1 = (uint32_t)((int32_t)t1 % 65521);
    s2 = (uint32_t)((int32_t)t2 % 65521);
  } else {
    // RFC 1950: exact arithmetic.  n*s1 + B + s2 can exceed 2^32, so widen.
    uint64_t w2 = (uint64_t)s2 + (uint64_t)n * s1 + B;
    s1 = t1 % 65521u;
    s2 = (uint32_t)(w2 % 65521u);
  }
}

int adler32_launch_buffer(zipc_b200_ctx *ctx, const uint8_t *d_src, uint64_t len, int mode, uint32_t *h_out) {
  if (len == 0) { *h_out = 1; return ZIPC_OK; }
  uint32_t first = (uint32_t)(len % kChunk);
  if (first == 0) first = kChunk;
  uint64_t nchunks64 = 1 + (len - first) / kChunk;
  if (nchunks64 > 0xFFFFFFFFull) return ZIPC_ERR_INVALID_ARG;
  uint32_t nchunks = (uint32_t)nchunks64;
  if (int st = ctx->d_scratch2.reserve((size_t)nchunks * sizeof(uint2))) return st;
  if (int st = ctx->h_res.reserve((size_t)nchunks * sizeof(uint2))) return st;
  uint2 *d_ab = ctx->d_scratch2.as<uint2>();
  uint32_t grid = (nchunks + kWarps - 1) / kWarps;
  uint32_t maxgrid = (uint32_t)ctx->sm_count * 4;
  if (grid > maxgrid) grid = maxgrid;
  adler_chunks_kernel<<<grid, kThreads, 0, ctx->stream>>>(d_src, first, nchunks, d_ab);
  ctx->launches++;
  ZB_CUDA(ctx, cudaGetLastError());
  uint2 *h_ab = ctx->h_res.as<uint2>();
  ZB_CUDA(ctx, cudaMemcpyAsync(h_ab, d_ab, (size_t)nchunks * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
  ZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint32_t s1 = 1, s2 = 0;
  for (uint32_t c = 0; c < nchunks; c++) adler_fold_step(s1, s2, c == 0 ? first : kChunk, h_ab[c].x, h_ab[c].y, mode);
  *h_out = (s2 << 16) + s1;
  return ZIPC_OK;
}

int adler32_ranges(zipc_b200_ctx *ctx, const uint8_t *const *d_ptrs, const uint64_t *lens, size_t n, int mode,
                   uint32_t *h_out) {
  // chunk every range on the reference's grid (first chunk = len mod 5552)
  size_t total = 0;
  for (size_t i = 0; i < n; i++) total += lens[i] ? 1 + (lens[i] - 1) / kChunk : 0;
  if (total > 0xFFFFFFFFull) return ZIPC_ERR_INVALID_ARG;
  if (total == 0) { for (size_t i = 0; i < n; i++) h_out[i] = 1; return ZIPC_OK; }
  if (int st = ctx->h_desc.reserve(total * sizeof(AdlerSeg))) return st;
  if (int st = ctx->d_desc.reserve(total * sizeof(AdlerSeg))) return st;
  if (int st = ctx->d_scratch2.reserve(total * sizeof(uint2))) return st;
  if (int st = ctx->h_res.reserve(total * sizeof(uint2))) return st;
  AdlerSeg *h = ctx->h_desc.as<AdlerSeg>();
  size_t k = 0;
  for (size_t i = 0; i < n; i++) {
    uint64_t len = lens[i], off = 0;
    if (!len) continue;
    uint32_t first = (uint32_t)(len % kChunk);
    if (first == 0) first = kChunk;
    for (uint32_t m = first; off < len; off += m, m = kChunk) { h[k].ptr = d_ptrs[i] + off; h[k].len = m; h[k]._pad = 0; k++; }
  }
  ZB_CUDA(ctx, cudaMemcpyAsync(ctx->d_desc.p, h, total * sizeof(AdlerSeg), cudaMemcpyHostToDevice, ctx->stream));
  uint32_t grid = (uint32_t)((total + kWarps - 1) / kWarps);
  uint32_t maxgrid = (uint32_t)ctx->sm_count * 4;
  if (grid > maxgrid) grid = maxgrid;
  adler_ranges_kernel<<<grid, kThreads, 0, ctx->stream>>>(ctx->d_desc.as<AdlerSeg>(), (uint32_t)total, ctx->d_scratch2.as<uint2>());
  ctx->launches++;
  ZB_CUDA(ctx, cudaGetLastError());
  uint2 *h_ab = ctx->h_res.as<uint2>();
  ZB_CUDA(ctx, cudaMemcpyAsync(h_ab, ctx->d_scratch2.p, total * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
  ZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  k = 0;
  for (size_t i = 0; i < n; i++) {
    uint32_t s1 = 1, s2 = 0;
    uint64_t len = lens[i], off = 0;
    if (len) {
      uint32_t first = (uint32_t)(len % kChunk);
      if (first == 0) first = kChunk;
      for (uint32_t m = first; off < len; off += m, m = kChunk, k++) adler_fold_step(s1, s2, m, h_ab[k].x, h_ab[k].y, mode);
    }
    h_out[i] = (s2 << 16) + s1;
  }
  return ZIPC_OK;
}

}  // namespace zb

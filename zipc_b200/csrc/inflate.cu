// inflate.cu -- batched DEFLATE decoder for sm_100a: one *lane* per stream, warp-cooperative copies.
//
// Replaces the decode side of the reference (src/zipc_deflate.ml:532-718): read_bits (:564-579),
// the bit-at-a-time read_symbol (:584-591), read_block_symbols (:593-616), the three block readers
// (:618-680) and the driver inflate_and_crc (:692-709).  Results are bit-exact, including which of
// the two errors ("Corrupted data stream" / "Expected decompression size exceeded") a bad stream
// yields: checks happen in the reference's order token by token.
//
// Design (why it is not a translation):
//   * Huffman decoding is serial per stream, so the parallelism is across streams: every lane of a
//     warp runs the decoder state machine of its own stream (10k ZIP members = 10k independent
//     streams).  A finished lane pulls the next stream from a global queue.
//   * Symbols are decoded through per-stream lookup tables in shared memory (2^LB entries for
//     literal/length, 2^DB for distance, 16-bit entries); codes longer than the table fall back to
//     the canonical walk over per-length counts.  The fixed-Huffman tables are built once per CTA.
//   * A round decodes one token per lane (uniform control flow), then the 32 tokens' bytes are
//     flattened over the warp: lane j of pass p moves byte 32p+j of the concatenated copies, so a
//     258-byte match costs 9 coalesced passes rather than stalling 31 lanes.  Back-references read
//     the stream's own earlier output from global memory (L1/L2 resident, at most 32 KiB back).
//   * Dynamic-block headers are parsed by the owning lanes, then the warp builds that lane's tables
//     cooperatively.
// The checksum of the output is produced by the CRC-32 / Adler-32 kernels over the freshly written
// (L2-warm) output, see api.cu.
//
// Algorithmic bytes: C + U per stream (compressed read + uncompressed written).  The kernel is
// issue/latency bound (serial bit parsing), not HBM bound: see DESIGN.md.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "adler_core.cuh"

namespace zb {
namespace {

constexpr int LB = 9;        // literal/length table bits
constexpr int DB = 7;        // distance table bits (>= 6: the area also hosts the 128-entry code-length table)
constexpr int G = 8;         // streams (decoder state machines) per warp: lanes 0..G-1 lead
constexpr int K = 8;         // tokens a leader decodes per round
constexpr int WARPS = 16;    // warps per CTA (one CTA per SM: tables fill shared memory)
constexpr int THREADS = WARPS * 32;
constexpr uint32_t ENT_LONG = 0xFFFFFFFFu;  // code longer than the table: canonical walk (0xFFFF in 16-bit tables)
constexpr int SYMS_PER_SLOT = 320;     // sorted symbols: 288 litlen + 32 dist

// per-lane shared memory record
// lit entry  (16 bit): [3:0] code length, [6:4] kind (0..5 = length symbol with that many extra bits,
//                       6 = end of block, 7 = literal), [15:7] value (literal byte / length base)
// dist entry (32 bit): [3:0] code length, [7:4] extra bits, [31:8] distance base
// 0 = invalid code (corrupt stream); all ones = code longer than the table (canonical walk)
struct __align__(16) LaneTabs {
  uint16_t lit[1 << LB];
  uint32_t dist[1 << DB];
  uint16_t lit_cnt[16];
  uint16_t dist_cnt[16];
};
struct __align__(16) WarpScratch {
  uint8_t len[320];
  uint32_t cnt[16];
  uint16_t next[16];
  uint16_t symoff[16];
  int err;
};

constexpr int QN = K * G;      // tokens per warp round
constexpr int kCompBytes = (QN + 34) * 2 + (QN + 32) + 4;  // per warp: cstart u16[QN+34], cidx u8[QN+32], padded to 4
constexpr size_t kSmemBytes = sizeof(LaneTabs) * (WARPS * G + 1) + sizeof(WarpScratch) * WARPS +
                              sizeof(uint32_t) * WARPS * QN * 2 + kCompBytes * WARPS + 64 + 128 + 64;

enum : uint32_t { S_IDLE = 0, S_HDR = 1, S_DATA = 2, S_STORED = 3, S_FINISH = 4, S_EXIT = 5 };

__constant__ uint16_t c_len_tab[29] = {  // base | extra << 9   (RFC 1951 3.2.5; reference :245-255)
    3 | 0 << 9,   4 | 0 << 9,   5 | 0 << 9,   6 | 0 << 9,   7 | 0 << 9,   8 | 0 << 9,   9 | 0 << 9,   10 | 0 << 9,
    11 | 1 << 9,  13 | 1 << 9,  15 | 1 << 9,  17 | 1 << 9,  19 | 2 << 9,  23 | 2 << 9,  27 | 2 << 9,  31 | 2 << 9,
    35 | 3 << 9,  43 | 3 << 9,  51 | 3 << 9,  59 | 3 << 9,  67 | 4 << 9,  83 | 4 << 9,  99 | 4 << 9,  115 | 4 << 9,
    131 | 5 << 9, 163 | 5 << 9, 195 | 5 << 9, 227 | 5 << 9, 258 | 0 << 9};
__constant__ uint32_t c_dist_tab[30] = {  // base | extra << 16   (reference :277-288)
    1 | 0 << 16,     2 | 0 << 16,     3 | 0 << 16,      4 | 0 << 16,      5 | 1 << 16,      7 | 1 << 16,
    9 | 2 << 16,     13 | 2 << 16,    17 | 3 << 16,     25 | 3 << 16,     33 | 4 << 16,     49 | 4 << 16,
    65 | 5 << 16,    97 | 5 << 16,    129 | 6 << 16,    193 | 6 << 16,    257 | 7 << 16,    385 | 7 << 16,
    513 | 8 << 16,   769 | 8 << 16,   1025 | 9 << 16,   1537 | 9 << 16,   2049 | 10 << 16,  3073 | 10 << 16,
    4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16,  12289 | 12 << 16, 16385 | 13 << 16, 24577 | 13 << 16};
__constant__ uint8_t c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---- bit reader: two 32-bit words + a bit offset; the next 32 stream bits are one funnel shift away ----------
// After refill() at least 32 valid bits are visible; callers peek / drop at most 32 bits between refills.
struct BitReader {
  const uint32_t *wp;    // next word to load into `ahead`
  const uint32_t *wend;  // first word not to load (words past the stream read as 0)
  uint32_t lo, hi;       // current and next word
  uint32_t pos;          // bit offset of the read position inside lo (may reach 63 before refill)
  uint32_t ahead;        // the word after hi, already fetched (hides the load latency)
  long long base;        // stream bit offset of bit 0 of lo (negative inside a leading partial word)
  long long limit;       // 8 * len
  bool tail;             // the visible window may reach past the end of the input
  __device__ __forceinline__ uint32_t fetch() {
    uint32_t w = wp < wend ? *wp : 0u;
    wp++;
    return w;
  }
  __device__ __forceinline__ void seek(const uint8_t *src, uint64_t len, uint64_t byte_pos) {
    const uint8_t *p = src + byte_pos;
    uint32_t a = (uint32_t)((uintptr_t)p & 3);
    wp = reinterpret_cast<const uint32_t *>(p - a);
    wend = reinterpret_cast<const uint32_t *>(((uintptr_t)(src + len) + 3) & ~(uintptr_t)3);
    lo = fetch(); hi = fetch(); ahead = fetch();
    pos = 8 * a;
    base = (long long)(byte_pos * 8) - 8 * a;
    limit = (long long)(len * 8);
    tail = base + 96 > limit;
  }
  __device__ __forceinline__ void refill() {  // afterwards pos < 32
    if (pos >= 32) {
      pos -= 32;
      lo = hi; hi = ahead;
      ahead = fetch();
      base += 32;
      tail = base + 96 > limit;
    }
  }
  // pos may have run past 32 since the last refill (a code followed by its extra bits): then the window
  // starts inside hi and continues in the prefetched word
  __device__ __forceinline__ uint32_t window() const {
    return pos < 32 ? __funnelshift_r(lo, hi, pos) : __funnelshift_r(hi, ahead, pos - 32);
  }
  __device__ __forceinline__ uint32_t peek(uint32_t cnt) const { return window() & ((1u << cnt) - 1u); }  // cnt < 32
  __device__ __forceinline__ void drop(uint32_t cnt) { pos += cnt; }
  __device__ __forceinline__ uint64_t consumed() const { return (uint64_t)(base + pos); }
  __device__ __forceinline__ bool overrun() const { return tail && base + (long long)pos > limit; }
};

// ---- canonical walk for codes longer than the table (the reference's read_symbol, :584-591) --------
// Returns the symbol and consumes its bits, or -1 (the reference would run off counts.(16)).
__device__ __forceinline__ int canon_decode(BitReader &br, const uint16_t *cnt, const uint16_t *syms) {
  int len = 1, base = 0, offs = 0;
  uint32_t bits = br.window();
  for (; len <= 15; len++) {
    offs = 2 * offs + (int)(bits & 1u);
    bits >>= 1;
    int count = cnt[len];
    if (offs < count) { br.drop(len); return syms[base + offs]; }
    base += count;
    offs -= count;
  }
  br.drop(15);
  return -1;
}

// ---- table construction (warp cooperative) -------------------------------------------------------------
// Builds one canonical decoder from ws.len[first .. first+n) into lut (2^bits entries), cnt[16] and the
// sorted symbol list syms (global).  Follows Huffman.init_decoder (reference :355-391) for what is accepted:
// over-subscribed -> corrupt; incomplete -> corrupt unless empty or a single code of length 1.
// max_valid_sym: symbols above it decode to "corrupt" (286/287 and 30/31 of the fixed codes).
template <bool IS_DIST>
__device__ void build_decoder_warp(WarpScratch &ws, int first, int n, int bits, void *lut_v, uint16_t *cnt,
                                   uint16_t *syms, const uint16_t *s_len_tab, const uint32_t *s_dist_tab, int lane) {
  uint16_t *lut16 = reinterpret_cast<uint16_t *>(lut_v);
  uint32_t *lut32 = reinterpret_cast<uint32_t *>(lut_v);
  if (lane < 16) ws.cnt[lane] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    int l = ws.len[first + i];
    if (l) atomicAdd(&ws.cnt[l], 1u);
  }
  {  // clear the table: 0 = invalid
    const int words = IS_DIST ? (1 << bits) : (1 << bits) / 2;
    for (int i = lane; i < words; i += 32) lut32[i] = 0;
  }
  __syncwarp();
  if (lane == 0) {
    int available = 1, num_codes = 0, code = 0, bad = 0;
    for (int l = 0; l < 16; l++) {
      int used = l ? (int)ws.cnt[l] : 0;
      if (used > available) bad = 1;                      // over-subscribed (:371)
      available = 2 * (available - used);
      ws.symoff[l] = (uint16_t)num_codes;                 // start of the length class in syms[]
      num_codes += used;
      int prev_used = l > 1 ? (int)ws.cnt[l - 1] : 0;     // canonical first code (RFC 1951 3.2.2)
      code = l ? (code + prev_used) << 1 : 0;
      ws.next[l] = (uint16_t)code;
      cnt[l] = (uint16_t)used;
    }
    if ((num_codes > 1 && available > 0) || (num_codes == 1 && ws.cnt[1] != 1)) bad = 1;  // (:377-378)
    if (bad) ws.err = 1;
    else {
      uint16_t off[16];
      for (int l = 0; l < 16; l++) off[l] = ws.symoff[l];
      int max_sym = -1;
      for (int i = 0; i < n; i++) {                       // symbols sorted by (length, symbol)
        int l = ws.len[first + i];
        if (!l) continue;
        max_sym = i;
        syms[off[l]++] = (uint16_t)i;
      }
      if (num_codes == 1) { cnt[1] = 2; syms[1] = (uint16_t)(max_sym + 1); }  // padding (:389-390)
    }
  }
  __syncwarp();
  if (ws.err) return;
  // Sorted position j of length class l carries code next[l] + (j - symoff[l]); codes are stored
  // bit-reversed because the stream delivers them most significant bit first.
  int total = 0;
  for (int l = 1; l < 16; l++) total += (int)ws.cnt[l];
  for (int j = lane; j < total; j += 32) {
    int l = 15;
    for (int t = 1; t < 15; t++)
      if (j < (int)ws.symoff[t + 1]) { l = t; break; }
    int sym = syms[j];
    uint32_t code = (uint32_t)ws.next[l] + (uint32_t)(j - (int)ws.symoff[l]);
    uint32_t rev = __brev(code) >> (32 - l);
    if (l <= bits) {
      if (IS_DIST) {
        uint32_t e = 0;                                   // 30, 31 never occur in valid data (:608)
        if (sym <= 29) { uint32_t dt = s_dist_tab[sym]; e = ((dt & 0xFFFFu) << 8) | ((dt >> 16) << 4) | (uint32_t)l; }
        for (uint32_t k = rev; k < (1u << bits); k += (1u << l)) lut32[k] = e;
      } else {
        uint16_t e = 0;                                   // 286, 287 never occur in valid data (:598)
        if (sym < 256) e = (uint16_t)((sym << 7) | (7 << 4) | l);
        else if (sym == 256) e = (uint16_t)((6 << 4) | l);
        else if (sym <= 285) { uint32_t lt = s_len_tab[sym - 257]; e = (uint16_t)(((lt & 0x1FFu) << 7) | ((lt >> 9) << 4) | (uint32_t)l); }
        for (uint32_t k = rev; k < (1u << bits); k += (1u << l)) lut16[k] = e;
      }
    } else {
      if (IS_DIST) lut32[rev & ((1u << bits) - 1u)] = ENT_LONG;
      else lut16[rev & ((1u << bits) - 1u)] = 0xFFFF;
    }
  }
  __syncwarp();
}

// ---- the kernel ------------------------------------------------------------------------------------------
// Each warp serves G streams: lanes 0..G-1 ("leaders") run one decoder state machine each and decode up to
// K tokens per round into a small queue; then all 32 lanes execute the queued tokens, token index by token
// index, with the bytes of the G concurrent tokens flattened over the lanes.
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(THREADS, 1)
inflate_kernel(const InflateTask *__restrict__ tasks, uint32_t ntasks, InflateResult *__restrict__ results,
               unsigned int *__restrict__ queue, uint16_t *__restrict__ g_syms, int adler_mode, int active_warps) {
  // adler_mode: -1 = no checksum in this kernel, else ZIPC_ADLER_* (fused per-block Adler-32 of the output)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  LaneTabs *tabs = reinterpret_cast<LaneTabs *>(smem_raw);                 // [WARPS*G] + fixed
  LaneTabs &fixed = tabs[WARPS * G];
  WarpScratch *wss = reinterpret_cast<WarpScratch *>(tabs + WARPS * G + 1);
  uint32_t *tokq = reinterpret_cast<uint32_t *>(wss + WARPS);              // [WARPS][2][K][G]
  uint8_t *compq = reinterpret_cast<uint8_t *>(tokq + WARPS * QN * 2);     // [WARPS] compacted token lists
  uint16_t *s_len_tab = reinterpret_cast<uint16_t *>(compq + ((WARPS * kCompBytes + 3) & ~3));
  uint32_t *s_dist_tab = reinterpret_cast<uint32_t *>(s_len_tab + 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpScratch &ws = wss[warp];
  const bool leader = lane < G;
  LaneTabs &mine = tabs[warp * G + (leader ? lane : 0)];
  uint2 *myq = reinterpret_cast<uint2 *>(tokq + warp * QN * 2);  // [j*G+g]: x = dep << 31 | len << 16 | dist-or-byte, y = output offset inside the round
  uint16_t *cstart = reinterpret_cast<uint16_t *>(compq + warp * kCompBytes);  // byte offset of compacted token c (+ end sentinel)
  uint8_t *cidx = reinterpret_cast<uint8_t *>(cstart + QN + 34);                // queue slot of compacted token c
  const uint32_t slot = (blockIdx.x * WARPS + warp) * G + (leader ? lane : 0);
  uint16_t *my_syms = g_syms + (size_t)slot * SYMS_PER_SLOT;
  uint16_t *fixed_syms = g_syms + (size_t)(gridDim.x * WARPS * G + blockIdx.x) * SYMS_PER_SLOT;

  if (threadIdx.x < 29) s_len_tab[threadIdx.x] = c_len_tab[threadIdx.x];
  if (threadIdx.x < 30) s_dist_tab[threadIdx.x] = c_dist_tab[threadIdx.x];
  __syncthreads();
  // fixed-Huffman decoders, once per CTA (reference :334-349)
  if (warp == 0) {
    for (int i = lane; i < 320; i += 32) {
      uint8_t l = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
      ws.len[i] = l;
    }
    if (lane == 0) ws.err = 0;
    __syncwarp();
    build_decoder_warp<false>(ws, 0, 288, LB, fixed.lit, fixed.lit_cnt, fixed_syms, s_len_tab, s_dist_tab, lane);
    build_decoder_warp<true>(ws, 288, 32, DB, fixed.dist, fixed.dist_cnt, fixed_syms + 288, s_len_tab, s_dist_tab, lane);
  }
  __syncthreads();

  // Streams are packed into the first `active_warps` warps of every CTA (a warp's round costs the same
  // whether 1 or G of its leaders are busy); the other warps only helped with the fixed tables.
  if (warp >= active_warps) return;
  // per-leader decoder state (lanes >= G carry dead copies)
  uint32_t state = leader ? S_IDLE : S_EXIT, task = 0, status = ZIPC_OK;
  BitReader br{};
  const uint8_t *src = nullptr;
  uint64_t src_len = 0;
  uint8_t *dst = nullptr;
  uint64_t out_pos = 0, out_cap = 0;
  bool final_blk = false, need_build = false, first_task = true, segment = false;
  uint32_t hlit = 0, hdist = 0;
  uint32_t stored_len = 0;
  const uint8_t *stored_src = nullptr;
  uint32_t ad_state = 1;      // running Adler-32 (reference :558, :682-690)
  uint64_t ad_from = 0;       // output offset up to which it is accounted
  bool ad_pending = false;    // a block just ended: fold [ad_from, out_pos)
  const uint16_t *lit_lut = nullptr, *lit_cnt = nullptr, *dist_cnt = nullptr;
  const uint32_t *dist_lut = nullptr;
  const uint16_t *lit_syms = nullptr, *dist_syms = nullptr;

#ifdef ZB_INFLATE_TIMING
  long long tA = 0, tD = 0, tE1 = 0, tE2 = 0, tG = 0, rounds = 0, ntoks = 0, t0 = clock64();
#define ZB_TICK(acc) { long long t1 = clock64(); acc += t1 - t0; t0 = t1; }
#else
#define ZB_TICK(acc)
#endif
  for (;;) {
    // ---- A: idle leaders pull work -------------------------------------------------------------------
    if (state == S_IDLE) {
      // first task: interleaved over the CTAs so the longest streams (sorted first) spread over all SMs;
      // afterwards from the shared queue, which starts behind the statically assigned ones
      if (first_task) { task = (uint32_t)((lane * active_warps + warp) * gridDim.x + blockIdx.x); first_task = false; }
      else task = atomicAdd(queue, 1u);
      if (task < ntasks) {
        const InflateTask t = tasks[task];
        src = t.src; src_len = t.src_len; dst = t.dst; out_cap = t.dst_cap; segment = (t.flags & kInflateSegment) != 0;
        out_pos = 0; status = ZIPC_OK; final_blk = false;
        ad_state = 1; ad_from = 0; ad_pending = false;
        br.seek(src, src_len, 0);
        state = S_HDR;
      } else {
        state = S_EXIT;
      }
    }
    if (__all_sync(0xffffffffu, state == S_EXIT)) break;

    // ---- B: block headers (reference :692-702, :623-661, :671-677) ------------------------------------
    if (state == S_HDR && segment && br.consumed() == src_len * 8) {
      state = S_FINISH;  // a segment ends at the block boundary where its input ends (byte aligned by construction)
    }
    if (state == S_HDR) {
      br.refill();
      uint32_t h = br.peek(3);
      br.drop(3);
      final_blk = h & 1u;
      uint32_t type = h >> 1;
      if (br.consumed() > src_len * 8) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
      else if (type == 0) {
        br.drop((8u - ((uint32_t)br.consumed() & 7u)) & 7u);  // to the byte boundary
        br.refill();
        uint32_t v = br.window();
        br.drop(32);
        uint32_t length = v & 0xFFFFu, inv = v >> 16;
        uint64_t pos = br.consumed() >> 3;
        if (br.consumed() > src_len * 8 || length != ((~inv) & 0xFFFFu) || src_len - pos < length) {
          status = ZIPC_ERR_CORRUPTED; state = S_FINISH;
        } else if (out_pos + length > out_cap) {
          status = ZIPC_ERR_SIZE_EXCEEDED; state = S_FINISH;
        } else {
          stored_len = length;
          stored_src = src + pos;
          br.seek(src, src_len, pos + length);
          state = S_STORED;
        }
      } else if (type == 1) {
        lit_lut = fixed.lit; dist_lut = fixed.dist; lit_cnt = fixed.lit_cnt; dist_cnt = fixed.dist_cnt;
        lit_syms = fixed_syms; dist_syms = fixed_syms + 288;
        state = S_DATA;
      } else if (type == 2) {
        br.refill();
        hlit = 257 + br.peek(5); br.drop(5);
        hdist = 1 + br.peek(5); br.drop(5);
        uint32_t hclen = 4 + br.peek(4); br.drop(4);
        bool bad = hlit > 286 || hdist > 30;
        // code length code lengths, 3 bits per symbol, packed by symbol
        uint64_t clc = 0;
        for (uint32_t i = 0; i < hclen; i++) {
          br.refill();
          clc |= (uint64_t)br.peek(3) << (3 * c_clen_order[i]);
          br.drop(3);
        }
        if (br.consumed() > src_len * 8) bad = true;
        // code-length decoder (7-bit table in the distance area, 8-bit entries: len << 5 | sym)
        uint8_t *cl_lut = reinterpret_cast<uint8_t *>(mine.dist);
        uint64_t next = 0;  // next code per length, 8 bits each
        if (!bad) {
          uint64_t cnt5 = 0;  // codes per length, 5 bits each (max 19)
          int max_sym = -1;
          for (int s = 0; s < 19; s++) {
            uint32_t l = (uint32_t)(clc >> (3 * s)) & 7u;
            if (l) { cnt5 += 1ull << (5 * l); max_sym = s; }
          }
          int available = 1, num_codes = 0, code = 0;
          for (int l = 0; l < 8; l++) {
            int used = l ? (int)((cnt5 >> (5 * l)) & 31u) : 0;
            if (used > available) bad = true;
            available = 2 * (available - used);
            num_codes += used;
            int prev_used = l > 1 ? (int)((cnt5 >> (5 * (l - 1))) & 31u) : 0;
            code = l ? (code + prev_used) << 1 : 0;
            next |= (uint64_t)(code & 0xff) << (8 * l);
          }
          if ((num_codes > 1 && available > 0) || (num_codes == 1 && ((cnt5 >> 5) & 31u) != 1) || max_sym == -1)
            bad = true;
        }
        if (!bad) {
          uint32_t *z = reinterpret_cast<uint32_t *>(cl_lut);
          for (int i = 0; i < 32; i++) z[i] = 0;
          for (int s = 0; s < 19; s++) {
            uint32_t l = (uint32_t)(clc >> (3 * s)) & 7u;
            if (!l) continue;
            uint32_t code = (uint32_t)(next >> (8 * l)) & 0xffu;
            next += 1ull << (8 * l);
            uint32_t rev = __brev(code) >> (32 - l);
            for (uint32_t k = rev; k < 128; k += (1u << l)) cl_lut[k] = (uint8_t)((l << 5) | s);
          }
          // decode hlit + hdist code lengths into the (currently unused) literal table area
          uint8_t *lens = reinterpret_cast<uint8_t *>(mine.lit);
          uint32_t num = 0, total = hlit + hdist, prev = 0;
          while (num < total && !bad) {
            br.refill();
            uint32_t e = cl_lut[br.peek(7)];
            if (!e) { bad = true; break; }
            br.drop(e >> 5);
            uint32_t sym = e & 31u, rep = 1, val = sym;
            if (sym == 16) {
              if (num == 0) { bad = true; break; }
              rep = 3 + br.peek(2); br.drop(2); val = prev;
            } else if (sym == 17) { rep = 3 + br.peek(3); br.drop(3); val = 0; }
            else if (sym == 18) { rep = 11 + br.peek(7); br.drop(7); val = 0; }
            if (rep > total - num) { bad = true; break; }
            for (uint32_t r = 0; r < rep; r++) lens[num++] = (uint8_t)val;
            prev = val;
          }
          if (br.consumed() > src_len * 8) bad = true;
        }
        if (bad) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
        else { need_build = true; }
      } else {
        status = ZIPC_ERR_CORRUPTED; state = S_FINISH;
      }
    }

    // ---- C: cooperative table builds ---------------------------------------------------------------
    {
      uint32_t m = __ballot_sync(0xffffffffu, need_build);
      while (m) {
        int L = __ffs(m) - 1;
        m &= m - 1;
        LaneTabs &lt = tabs[warp * G + L];
        uint32_t nl = __shfl_sync(0xffffffffu, hlit, L), nd = __shfl_sync(0xffffffffu, hdist, L);
        uint16_t *syms = g_syms + (size_t)((blockIdx.x * WARPS + warp) * G + L) * SYMS_PER_SLOT;
        const uint8_t *lens = reinterpret_cast<const uint8_t *>(lt.lit);
        __syncwarp();
        for (uint32_t i = lane; i < 320; i += 32) ws.len[i] = i < nl + nd ? lens[i] : 0;
        if (lane == 0) ws.err = 0;
        __syncwarp();
        if (ws.len[256] == 0) { if (lane == 0) ws.err = 1; }  // no end-of-block code (:662)
        __syncwarp();
        if (!ws.err) build_decoder_warp<false>(ws, 0, (int)nl, LB, lt.lit, lt.lit_cnt, syms, s_len_tab, s_dist_tab, lane);
        __syncwarp();
        if (!ws.err) build_decoder_warp<true>(ws, (int)nl, (int)nd, DB, lt.dist, lt.dist_cnt, syms + 288, s_len_tab, s_dist_tab, lane);
        __syncwarp();
        int err = ws.err;
        if (lane == L) {
          need_build = false;
          if (err) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
          else {
            lit_lut = mine.lit; dist_lut = mine.dist; lit_cnt = mine.lit_cnt; dist_cnt = mine.dist_cnt;
            lit_syms = my_syms; dist_syms = my_syms + 288;
            state = S_DATA;
          }
        }
        __syncwarp();
      }
    }

    ZB_TICK(tA)
    // ---- D: leaders decode up to K tokens each (reference :593-616) -----------------------------------------
    uint32_t ntok = 0, depmask = 0;
    const uint64_t batch_pos = out_pos;  // output position of this leader's first queued token
    if (state == S_DATA) {
      uint32_t rel = 0;                                                     // bytes produced in this round
      uint32_t hist = out_pos < 32768 ? (uint32_t)out_pos : 32768u;         // reachable history, capped
      const uint64_t room64 = out_cap - out_pos;
      uint32_t room = room64 > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)room64;  // a round produces < 2^12 bytes
      for (int k = 0; k < K; k++) {
        br.refill();
        uint32_t e = lit_lut[br.peek(LB)];
        uint32_t kind, val;
        if ((uint16_t)(e + 1u) > 1u) { br.drop(e & 15u); kind = (e >> 4) & 7u; val = e >> 7; }   // neither 0 nor 0xFFFF
        else {
          int sym = e ? canon_decode(br, lit_cnt, lit_syms) : -1;
          if (sym < 0 || sym > 285) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; break; }
          if (sym < 256) { kind = 7; val = (uint32_t)sym; }
          else if (sym == 256) { kind = 6; val = 0; }
          else { uint32_t lt = s_len_tab[sym - 257]; kind = lt >> 9; val = lt & 0x1FFu; }
        }
        if (kind == 6) {                                                    // end of block
          if (br.overrun()) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
          else { state = final_blk ? S_FINISH : S_HDR; ad_pending = true; }
          break;
        }
        // literal and match lanes of the warp share everything below except the distance decode, so a round
        // of mixed tokens does not pay for two copies of the checks / queue write / bookkeeping
        uint32_t length = 1, dist = 0;
        if (kind != 7) {
          length = val + br.peek(kind);
          br.drop(kind);
          br.refill();
          uint32_t e2 = dist_lut[br.peek(DB)];
          if (e2 + 1u > 1u) {                                               // neither 0 nor ENT_LONG
            br.drop(e2 & 15u);
            uint32_t deb = (e2 >> 4) & 15u;
            dist = (e2 >> 8) + br.peek(deb);
            br.drop(deb);
          } else {
            int dsym = e2 ? canon_decode(br, dist_cnt, dist_syms) : -1;
            if (dsym < 0 || dsym > 29) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; break; }
            uint32_t dt = s_dist_tab[dsym];
            dist = (dt & 0xFFFFu) + br.peek(dt >> 16);
            br.drop(dt >> 16);
          }
        }
        if (br.overrun() || dist > hist) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; break; }
        if (length > room) { status = ZIPC_ERR_SIZE_EXCEEDED; state = S_FINISH; break; }
        if (!COUNT_ONLY) {
          // a match is independent of this round when all of its source bytes precede the round
          uint32_t reach = min(dist, length);                               // source bytes actually read
          bool dep = kind != 7 && dist < rel + reach;                       // pos - dist + reach > batch_pos
          depmask |= dep ? (1u << k) : 0u;
          uint32_t x = kind == 7 ? val : ((dep ? 0x80000000u : 0u) | (length << 16) | dist);
          myq[k * G + lane] = make_uint2(x, rel);
        }
        ntok = k + 1; rel += length; room -= length;
        hist = min(hist + length, 32768u);
      }
      out_pos += rel;
    }

    ZB_TICK(tD)
#ifdef ZB_INFLATE_TIMING
    rounds++; ntoks += ntok;
#endif
    // ---- E: the warp executes the queued tokens ------------------------------------------------------------
    if (!COUNT_ONLY) {
      __syncwarp();
      const unsigned long long base_ptr = (unsigned long long)(uintptr_t)(dst + batch_pos);
      // E1: literals and matches whose source lies before this round: no ordering needed, so all their
      // bytes are flattened over the lanes.  The non-empty independent tokens are first compacted (QN queue
      // slots, 2 per lane) so that a pass can find the owner of each byte with one warp OR-reduction:
      // bit (end - base - 1) of M marks where a token ends inside the 32-byte window, and a byte's owner is
      // the window's first token plus the number of ends before it.
      uint32_t l0, l1;
      {
        uint32_t n0 = __shfl_sync(0xffffffffu, ntok, lane & (G - 1));
        uint32_t e0 = myq[lane].x, e1 = myq[lane + 32].x;
        uint32_t j0 = lane / G, j1 = (lane + 32) / G;
        l0 = (j0 < n0 && !(e0 >> 31)) ? (((e0 >> 16) & 0x1FFu) ? ((e0 >> 16) & 0x1FFu) : 1u) : 0u;
        l1 = (j1 < n0 && !(e1 >> 31)) ? (((e1 >> 16) & 0x1FFu) ? ((e1 >> 16) & 0x1FFu) : 1u) : 0u;
      }
      uint32_t i0 = l0, i1 = l1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v0 = __shfl_up_sync(0xffffffffu, i0, o), v1 = __shfl_up_sync(0xffffffffu, i1, o);
        if (lane >= o) { i0 += v0; i1 += v1; }
      }
      const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
      const uint32_t total = tot0 + __shfl_sync(0xffffffffu, i1, 31);
      const uint32_t lt_mask = (1u << lane) - 1u;
      const uint32_t b0 = __ballot_sync(0xffffffffu, l0 != 0), b1 = __ballot_sync(0xffffffffu, l1 != 0);
      const uint32_t ncomp = __popc(b0) + __popc(b1);
      if (l0) { uint32_t r = __popc(b0 & lt_mask); cstart[r] = (uint16_t)(i0 - l0); cidx[r] = (uint8_t)lane; }
      if (l1) { uint32_t r = __popc(b0) + __popc(b1 & lt_mask); cstart[r] = (uint16_t)(tot0 + i1 - l1); cidx[r] = (uint8_t)(lane + 32); }
      cstart[ncomp + lane] = lane == 0 ? (uint16_t)total : (uint16_t)0xFFFF;  // end sentinel, then "never ends"
      if (lane < 2) cstart[ncomp + 32 + lane] = 0xFFFF;
      __syncwarp();
      // four passes at a time: all loads are issued before the first store, so one L2 round trip
      // covers 128 bytes of copies
      uint32_t cbase = 0;  // first compacted token that reaches into the current window
      for (uint32_t base = 0; base < total; base += 128) {
        uint8_t val[4];
        uint8_t *dq[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t wb = base + 32 * u;                 // window [wb, wb + 32)
          uint32_t b = wb + lane;
          bool act = b < total;
          uint32_t endrel = (uint32_t)cstart[cbase + lane + 1] - wb;   // end of token cbase+lane, relative to the window
          uint32_t M = __reduce_or_sync(0xffffffffu, (endrel - 1u) < 32u ? 1u << (endrel - 1u) : 0u);
          uint32_t c = cbase + __popc(M & lt_mask);
          cbase += __popc(M);
          if (!act) c = 0;
          uint32_t q = b - cstart[c];
          uint2 ent = myq[cidx[c]];
          unsigned long long d = __shfl_sync(0xffffffffu, base_ptr, (int)(cidx[c] & (G - 1)));
          uint8_t *dp = reinterpret_cast<uint8_t *>((uintptr_t)d) + ent.y;
          uint32_t mlen = (ent.x >> 16) & 0x1FFu, mdist = ent.x & 0xFFFFu;
          dq[u] = act ? dp + q : nullptr;
          val[u] = (uint8_t)mdist;
          if (act && mlen) {
            const uint8_t *sp = dp - mdist + (mdist < mlen ? q % mdist : q);
            unsigned int v;
            asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(sp) : "memory");
            val[u] = (uint8_t)v;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (dq[u]) *dq[u] = val[u];
      }
      __syncwarp();
      ZB_TICK(tE1)
      // E2: matches that read bytes produced in this round, in token order (rare for text)
      uint32_t levels = __reduce_or_sync(0xffffffffu, depmask);
      while (levels) {
        const uint32_t j = (uint32_t)__ffs((int)levels) - 1u;
        levels &= levels - 1u;
        uint2 ent2 = (leader && j < ntok) ? myq[j * G + lane] : make_uint2(0u, 0u);
        uint32_t ent = ent2.x;
        uint32_t tlen = (ent >> 31) ? ((ent >> 16) & 0x1FFu) : 0u;
        uint32_t pos = ent2.y;
        uint32_t incl = tlen;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
          uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        uint32_t excl = incl - tlen;
        uint32_t tot = __shfl_sync(0xffffffffu, incl, G - 1);
        for (uint32_t base = 0; base < tot; base += 32) {
          uint32_t g = base + lane;
          uint32_t lo = 0;  // owner = number of leaders whose inclusive sum is <= g
#pragma unroll
          for (int step = G / 2; step > 0; step >>= 1) {
            uint32_t v = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1));
            if (v <= g) lo += step;
          }
          uint32_t t = lo & (G - 1);
          uint32_t q = g - __shfl_sync(0xffffffffu, excl, (int)t);
          uint32_t oe = __shfl_sync(0xffffffffu, ent, (int)t);
          uint32_t op = __shfl_sync(0xffffffffu, pos, (int)t);
          unsigned long long d = __shfl_sync(0xffffffffu, base_ptr, (int)t);
          if (g < tot) {
            uint8_t *dp = reinterpret_cast<uint8_t *>((uintptr_t)d) + op;
            uint32_t mlen = (oe >> 16) & 0x1FFu, mdist = oe & 0xFFFFu;
            const uint8_t *sp = dp - mdist + (mdist < mlen ? q % mdist : q);
            unsigned int v;
            asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(sp) : "memory");
            dp[q] = (uint8_t)v;
          }
        }
        __syncwarp();
      }
      ZB_TICK(tE2)
      // stored blocks: one coalesced copy from the input per leader (reference :678-680)
      uint32_t sm = __ballot_sync(0xffffffffu, state == S_STORED);
      while (sm) {
        int L = __ffs(sm) - 1;
        sm &= sm - 1;
        uint32_t n = __shfl_sync(0xffffffffu, stored_len, L);
        unsigned long long s = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)stored_src, L);
        unsigned long long d = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(dst + out_pos), L);
        const uint8_t *sp = reinterpret_cast<const uint8_t *>((uintptr_t)s);
        uint8_t *dp = reinterpret_cast<uint8_t *>((uintptr_t)d);
        for (uint32_t i = lane; i < n; i += 32) dp[i] = sp[i];
      }
      __syncwarp();
    }
    if (state == S_STORED) {
      out_pos += stored_len;
      state = final_blk ? S_FINISH : S_HDR;
      ad_pending = true;
    }

    // ---- G: checksum of the block that just ended (reference inflated_block_crc, :682-690) ---------------------
    // Adler-32 restarts its 5552-byte chunk grid at every block and, as written in the reference, reduces
    // with a signed remainder, so it has to be folded block by block to stay bit-exact.
    if (!COUNT_ONLY && adler_mode >= 0) {
      uint32_t am = __ballot_sync(0xffffffffu, ad_pending);
      while (am) {
        int L = __ffs(am) - 1;
        am &= am - 1;
        unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(dst + ad_from), L);
        unsigned long long n = __shfl_sync(0xffffffffu, (unsigned long long)(out_pos - ad_from), L);
        uint32_t stt = __shfl_sync(0xffffffffu, ad_state, L);
        uint32_t upd = adler_update_warp<true>(stt, reinterpret_cast<const uint8_t *>((uintptr_t)p), n, adler_mode, lane);
        if (lane == L) { ad_state = upd; ad_from = out_pos; }
      }
    }
    ad_pending = false;
    ZB_TICK(tG)

    // ---- F: finished streams report ---------------------------------------------------------------------
    if (state == S_FINISH) {
      InflateResult r;
      r.out_len = status == ZIPC_OK ? out_pos : 0;
      r.status = status;
      r._pad = status == ZIPC_OK ? ad_state : 0;  // fused Adler-32 of the output (when requested)
      results[task] = r;
      state = S_IDLE;
    }
  }
#ifdef ZB_INFLATE_TIMING
  if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && warp < 4)
    printf("blk %d warp %d: hdr/build %lld  decode %lld  copyE1 %lld  copyE2 %lld  adler %lld  clk | rounds %lld tokens(lane0) %lld\n",
           blockIdx.x, warp, tA, tD, tE1, tE2, tG, rounds, ntoks);
#endif
}

bool g_attr_set = false;

}  // namespace

int inflate_launch(zipc_b200_ctx *ctx, const InflateTask *d_tasks, uint32_t n, InflateResult *d_results,
                   bool count_only, int adler_mode) {
  if (n == 0) return ZIPC_OK;
  if (!g_attr_set) {
    ZB_CUDA(ctx, cudaFuncSetAttribute(inflate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    ZB_CUDA(ctx, cudaFuncSetAttribute(inflate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    g_attr_set = true;
  }
  // spread the streams over all SMs first (a CTA serves up to WARPS*G at a time)
  uint32_t grid = (uint32_t)ctx->sm_count;
  if (grid > n) grid = n;
  size_t sym_bytes = (size_t)(grid * WARPS * G + grid) * SYMS_PER_SLOT * sizeof(uint16_t);
  if (int st = ctx->d_scratch.reserve(sym_bytes + 256 + 4096)) return st;
  unsigned int *queue = reinterpret_cast<unsigned int *>(ctx->d_scratch.as<uint8_t>() + sym_bytes);
  // how many warps per CTA take streams: enough for `oversub` streams per leader over the whole batch
  int oversub = 1;
  if (const char *e = getenv("ZIPC_B200_INFLATE_OVERSUB")) oversub = std::max(1, atoi(e));
  // default: spread the streams over all warps (measured: the kernel is bound by per-warp latency, a warp
  // with few busy leaders finishes its rounds sooner); ZIPC_B200_INFLATE_OVERSUB packs them instead
  int active_warps = WARPS;
  if (getenv("ZIPC_B200_INFLATE_OVERSUB")) {
    active_warps = (int)((n + (size_t)grid * G * oversub - 1) / ((size_t)grid * G * oversub));
    active_warps = std::max(1, std::min(active_warps, WARPS));
  }
  if (const char *e = getenv("ZIPC_B200_INFLATE_WARPS")) active_warps = std::max(1, std::min(atoi(e), WARPS));
  {
    unsigned int start = grid * active_warps * G;  // tasks [0, start) are assigned statically
    ZB_CUDA(ctx, cudaMemcpyAsync(queue, &start, sizeof start, cudaMemcpyHostToDevice, ctx->stream));
  }
  KernelTimer kt(ctx);
  if (count_only)
    inflate_kernel<true><<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint16_t>(), -1, active_warps);
  else
    inflate_kernel<false><<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint16_t>(), adler_mode, active_warps);
  ctx->launches++;
  ZB_CUDA(ctx, cudaGetLastError());
  return ZIPC_OK;
}

}  // namespace zb

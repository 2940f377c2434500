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
#include "common.cuh"

namespace zb {
namespace {

constexpr int LB = 9;        // literal/length table bits
constexpr int DB = 7;        // distance table bits (>= 6: the area also hosts the 128-entry code-length table)
constexpr int WARPS = 4;     // warps per CTA (one CTA per SM: tables fill shared memory)
constexpr int THREADS = WARPS * 32;
constexpr uint16_t ENT_LONG = 0xFFFF;  // code longer than the table: canonical walk
constexpr int SYMS_PER_SLOT = 320;     // sorted symbols: 288 litlen + 32 dist

// per-lane shared memory record
struct __align__(16) LaneTabs {
  uint16_t lit[1 << LB];
  uint16_t dist[1 << DB];
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

constexpr size_t kSmemBytes = sizeof(LaneTabs) * (THREADS + 1) + sizeof(WarpScratch) * WARPS + 256;

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

// ---- bit reader: 64-bit buffer refilled with aligned 32-bit words --------------------------------
struct BitReader {
  const uint32_t *wp;    // next word to load
  const uint32_t *wend;  // first word not to load (words past the stream read as 0)
  uint64_t buf;
  uint32_t n;            // valid bits in buf
  uint64_t loaded;       // stream bits loaded so far (can exceed 8*len by < 64+32 bits)
  __device__ __forceinline__ void seek(const uint8_t *base, uint64_t len, uint64_t byte_pos) {
    const uint8_t *p = base + byte_pos;
    uint32_t a = (uint32_t)((uintptr_t)p & 3);
    wp = reinterpret_cast<const uint32_t *>(p - a);
    wend = reinterpret_cast<const uint32_t *>(((uintptr_t)(base + len) + 3) & ~(uintptr_t)3);
    uint32_t w = wp < wend ? *wp : 0u;
    wp++;
    buf = (uint64_t)(w >> (8 * a));
    n = 32 - 8 * a;
    loaded = byte_pos * 8 + n;
  }
  __device__ __forceinline__ void refill() {  // afterwards n >= 33
    if (n <= 32) {
      uint32_t w = wp < wend ? *wp : 0u;
      wp++;
      buf |= (uint64_t)w << n;
      n += 32;
      loaded += 32;
    }
  }
  __device__ __forceinline__ uint32_t peek(uint32_t cnt) const { return (uint32_t)buf & ((1u << cnt) - 1u); }
  __device__ __forceinline__ void drop(uint32_t cnt) { buf >>= cnt; n -= cnt; }
  __device__ __forceinline__ uint64_t consumed() const { return loaded - n; }
};

// ---- canonical walk for codes longer than the table (the reference's read_symbol, :584-591) --------
// Returns the symbol and consumes its bits, or -1 (the reference would run off counts.(16)).
__device__ __forceinline__ int canon_decode(BitReader &br, const uint16_t *cnt, const uint16_t *syms) {
  int len = 1, base = 0, offs = 0;
  uint32_t bits = (uint32_t)br.buf;
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
__device__ void build_decoder_warp(WarpScratch &ws, int first, int n, int bits, uint16_t *lut, uint16_t *cnt,
                                   uint16_t *syms, int max_valid_sym, int lane) {
  if (lane < 16) ws.cnt[lane] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    int l = ws.len[first + i];
    if (l) atomicAdd(&ws.cnt[l], 1u);
  }
  uint32_t *lut32 = reinterpret_cast<uint32_t *>(lut);  // clear the table: 0 = invalid
  for (int i = lane; i < (1 << bits) / 2; i += 32) lut32[i] = 0;
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
      uint16_t e = sym > max_valid_sym ? (uint16_t)0 : (uint16_t)((l << 12) | sym);
      for (uint32_t k = rev; k < (1u << bits); k += (1u << l)) lut[k] = e;
    } else {
      lut[rev & ((1u << bits) - 1u)] = ENT_LONG;
    }
  }
  __syncwarp();
}

// ---- the kernel ------------------------------------------------------------------------------------------
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(THREADS, 1)
inflate_kernel(const InflateTask *__restrict__ tasks, uint32_t ntasks, InflateResult *__restrict__ results,
               unsigned int *__restrict__ queue, uint16_t *__restrict__ g_syms) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  LaneTabs *tabs = reinterpret_cast<LaneTabs *>(smem_raw);
  LaneTabs &fixed = tabs[THREADS];
  WarpScratch *wss = reinterpret_cast<WarpScratch *>(tabs + THREADS + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpScratch &ws = wss[warp];
  LaneTabs &mine = tabs[threadIdx.x];
  const uint32_t slot = blockIdx.x * THREADS + threadIdx.x;
  uint16_t *my_syms = g_syms + (size_t)slot * SYMS_PER_SLOT;
  uint16_t *fixed_syms = g_syms + (size_t)(gridDim.x * THREADS + blockIdx.x) * SYMS_PER_SLOT;

  // fixed-Huffman decoders, once per CTA (reference :334-349)
  if (warp == 0) {
    for (int i = lane; i < 320; i += 32) {
      uint8_t l = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
      ws.len[i] = l;
    }
    if (lane == 0) ws.err = 0;
    __syncwarp();
    build_decoder_warp(ws, 0, 288, LB, fixed.lit, fixed.lit_cnt, fixed_syms, 285, lane);
    build_decoder_warp(ws, 288, 32, DB, fixed.dist, fixed.dist_cnt, fixed_syms + 288, 29, lane);
  }
  __syncthreads();

  // per-lane decoder state
  uint32_t state = S_IDLE, task = 0, status = ZIPC_OK;
  BitReader br{};
  const uint8_t *src = nullptr;
  uint64_t src_len = 0;
  uint8_t *dst = nullptr;
  uint64_t out_pos = 0, out_cap = 0;
  bool final_blk = false, need_build = false;
  uint32_t hlit = 0, hdist = 0;
  uint32_t stored_left = 0;
  const uint16_t *lit_lut = nullptr, *dist_lut = nullptr, *lit_cnt = nullptr, *dist_cnt = nullptr;
  const uint16_t *lit_syms = nullptr, *dist_syms = nullptr;

  for (;;) {
    // ---- A: idle lanes pull work ---------------------------------------------------------------------
    if (state == S_IDLE) {
      task = atomicAdd(queue, 1u);
      if (task < ntasks) {
        const InflateTask t = tasks[task];
        src = t.src; src_len = t.src_len; dst = t.dst; out_cap = t.dst_cap;
        out_pos = 0; status = ZIPC_OK; final_blk = false;
        br.seek(src, src_len, 0);
        state = S_HDR;
      } else {
        state = S_EXIT;
      }
    }
    if (__all_sync(0xffffffffu, state == S_EXIT)) break;

    // ---- B: block headers (reference :692-702, :623-661, :671-677) ------------------------------------
    if (state == S_HDR) {
      br.refill();
      uint32_t h = br.peek(3);
      br.drop(3);
      final_blk = h & 1u;
      uint32_t type = h >> 1;
      if (br.consumed() > src_len * 8) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
      else if (type == 0) {
        br.drop(br.n & 7u);  // to the byte boundary
        br.refill();
        uint32_t v = (uint32_t)br.buf;
        br.drop(32);
        uint32_t length = v & 0xFFFFu, inv = v >> 16;
        uint64_t pos = br.consumed() >> 3;
        if (br.consumed() > src_len * 8 || length != ((~inv) & 0xFFFFu) || src_len - pos < length) {
          status = ZIPC_ERR_CORRUPTED; state = S_FINISH;
        } else if (out_pos + length > out_cap) {
          status = ZIPC_ERR_SIZE_EXCEEDED; state = S_FINISH;
        } else {
          stored_left = length;
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
        LaneTabs &lt = tabs[warp * 32 + L];
        uint32_t nl = __shfl_sync(0xffffffffu, hlit, L), nd = __shfl_sync(0xffffffffu, hdist, L);
        uint16_t *syms = g_syms + (size_t)(blockIdx.x * THREADS + warp * 32 + L) * SYMS_PER_SLOT;
        const uint8_t *lens = reinterpret_cast<const uint8_t *>(lt.lit);
        __syncwarp();
        for (uint32_t i = lane; i < 320; i += 32) ws.len[i] = i < nl + nd ? lens[i] : 0;
        if (lane == 0) ws.err = 0;
        __syncwarp();
        if (ws.len[256] == 0) { if (lane == 0) ws.err = 1; }  // no end-of-block code (:662)
        __syncwarp();
        if (!ws.err) build_decoder_warp(ws, 0, (int)nl, LB, lt.lit, lt.lit_cnt, syms, 285, lane);
        __syncwarp();
        if (!ws.err) build_decoder_warp(ws, (int)nl, (int)nd, DB, lt.dist, lt.dist_cnt, syms + 288, 29, lane);
        __syncwarp();
        int err = ws.err;
        __threadfence_block();
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

    // ---- D: one token per lane (reference :593-616) ----------------------------------------------------
    uint32_t tok_len = 0, tok_per = 0;
    const uint8_t *tok_src = nullptr;
    uint8_t *tok_dst = nullptr;
    if (state == S_DATA) {
      br.refill();
      uint32_t e = lit_lut[br.peek(LB)];
      int sym;
      if (e == ENT_LONG) sym = canon_decode(br, lit_cnt, lit_syms);
      else if (e == 0) sym = -1;
      else { sym = (int)(e & 0x1FFu); br.drop(e >> 12); }
      if (sym < 0 || sym > 285) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
      else if (sym < 256) {
        if (br.consumed() > src_len * 8) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
        else if (out_pos + 1 > out_cap) { status = ZIPC_ERR_SIZE_EXCEEDED; state = S_FINISH; }
        else {
          if (!COUNT_ONLY) dst[out_pos] = (uint8_t)sym;
          out_pos++;
        }
      } else if (sym == 256) {
        if (br.consumed() > src_len * 8) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
        else state = final_blk ? S_FINISH : S_HDR;
      } else {
        uint32_t lt = c_len_tab[sym - 257];
        uint32_t eb = lt >> 9;
        uint32_t length = (lt & 0x1FFu) + br.peek(eb);
        br.drop(eb);
        br.refill();
        uint32_t e2 = dist_lut[br.peek(DB)];
        int dsym;
        if (e2 == ENT_LONG) dsym = canon_decode(br, dist_cnt, dist_syms);
        else if (e2 == 0) dsym = -1;
        else { dsym = (int)(e2 & 0x1FFu); br.drop(e2 >> 12); }
        if (dsym < 0 || dsym > 29) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
        else {
          uint32_t dt = c_dist_tab[dsym];
          uint32_t deb = dt >> 16;
          uint32_t dist = (dt & 0xFFFFu) + br.peek(deb);
          br.drop(deb);
          if (br.consumed() > src_len * 8 || (uint64_t)dist > out_pos) { status = ZIPC_ERR_CORRUPTED; state = S_FINISH; }
          else if (out_pos + length > out_cap) { status = ZIPC_ERR_SIZE_EXCEEDED; state = S_FINISH; }
          else {
            if (!COUNT_ONLY) {
              tok_len = length;
              tok_dst = dst + out_pos;
              tok_src = tok_dst - dist;
              tok_per = dist < length ? dist : 0u;
            }
            out_pos += length;
          }
        }
      }
    } else if (state == S_STORED) {
      // the whole stored block as one copy from the input (reference :678-680)
      uint64_t pos = br.consumed() >> 3;
      if (!COUNT_ONLY) {
        tok_len = stored_left;
        tok_dst = dst + out_pos;
        tok_src = src + pos;
        tok_per = 0;
      }
      out_pos += stored_left;
      br.seek(src, src_len, pos + stored_left);
      state = final_blk ? S_FINISH : S_HDR;
    }

    // ---- E: flattened warp copy of this round's tokens ----------------------------------------------------
    if (!COUNT_ONLY) {
      uint32_t incl = tok_len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      uint32_t excl = incl - tok_len;
      uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      for (uint32_t base = 0; base < total; base += 32) {
        uint32_t g = base + lane;
        // owner = number of lanes whose inclusive sum is <= g
        uint32_t lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          uint32_t v = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1));
          if (v <= g) lo += step;
        }
        uint32_t t = lo & 31u;
        uint32_t q = g - __shfl_sync(0xffffffffu, excl, (int)t);
        unsigned long long s = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tok_src, (int)t);
        unsigned long long d = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tok_dst, (int)t);
        uint32_t per = __shfl_sync(0xffffffffu, tok_per, (int)t);
        if (g < total) {
          uint32_t idx = per ? q % per : q;
          uint8_t b = *(reinterpret_cast<const volatile uint8_t *>((uintptr_t)s) + idx);
          *(reinterpret_cast<uint8_t *>((uintptr_t)d) + q) = b;
        }
      }
    }
    __syncwarp();

    // ---- F: finished streams report ---------------------------------------------------------------------
    if (state == S_FINISH) {
      InflateResult r;
      r.out_len = status == ZIPC_OK ? out_pos : 0;
      r.status = status;
      r._pad = 0;
      results[task] = r;
      state = S_IDLE;
    }
  }
}

bool g_attr_set = false;

}  // namespace

int inflate_launch(zipc_b200_ctx *ctx, const InflateTask *d_tasks, uint32_t n, InflateResult *d_results,
                   bool count_only) {
  if (n == 0) return ZIPC_OK;
  if (!g_attr_set) {
    ZB_CUDA(ctx, cudaFuncSetAttribute(inflate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    ZB_CUDA(ctx, cudaFuncSetAttribute(inflate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    g_attr_set = true;
  }
  uint32_t grid = (n + THREADS - 1) / THREADS;
  if (grid > (uint32_t)ctx->sm_count) grid = (uint32_t)ctx->sm_count;
  size_t sym_bytes = (size_t)(grid * THREADS + grid) * SYMS_PER_SLOT * sizeof(uint16_t);
  if (int st = ctx->d_scratch.reserve(sym_bytes + 256)) return st;
  unsigned int *queue = reinterpret_cast<unsigned int *>(ctx->d_scratch.as<uint8_t>() + sym_bytes);
  ZB_CUDA(ctx, cudaMemsetAsync(queue, 0, sizeof(unsigned int), ctx->stream));
  if (count_only)
    inflate_kernel<true><<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint16_t>());
  else
    inflate_kernel<false><<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint16_t>());
  ctx->launches++;
  ZB_CUDA(ctx, cudaGetLastError());
  return ZIPC_OK;
}

}  // namespace zb

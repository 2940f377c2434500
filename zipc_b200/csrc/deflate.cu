// deflate.cu -- batched DEFLATE encoder for sm_100a: one CTA per ZIP member (or independent segment).
//
// Replaces the encode side of the reference (src/zipc_deflate.ml:742-1277): Lz77.compress and its
// hash-chain matcher (:1140-1245), the block writer and block-type choice (:873-1104) and the deflate /
// crc_32_and_deflate / zlib_compress entry points (:1247-1277).  The output is valid RFC 1951, inflates to
// the input bit-exactly through the reference's inflate, and lands within about 1 % of the reference's
// compressed size per level (measured with the host model of the same algorithm).
//
// Per member, tile by tile (2048 input bytes), everything in shared memory:
//   1. stage the next tile into a 64 KiB input ring,
//   2. hash every position; partition the tile's positions by hash class so that 16 warps insert them into
//      the head / prev chain tables concurrently yet in exact position order (__match_any_sync resolves
//      equal hashes inside a 32-lane batch), which reproduces the serial chain semantics,
//   3. search the chain of EVERY position in parallel (4-byte compares, early rejection on the byte that
//      would have to improve the match),
//   4. resolve the one-step lazy parse (reference :1224-1241) as a successor function over (position,
//      pending?) nodes and follow it with pointer jumping, 11 doubling rounds per tile,
//   5. compact the visited nodes into tokens (CTA prefix sum), update the literal/length and distance
//      histograms with shared-memory atomics.
// Every 30 tiles (61440 bytes) the block is closed: Huffman lengths from bitonic-sorted frequencies
// (length-limited, complete), exact stored / fixed / dynamic cost comparison, then header and tokens go
// through a CTA-wide bit packer (per-item bit lengths -> prefix sum -> atomicOr into a shared staging
// window -> coalesced 32-bit stores).
//
// Algorithmic bytes: U + C per member.  Issue/latency bound (hash-chain walks, serial Huffman step), not
// HBM bound: DESIGN.md.
#include <algorithm>
#include <numeric>

#include "common.cuh"
#include "deflate_core.h"

namespace zb {
namespace {

using namespace dfl;

#ifndef ZB_DEFLATE_THREADS
#define ZB_DEFLATE_THREADS 1024
#endif
constexpr int THREADS = ZB_DEFLATE_THREADS;   // 512 or 1024
constexpr int NWARPS = THREADS / 32;
constexpr int PPT = kTile / THREADS;  // positions per thread per tile
constexpr int kClasses = NWARPS;      // hash classes of the chain insertion: one warp each
constexpr int kClassBits = NWARPS == 32 ? 5 : 4;
constexpr int PPW = kTile / NWARPS;   // positions a warp ranks in the partition step
constexpr int SORTN = 512;            // keys of the literal/length frequency sort
static_assert(THREADS == 512 || THREADS == 1024, "the kernel is written for 16 or 32 warps");
constexpr int kTokCap = 65536;        // tokens per block scratch (block source <= 61440 + 258)
constexpr int kChunk = 2048;          // items per bit-packer chunk
constexpr int IPT = kChunk / THREADS; // items per thread per chunk
constexpr int kStageWords = kChunk * 48 / 32 + 8;

// ---- shared memory map ------------------------------------------------------------------------------------
constexpr int OFF_RING = 0;                                  // u32[16384]
constexpr int OFF_PREV = OFF_RING + kRing;                   // u16[32768]
constexpr int OFF_HEAD = OFF_PREV + kWindow * 2;             // u16[1 << kHashBits]
constexpr int OFF_X = OFF_HEAD + (2 << kHashBits);           // parse arrays | bit staging | header items
constexpr int X_BYTES = 2 * kTile * 2 * 2 + 2 * kTile;       //   jumpA, jumpB (u16[2T]) + mark (u8[2T]) = 20480
constexpr int OFF_Y = OFF_X + X_BYTES;                       // tile arrays | block-finalize scratch
constexpr int Y_BYTES = kTile * 2 + (kTile + 8) * 2 * 2 + kTile * 2;  // first, mlen, mdist, poslist
constexpr int OFF_HIST = OFF_Y + Y_BYTES;                    // u32[288] lit, u32[32] dist
constexpr int OFF_CODE = OFF_HIST + 320 * 4;                 // u32[288] lit codes, u32[32] dist codes
constexpr int OFF_MISC = OFF_CODE + 320 * 4;                 // class counters, scan scratch, scalars
constexpr int MISC_BYTES = 4096;
constexpr int kSmemBytes = OFF_MISC + MISC_BYTES;
static_assert(kStageWords * 4 + 704 * 5 + 64 <= X_BYTES, "bit staging + header items must fit the parse area");
static_assert(512 * 4 * 2 + 320 * 2 + 1024 <= Y_BYTES, "finalize scratch must fit the tile area");

struct Shared {
  uint32_t *ring;
  uint8_t *ringb;
  uint16_t *prev, *head;
  // parse view of X
  uint16_t *jumpA, *jumpB;
  uint8_t *mark;
  uint16_t *hsh;  // aliases jumpA (hashes are dead before the parse starts)
  // emit view of X
  uint32_t *stage, *hdr_val;
  uint8_t *hdr_nb;
  // tile view of Y
  uint16_t *first, *mlen, *mdist, *poslist;
  // finalize view of Y
  uint32_t *keys, *hscratch;
  uint16_t *rsyms;
  uint8_t *ll, *dl, *both, *cl;
  uint32_t *dkeys, *dscratch, *ckeys, *cscratch, *cfreq;
  uint32_t *hist_l, *hist_d, *lcode, *dcode;
  uint16_t *cnt;     // [NWARPS][kClasses] per-warp class counts, then offsets
  uint16_t *cstart;  // [kClasses + 1] class starts
  uint32_t *scan;    // [NWARPS + 1]
  uint32_t *sc;      // scalars
};

// scalar slots in sh.sc
enum { SC_TASK = 0, SC_OUTW, SC_CARRY, SC_CBITS, SC_OVERFLOW, SC_M_L, SC_M_D, SC_NHDR, SC_BTYPE, SC_HLIT, SC_HDIST,
       SC_EXIT, SC_CARRY_LEN, SC_CARRY_DIST, SC_SUMDYN, SC_SUMFIX, SC_BLK_SRCLEN, SC_NRSYM, SC_HCLEN, SC_HDRBITS, SC_BATCH };

__device__ __forceinline__ Shared carve(uint8_t *base) {
  Shared s;
  s.ring = reinterpret_cast<uint32_t *>(base + OFF_RING);
  s.ringb = base + OFF_RING;
  s.prev = reinterpret_cast<uint16_t *>(base + OFF_PREV);
  s.head = reinterpret_cast<uint16_t *>(base + OFF_HEAD);
  uint8_t *x = base + OFF_X;
  s.jumpA = reinterpret_cast<uint16_t *>(x);
  s.jumpB = s.jumpA + 2 * kTile;
  s.mark = reinterpret_cast<uint8_t *>(s.jumpB + 2 * kTile);
  s.hsh = s.jumpA;
  s.stage = reinterpret_cast<uint32_t *>(x);
  s.hdr_val = s.stage + kStageWords;
  s.hdr_nb = reinterpret_cast<uint8_t *>(s.hdr_val + 704);
  uint8_t *y = base + OFF_Y;
  s.first = reinterpret_cast<uint16_t *>(y);
  s.mlen = s.first + kTile;
  s.mdist = s.mlen + kTile + 8;
  s.poslist = s.mdist + kTile + 8;
  s.keys = reinterpret_cast<uint32_t *>(y);
  s.hscratch = s.keys + 512;
  s.rsyms = reinterpret_cast<uint16_t *>(s.hscratch + 512);
  s.ll = reinterpret_cast<uint8_t *>(s.rsyms + 320);
  s.dl = s.ll + 288;
  s.both = s.dl + 32;
  s.cl = s.both + 320;
  s.dkeys = reinterpret_cast<uint32_t *>(s.cl + 32);
  s.dscratch = s.dkeys + 32;
  s.ckeys = s.dscratch + 32;
  s.cscratch = s.ckeys + 32;
  s.cfreq = s.cscratch + 32;
  s.hist_l = reinterpret_cast<uint32_t *>(base + OFF_HIST);
  s.hist_d = s.hist_l + 288;
  s.lcode = reinterpret_cast<uint32_t *>(base + OFF_CODE);
  s.dcode = s.lcode + 288;
  uint8_t *m = base + OFF_MISC;
  s.cnt = reinterpret_cast<uint16_t *>(m);
  s.cstart = s.cnt + NWARPS * kClasses;
  s.scan = reinterpret_cast<uint32_t *>(m + 2048 + 128);
  s.sc = s.scan + 40;
  return s;
}

// lanes of the warp whose key equals mine, from one ballot per key bit (match.any takes a hardware loop over
// the distinct values: ~400 clk for 32 different hashes; this is BITS ballots + logic ops)
template <int BITS>
__device__ __forceinline__ uint32_t match_any_bits(uint32_t key) {
  uint32_t m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < BITS; b++) {
    const uint32_t v = __ballot_sync(0xffffffffu, (key >> b) & 1u);
    m &= ((key >> b) & 1u) ? v : ~v;
  }
  return m;
}

struct RingView {
  const uint32_t *w;
  __device__ __forceinline__ uint32_t word(uint32_t a) const { return w[a]; }
  __device__ __forceinline__ uint32_t byte(uint32_t i) const { return reinterpret_cast<const uint8_t *>(w)[i]; }
};
struct PrevView {
  const uint16_t *l;
  __device__ __forceinline__ uint32_t link(uint32_t pos) const { return l[pos & (kWindow - 1)]; }
};

// exclusive prefix sum of one value per thread over the CTA; *total gets the sum
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t *scratch, uint32_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // scratch may still be read by a previous scan
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NWARPS ? scratch[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < NWARPS) scratch[lane] = wi - w;
    if (lane == NWARPS - 1) scratch[NWARPS] = wi;
  }
  __syncthreads();
  *total = scratch[NWARPS];
  return scratch[warp] + incl - v;
}

// ---- bit packer ----------------------------------------------------------------------------------------------
// Appends `count` items to the member's bit stream.  fetch(i, bits, nbits) yields item i (nbits <= 48).
template <class Fetch>
__device__ void pack_items(const Shared &sh, uint32_t *out_words, uint64_t out_cap_words, uint32_t count, Fetch fetch) {
  const int tid = threadIdx.x;
  for (uint32_t c0 = 0; c0 < count; c0 += kChunk) {
    uint64_t bits[IPT];
    uint32_t nb[IPT], mine = 0;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      uint32_t i = c0 + tid * IPT + k;
      bits[k] = 0; nb[k] = 0;
      if (i < count) fetch(i, bits[k], nb[k]);
      mine += nb[k];
    }
    for (int i = tid; i < kStageWords; i += THREADS) sh.stage[i] = 0;
    uint32_t total;
    uint32_t off = block_scan(mine, sh.scan, &total);  // (contains the barriers that order the zeroing)
    const uint32_t cbits = sh.sc[SC_CBITS];
    if (tid == 0) sh.stage[0] = sh.sc[SC_CARRY];
    __syncthreads();
    off += cbits;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      if (!nb[k]) continue;
      uint32_t w = off >> 5, s = off & 31;
      uint64_t lo = bits[k] << s;
      atomicOr(&sh.stage[w], (uint32_t)lo);
      if (s + nb[k] > 32) atomicOr(&sh.stage[w + 1], (uint32_t)(lo >> 32));
      if (s + nb[k] > 64) atomicOr(&sh.stage[w + 2], (uint32_t)(bits[k] >> (64 - s)));
      off += nb[k];
    }
    __syncthreads();
    const uint32_t end_bits = cbits + total, full = end_bits >> 5;
    const uint64_t outw = sh.sc[SC_OUTW];
    if (outw + full > out_cap_words) { if (tid == 0) sh.sc[SC_OVERFLOW] = 1; }
    else for (uint32_t i = tid; i < full; i += THREADS) out_words[outw + i] = sh.stage[i];
    uint32_t carry = sh.stage[full];
    __syncthreads();
    if (tid == 0) {
      sh.sc[SC_OUTW] = (uint32_t)(outw + full);
      sh.sc[SC_CARRY] = (end_bits & 31) ? carry : 0;
      sh.sc[SC_CBITS] = end_bits & 31;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t fixed_lit_code(uint32_t sym) {  // bit-reversed | len << 16 (RFC 1951 3.2.6)
  uint32_t code, len;
  if (sym < 144) { code = 0x30 + sym; len = 8; }
  else if (sym < 256) { code = 0x190 + (sym - 144); len = 9; }
  else if (sym < 280) { code = sym - 256; len = 7; }
  else { code = 0xC0 + (sym - 280); len = 8; }
  return (__brev(code) >> (32 - len)) | (len << 16);
}

// ---- block finalisation -----------------------------------------------------------------------------------
// Closes the current block: builds codes, picks the block type and appends its bits.
__device__ void finalize_block(const Shared &sh, const uint8_t *src, uint64_t blk_src_start, uint32_t ntok,
                               const uint32_t *toks, bool final, uint32_t *out_words, uint64_t out_cap_words) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (tid == 0) sh.hist_l[256] += 1;  // end of block symbol (reference :1088-1092)
  __syncthreads();
  // -- literal/length frequencies: bitonic sort of 512 keys (freq << 9 | sym; unused symbols sort last)
  if (tid < SORTN) sh.keys[tid] = (tid < kNumLit && sh.hist_l[tid]) ? ((sh.hist_l[tid] << 9) | (uint32_t)tid) : 0xFFFFFFFFu;
  __syncthreads();
  for (int k = 2; k <= SORTN; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      int partner = tid ^ j;
      if (tid < SORTN && partner > tid) {
        uint32_t a = sh.keys[tid], b = sh.keys[partner];
        bool up = (tid & k) == 0;
        if ((a > b) == up) { sh.keys[tid] = b; sh.keys[partner] = a; }
      }
      __syncthreads();
    }
  if (tid < SORTN && sh.keys[tid] != 0xFFFFFFFFu && (tid == SORTN - 1 || sh.keys[tid + 1] == 0xFFFFFFFFu)) sh.sc[SC_M_L] = tid + 1;
  // -- distance frequencies: one warp, shuffle bitonic sort of 32 keys
  if (warp == 1) {
    uint32_t key = (lane < kNumDist && sh.hist_d[lane]) ? ((sh.hist_d[lane] << 9) | (uint32_t)lane) : 0xFFFFFFFFu;
    for (int k = 2; k <= 32; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        uint32_t o = __shfl_xor_sync(0xffffffffu, key, j);
        bool up = (lane & k) == 0, lower = (lane & j) == 0;
        key = (lower == up) ? min(key, o) : max(key, o);
      }
    sh.dkeys[lane] = key;
    uint32_t used = __ballot_sync(0xffffffffu, key != 0xFFFFFFFFu);
    if (lane == 0) sh.sc[SC_M_D] = __popc(used);
  }
  __syncthreads();
  // -- code lengths (serial per alphabet, two alphabets side by side)
  if (tid == 0) huff_lengths_from_sorted(sh.keys, (int)sh.sc[SC_M_L], kNumLit, 15, sh.ll, sh.hscratch);
  if (tid == 32) huff_lengths_from_sorted(sh.dkeys, (int)sh.sc[SC_M_D], kNumDist, 15, sh.dl, sh.dscratch);
  __syncthreads();
  // -- exact symbol costs of the dynamic and fixed alternatives (reference :1049-1069), in parallel
  {
    uint32_t dyn = 0, fix = 0;
    if (tid < kNumLit) {
      uint32_t eb = tid >= 257 ? len_extra_bits_of_sym((uint32_t)tid) : 0, f = sh.hist_l[tid];
      dyn = f * (sh.ll[tid] + eb); fix = f * (fixed_lit_len((uint32_t)tid) + eb);
    } else if (tid >= 288 && tid < 288 + kNumDist) {
      uint32_t s = tid - 288, eb = dist_extra_bits_of_sym(s), f = sh.hist_d[s];
      dyn = f * (sh.dl[s] + eb); fix = f * (5 + eb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dyn += __shfl_xor_sync(0xffffffffu, dyn, o); fix += __shfl_xor_sync(0xffffffffu, fix, o); }
    if (tid == 0) { sh.sc[SC_SUMDYN] = 0; sh.sc[SC_SUMFIX] = 0; }
    __syncthreads();
    if (lane == 0) { atomicAdd(&sh.sc[SC_SUMDYN], dyn); atomicAdd(&sh.sc[SC_SUMFIX], fix); }
  }
  // -- code length alphabet, header cost, block type (thread 0; reference :959-1043, :1071-1104)
  if (tid == 0) {
    int hlit = kNumLit;
    while (hlit > 257 && sh.ll[hlit - 1] == 0) hlit--;
    int hdist = kNumDist;
    while (hdist > 1 && sh.dl[hdist - 1] == 0) hdist--;
    for (int i = 0; i < hlit; i++) sh.both[i] = sh.ll[i];
    for (int i = 0; i < hdist; i++) sh.both[hlit + i] = sh.dl[i];
    int nr = rle_code_lengths(sh.both, hlit + hdist, sh.rsyms, sh.cfreq);
    int m = 0;  // insertion sort of the (at most 19) used code length symbols
    for (int s = 0; s < kNumClen; s++) {
      if (!sh.cfreq[s]) continue;
      uint32_t key = (sh.cfreq[s] << 9) | (uint32_t)s;
      int j = m++;
      while (j > 0 && sh.ckeys[j - 1] > key) { sh.ckeys[j] = sh.ckeys[j - 1]; j--; }
      sh.ckeys[j] = key;
    }
    huff_lengths_from_sorted(sh.ckeys, m, kNumClen, 7, sh.cl, sh.cscratch);
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && sh.cl[order[hclen - 1]] == 0) hclen--;
    uint32_t hdr = 3 + 5 + 5 + 4 + 3 * (uint32_t)hclen;
    for (int s = 0; s < kNumClen; s++) hdr += sh.cfreq[s] * (sh.cl[s] + (s == 16 ? 2u : s == 17 ? 3u : s == 18 ? 7u : 0u));
    sh.sc[SC_HLIT] = (uint32_t)hlit; sh.sc[SC_HDIST] = (uint32_t)hdist; sh.sc[SC_NRSYM] = (uint32_t)nr;
    sh.sc[SC_HCLEN] = (uint32_t)hclen; sh.sc[SC_HDRBITS] = hdr;
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t src_len = sh.sc[SC_BLK_SRCLEN];
    uint64_t dlen = (uint64_t)sh.sc[SC_HDRBITS] + sh.sc[SC_SUMDYN], flen = 3 + (uint64_t)sh.sc[SC_SUMFIX];
    uint32_t pad = (8 - ((sh.sc[SC_CBITS] + 3) & 7)) & 7;
    uint64_t nlen = 3 + pad + 32 + 8ull * src_len;
    uint32_t btype = (nlen <= dlen && nlen <= flen) ? 0u : (flen <= dlen ? 1u : 2u);
    sh.sc[SC_BTYPE] = btype;
    // header items
    uint32_t k = 0;
    auto put = [&](uint32_t v, uint32_t n) { sh.hdr_val[k] = v; sh.hdr_nb[k] = (uint8_t)n; k++; };
    put((final ? 1u : 0u) | (btype << 1), 3);
    if (btype == 0) {
      if (pad) put(0, pad);
      put(src_len & 0xFFFFu, 16);
      put((~src_len) & 0xFFFFu, 16);
    } else if (btype == 2) {
      const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint32_t ccode[kNumClen];
      canonical_codes(sh.cl, kNumClen, ccode);
      put(sh.sc[SC_HLIT] - 257, 5);
      put(sh.sc[SC_HDIST] - 1, 5);
      put(sh.sc[SC_HCLEN] - 4, 4);
      for (uint32_t i = 0; i < sh.sc[SC_HCLEN]; i++) put(sh.cl[order[i]], 3);
      for (uint32_t i = 0; i < sh.sc[SC_NRSYM]; i++) {
        uint32_t s = sh.rsyms[i] & 0xFFu, ex = sh.rsyms[i] >> 8, c = ccode[s];
        uint32_t n = c >> 16, v = c & 0xFFFFu;
        if (s == 16) { v |= ex << n; n += 2; } else if (s == 17) { v |= ex << n; n += 3; } else if (s == 18) { v |= ex << n; n += 7; }
        put(v, n);
      }
    }
    sh.sc[SC_NHDR] = k;
  }
  __syncthreads();
  const uint32_t btype = sh.sc[SC_BTYPE];
  // -- code tables for the token pass
  if (btype == 1) {
    if (tid < 288) sh.lcode[tid] = fixed_lit_code((uint32_t)tid);
    else if (tid < 320) sh.dcode[tid - 288] = (__brev((uint32_t)(tid - 288)) >> 27) | (5u << 16);
  } else if (btype == 2) {
    // canonical code of a symbol = first code of its length + number of smaller symbols with that length
    const uint8_t *lens = tid < 288 ? sh.ll : sh.dl;
    int nsym = tid < 288 ? kNumLit : kNumDist, s = tid < 288 ? tid : tid - 288;
    if (tid < 320 && s < nsym) {
      uint32_t l = lens[s], code = 0;
      if (l) {
        uint32_t rank = 0, cnt[16];
#pragma unroll
        for (int i = 0; i < 16; i++) cnt[i] = 0;
        for (int t = 0; t < nsym; t++) { uint32_t lt = lens[t]; cnt[lt]++; rank += (lt == l && t < s); }
        cnt[0] = 0;
        uint32_t c = 0;
        for (uint32_t i = 1; i <= l; i++) c = (c + cnt[i - 1]) << 1;
        code = (__brev(c + rank) >> (32 - l)) | (l << 16);
      }
      (tid < 288 ? sh.lcode : sh.dcode)[s] = code;
    }
  }
  // the header items live in the parse area next to the staging window: pack them first
  {
    const uint32_t nh = sh.sc[SC_NHDR];
    const uint32_t *hv = sh.hdr_val;
    const uint8_t *hn = sh.hdr_nb;
    __syncthreads();
    pack_items(sh, out_words, out_cap_words, nh, [&](uint32_t i, uint64_t &b, uint32_t &n) { b = hv[i]; n = hn[i]; });
  }
  if (btype == 0) {
    const uint8_t *p = src + blk_src_start;
    pack_items(sh, out_words, out_cap_words, sh.sc[SC_BLK_SRCLEN], [&](uint32_t i, uint64_t &b, uint32_t &n) { b = p[i]; n = 8; });
  } else {
    const uint32_t *lc = sh.lcode, *dc = sh.dcode;
    pack_items(sh, out_words, out_cap_words, ntok + 1, [&](uint32_t i, uint64_t &b, uint32_t &n) {
      if (i == ntok) { uint32_t c = lc[256]; b = c & 0xFFFFu; n = c >> 16; return; }
      uint32_t t = toks[i], dist = t >> 9, len = t & 0x1FFu;
      if (!dist) { uint32_t c = lc[len]; b = c & 0xFFFFu; n = c >> 16; return; }
      uint32_t eb, ev, s = len_sym_of(len, eb, ev), c = lc[s];
      uint64_t acc = c & 0xFFFFu;
      uint32_t na = c >> 16;
      acc |= (uint64_t)ev << na; na += eb;
      s = dist_sym_of(dist, eb, ev); c = dc[s];
      acc |= (uint64_t)(c & 0xFFFFu) << na; na += c >> 16;
      acc |= (uint64_t)ev << na; na += eb;
      b = acc; n = na;
    });
  }
  // reset the per-block state
  for (int i = tid; i < 320; i += THREADS) sh.hist_l[i] = 0;  // (hist_d follows hist_l)
  if (tid == 0) sh.sc[SC_BLK_SRCLEN] = 0;
  __syncthreads();
}

// ---- one member ----------------------------------------------------------------------------------------------
__device__ void encode_member(const Shared &sh, const DeflateTask t, int level, uint32_t *toks, DeflateResult *res,
                              uint32_t *blk_lens) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t *src = t.src;
  const uint32_t n = (uint32_t)t.src_len;
  uint32_t *out_words = reinterpret_cast<uint32_t *>(t.dst);
  const uint64_t out_cap_words = t.dst_cap / 4;
  const LevelParams lp = level_params(level);
  const bool not_final = (t.flags & kDeflateNotFinal) != 0;
  const RingView ring{sh.ring};
  const PrevView prevv{sh.prev};

  for (int i = tid; i < (1 << kHashBits) / 2; i += THREADS) reinterpret_cast<uint32_t *>(sh.head)[i] = 0;
  for (int i = tid; i < kWindow / 2; i += THREADS) reinterpret_cast<uint32_t *>(sh.prev)[i] = 0;
  for (int i = tid; i < 320; i += THREADS) sh.hist_l[i] = 0;
  if (tid == 0) {
    sh.sc[SC_OUTW] = 0; sh.sc[SC_CARRY] = 0; sh.sc[SC_CBITS] = 0; sh.sc[SC_OVERFLOW] = 0;
    sh.sc[SC_CARRY_LEN] = 0; sh.sc[SC_CARRY_DIST] = 0; sh.sc[SC_BLK_SRCLEN] = 0;
  }
  __syncthreads();

  uint32_t pos = 0, kind = 0, loaded = 0, ntok = 0, nblocks = 0;
  uint64_t blk_src_start = 0;
  int tiles_in_block = 0;

  for (uint32_t ts = 0; ts < n; ts += kTile) {
    const uint32_t te = min(ts + (uint32_t)kTile, n), want = min(n, te + (uint32_t)kTile);
    // 1. stage input
    if ((((uintptr_t)src) & 15) == 0 && (loaded & 15) == 0) {  // 16-byte vectors (slots of the packed arena are aligned)
      const uint32_t nv = (want - loaded) >> 4;
      const uint4 *sv = reinterpret_cast<const uint4 *>(src + loaded);
      for (uint32_t i = tid; i < nv; i += THREADS)
        *reinterpret_cast<uint4 *>(sh.ringb + ((loaded + 16 * i) & (kRing - 1))) = __ldg(sv + i);
      for (uint32_t i = loaded + 16 * nv + tid; i < want; i += THREADS) sh.ringb[i & (kRing - 1)] = src[i];
    } else {
      for (uint32_t i = loaded + tid; i < want; i += THREADS) sh.ringb[i & (kRing - 1)] = src[i];
    }
    loaded = want;
    __syncthreads();
    // 2a. hashes (0xFFFF = position cannot start a match)
    for (int j = 0; j < PPT; j++) {
      uint32_t i = tid + THREADS * j, p = ts + i;
      sh.hsh[i] = (p < te && p + 4 <= n) ? (uint16_t)hash4(ring_load32(ring, p)) : (uint16_t)0xFFFF;
    }
    for (int i = tid; i < NWARPS * kClasses; i += THREADS) sh.cnt[i] = 0;
    if (tid == 0) sh.sc[SC_BATCH] = 0;
    __syncthreads();
    // 2b. partition the tile's positions by hash class (low bits), keeping position order:
    //     warp w ranks its PPW consecutive positions, 32 at a time
    uint32_t lrank[PPW / 32];
    for (int b = 0; b < PPW / 32; b++) {
      uint32_t i = warp * PPW + b * 32 + lane;
      uint32_t h = sh.hsh[i];
      bool valid = h != 0xFFFFu;
      const uint32_t cls = h & (kClasses - 1);
      uint32_t m = match_any_bits<kClassBits>(cls) & __ballot_sync(0xffffffffu, valid);  // valid lanes of my class
      uint32_t below = m & ((1u << lane) - 1u);
      uint32_t base = valid ? sh.cnt[warp * kClasses + cls] : 0;
      lrank[b] = base + __popc(below);
      __syncwarp();
      if (valid && below == 0) sh.cnt[warp * kClasses + cls] = (uint16_t)(base + __popc(m));
      __syncwarp();
    }
    __syncthreads();
    if (tid < kClasses) {  // per class: exclusive offsets over the warps, then the class total
      uint32_t run = 0;
      for (int w = 0; w < NWARPS; w++) { uint32_t c = sh.cnt[w * kClasses + tid]; sh.cnt[w * kClasses + tid] = (uint16_t)run; run += c; }
      sh.cstart[tid + 1] = (uint16_t)run;
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t run = 0;
      for (int c = 0; c < kClasses; c++) { uint32_t v = sh.cstart[c + 1]; sh.cstart[c] = (uint16_t)run; run += v; }
      sh.cstart[kClasses] = (uint16_t)run;
    }
    __syncthreads();
    for (int b = 0; b < PPW / 32; b++) {
      uint32_t i = warp * PPW + b * 32 + lane;
      uint32_t h = sh.hsh[i];
      const uint32_t cls = h & (kClasses - 1);
      if (h != 0xFFFFu) sh.poslist[sh.cstart[cls] + sh.cnt[warp * kClasses + cls] + lrank[b]] = (uint16_t)i;
    }
    __syncthreads();
    // 2c. chain insertion: warp w owns hash class w, walks its positions in order, 32 per step
    {
      const uint32_t c0 = sh.cstart[warp], c1 = sh.cstart[warp + 1];
      for (uint32_t base = c0; base < c1; base += 32) {
        uint32_t k = base + lane;
        bool valid = k < c1;
        uint32_t i = valid ? sh.poslist[k] : 0, p = ts + i;
        uint32_t h = valid ? sh.hsh[i] : 0u;
        uint32_t m = match_any_bits<kHashBits - kClassBits>(h >> kClassBits) & __ballot_sync(0xffffffffu, valid);  // valid lanes with my hash (the class bits are equal anyway)
        uint32_t below = m & ((1u << lane) - 1u);
        int srcl = below ? 31 - __clz(below) : 0;
        uint32_t pp = __shfl_sync(0xffffffffu, p, srcl);
        uint16_t cand = 0;
        if (valid) cand = below ? (uint16_t)pp : sh.head[h];
        __syncwarp();  // every head is read before any head of this batch is replaced
        if (valid) {
          sh.first[i] = cand;
          sh.prev[p & (kWindow - 1)] = cand;
          if ((m >> lane) == 1u) sh.head[h] = (uint16_t)p;  // highest lane of the group
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // 3. longest match at every position (slot 0 carries position ts-1 from the previous tile)
    if (tid == 0) { sh.mlen[0] = (uint16_t)sh.sc[SC_CARRY_LEN]; sh.mdist[0] = (uint16_t)sh.sc[SC_CARRY_DIST]; }
    // batches of 32 consecutive positions are handed out through a counter: chains are much longer in some
    // stretches of the input than in others, and with two fixed batches per warp the CTA waited for its slowest warp
    for (;;) {
      // begin / store are uniform over the warp; only the chain steps diverge (lanes with short chains idle)
      uint32_t b = 0;
      if (lane == 0) b = atomicAdd(&sh.sc[SC_BATCH], 1u);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b >= kTile / 32) break;
      uint32_t i = b * 32 + lane, p = ts + i, d = 0, l = 0;
      if (p < te && p + 4 <= n) l = find_match(ring, prevv, p, n, sh.first[i], lp.depth, lp.nice, d);
      sh.mlen[1 + i] = (uint16_t)l;
      sh.mdist[1 + i] = (uint16_t)d;
    }
    __syncthreads();
    // 4. lazy parse by pointer jumping over nodes v = 2 * (p - ts) + kind; codes >= 2T are exits
    if (pos < te) {
      // both nodes of a position are handled by one thread: their jump codes share a 32-bit word and their
      // marks a 16-bit word, which saves a third of the shared-memory operations of every round
      uint32_t *jA32 = reinterpret_cast<uint32_t *>(sh.jumpA);
      uint16_t *mk16 = reinterpret_cast<uint16_t *>(sh.mark);
      for (int j = 0; j < PPT; j++) {
        const uint32_t i = tid + THREADS * j, p = ts + i;
        const uint32_t ml = sh.mlen[1 + i], mp = sh.mlen[i];
        uint32_t code[2];
#pragma unroll
        for (uint32_t k = 0; k < 2; k++) {
          uint32_t nk, em;
          const uint32_t np = lazy_next(p, k, ml, mp, nk, em);
          code[k] = np < te ? 2 * (np - ts) + nk : 2 * kTile + 2 * (np - te) + nk;
        }
        jA32[i] = code[0] | (code[1] << 16);
        mk16[i] = 0;
      }
      __syncthreads();
      const uint32_t entry = 2 * (pos - ts) + kind;
      if (tid == 0) sh.mark[entry] = 1;
      __syncthreads();
      uint16_t *ja = sh.jumpA, *jb = sh.jumpB;
      // Marks only ever move along the path (a marked node marks J_r of itself), so a mark that becomes visible
      // within the round it is written (there is no barrier between the reads and the writes of sh.mark) can
      // only mark further path nodes early: the race is benign and the final set is exactly the path.
      for (int r = 0; r < 11; r++) {
        const uint32_t *ja32 = reinterpret_cast<const uint32_t *>(ja);
        uint32_t *jb32 = reinterpret_cast<uint32_t *>(jb);
        for (int j = 0; j < PPT; j++) {
          const uint32_t i = tid + THREADS * j;
          const uint32_t jj = ja32[i], mm = mk16[i];
          const uint32_t w0 = jj & 0xFFFFu, w1 = jj >> 16;
          if ((mm & 0xFFu) && w0 < 2 * kTile) sh.mark[w0] = 1;
          if ((mm >> 8) && w1 < 2 * kTile) sh.mark[w1] = 1;
          const uint32_t n0 = w0 < 2 * kTile ? ja[w0] : w0, n1 = w1 < 2 * kTile ? ja[w1] : w1;
          jb32[i] = n0 | (n1 << 16);
        }
        __syncthreads();
        uint16_t *tmp = ja; ja = jb; jb = tmp;
        if (ja[entry] >= 2 * kTile) break;  // the entry has left the tile: every node of the path is marked
      }
      const uint32_t ex = ja[entry] - 2 * kTile;  // after 2^11 >= T steps the entry has left the tile
      pos = te + (ex >> 1);
      kind = ex & 1u;
      // 5. tokens of the visited nodes, in position order (2 * PPT consecutive nodes per thread)
      uint32_t tk[2 * PPT], cnt = 0, srcsum = 0;
      for (int j = 0; j < 2 * PPT; j++) {
        uint32_t v = tid * (2 * PPT) + j, i = v >> 1, k = v & 1u, p = ts + i;
        if (!sh.mark[v] || p >= te) continue;
        uint32_t nk, em;
        uint32_t mp = sh.mlen[i];
        lazy_next(p, k, sh.mlen[1 + i], mp, nk, em);
        if (em == 1 || em == 2) {
          uint32_t b = sh.ringb[(em == 1 ? p : p - 1) & (kRing - 1)];
          tk[cnt++] = tok_lit(b);
          atomicAdd(&sh.hist_l[b], 1u);
          srcsum += 1;
        } else if (em == 3) {
          uint32_t d = sh.mdist[i], eb, ev;
          tk[cnt++] = tok_match(mp, d);
          atomicAdd(&sh.hist_l[len_sym_of(mp, eb, ev)], 1u);
          atomicAdd(&sh.hist_d[dist_sym_of(d, eb, ev)], 1u);
          srcsum += mp;
        }
      }
      uint32_t total;
      uint32_t off = block_scan(cnt, sh.scan, &total);
      for (uint32_t j = 0; j < cnt; j++) toks[ntok + off + j] = tk[j];
      ntok += total;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) srcsum += __shfl_xor_sync(0xffffffffu, srcsum, o);
      if (lane == 0 && srcsum) atomicAdd(&sh.sc[SC_BLK_SRCLEN], srcsum);
    }
    __syncthreads();
    if (tid == 0) { sh.sc[SC_CARRY_LEN] = sh.mlen[te - ts]; sh.sc[SC_CARRY_DIST] = sh.mdist[te - ts]; }
    tiles_in_block++;
    const bool last = te == n;
    if (tiles_in_block == kTilesPerBlock || last) {
      __syncthreads();
      const uint32_t blen = sh.sc[SC_BLK_SRCLEN];
      finalize_block(sh, src, blk_src_start, ntok, toks, last && !not_final, out_words, out_cap_words);
      if (tid == 0 && blk_lens) blk_lens[t.blk_off + nblocks] = blen;
      blk_src_start += blen;
      ntok = 0; tiles_in_block = 0; nblocks++;
    }
  }
  if (n == 0) {
    finalize_block(sh, src, 0, 0, toks, !not_final, out_words, out_cap_words);
    if (tid == 0 && blk_lens) blk_lens[t.blk_off] = 0;
    nblocks++;
  }
  if (not_final) {
    // segment of a larger stream: close with an empty stored block (000, pad to a byte, LEN 0, NLEN ffff) so
    // that the next segment starts byte aligned and the concatenation is one valid RFC 1951 stream
    __syncthreads();
    if (tid == 0) {
      uint32_t pad = (8 - ((sh.sc[SC_CBITS] + 3) & 7)) & 7, k = 0;
      sh.hdr_val[k] = 0; sh.hdr_nb[k++] = 3;
      if (pad) { sh.hdr_val[k] = 0; sh.hdr_nb[k++] = (uint8_t)pad; }
      sh.hdr_val[k] = 0; sh.hdr_nb[k++] = 16;
      sh.hdr_val[k] = 0xFFFFu; sh.hdr_nb[k++] = 16;
      sh.sc[SC_NHDR] = k;
    }
    __syncthreads();
    const uint32_t nh = sh.sc[SC_NHDR];
    const uint32_t *hv = sh.hdr_val;
    const uint8_t *hn = sh.hdr_nb;
    pack_items(sh, out_words, out_cap_words, nh, [&](uint32_t i, uint64_t &b, uint32_t &nn) { b = hv[i]; nn = hn[i]; });
  }
  __syncthreads();
  if (tid == 0) {
    uint64_t outw = sh.sc[SC_OUTW];
    uint32_t cbits = sh.sc[SC_CBITS], tail = (cbits + 7) >> 3;
    bool ovf = sh.sc[SC_OVERFLOW] != 0 || outw * 4 + tail > t.dst_cap;
    if (!ovf) {
      uint32_t carry = sh.sc[SC_CARRY];
      for (uint32_t b = 0; b < tail; b++) t.dst[outw * 4 + b] = (uint8_t)(carry >> (8 * b));
    }
    res->out_len = ovf ? 0 : outw * 4 + tail;
    res->status = ovf ? ZIPC_ERR_DST_TOO_SMALL : ZIPC_OK;
    res->blocks = nblocks;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1)
deflate_kernel(const DeflateTask *__restrict__ tasks, uint32_t ntasks, DeflateResult *__restrict__ results,
               unsigned int *__restrict__ queue, uint32_t *__restrict__ tok_scratch, int level,
               uint32_t *__restrict__ blk_lens) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const Shared sh = carve(smem_raw);
  uint32_t *toks = tok_scratch + (size_t)blockIdx.x * kTokCap;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh.sc[SC_TASK] = atomicAdd(queue, 1u);
    __syncthreads();
    const uint32_t task = sh.sc[SC_TASK];
    if (task >= ntasks) break;
    encode_member(sh, tasks[task], level, toks, &results[task], blk_lens);
  }
}

// level `None (reference :1106-1116): stored blocks only, one CTA per member.  Blocks hold 65534 bytes as in the
// reference (:747-750), so the output is byte-identical to Zipc_deflate.deflate ~level:`None.
__global__ void __launch_bounds__(256)
stored_kernel(const DeflateTask *__restrict__ tasks, uint32_t ntasks, DeflateResult *__restrict__ results) {
  const uint32_t task = blockIdx.x;
  if (task >= ntasks) return;
  const DeflateTask t = tasks[task];
  constexpr uint64_t B = kStoredBlock;
  const uint64_t n = t.src_len, nblk = n ? (n + B - 1) / B : 1, need = n + 5 * nblk;
  if (need > t.dst_cap) {
    if (threadIdx.x == 0) { results[task].out_len = 0; results[task].status = ZIPC_ERR_DST_TOO_SMALL; results[task].blocks = 0; }
    return;
  }
  for (uint64_t b = threadIdx.x; b < nblk; b += blockDim.x) {
    uint64_t len = b + 1 < nblk ? B : n - b * B;
    uint8_t *h = t.dst + b * (B + 5);
    h[0] = (b + 1 == nblk && !(t.flags & kDeflateNotFinal)) ? 1 : 0;
    h[1] = (uint8_t)len; h[2] = (uint8_t)(len >> 8); h[3] = (uint8_t)~len; h[4] = (uint8_t)(~len >> 8);
  }
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) t.dst[i + 5 * (i / B + 1)] = t.src[i];
  if (threadIdx.x == 0) { results[task].out_len = need; results[task].status = ZIPC_OK; results[task].blocks = (uint32_t)nblk; }
}

unsigned long long g_attr_devs = 0;  // bit d: attributes set on device d (function attributes are per device)

}  // namespace

int deflate_launch(zipc_b200_ctx *ctx, const DeflateTask *d_tasks, uint32_t n, DeflateResult *d_results, int level,
                   uint32_t *d_blk_lens) {
  if (n == 0) return ZIPC_OK;
  if (level == ZIPC_LEVEL_NONE) {
    stored_kernel<<<n, 256, 0, ctx->stream>>>(d_tasks, n, d_results);
    ctx->launches++;
    ZB_CUDA(ctx, cudaGetLastError());
    return ZIPC_OK;
  }
  if (!(g_attr_devs >> (ctx->device & 63) & 1ull)) {
    ZB_CUDA(ctx, cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    g_attr_devs |= 1ull << (ctx->device & 63);
  }
  uint32_t grid = (uint32_t)ctx->sm_count;
  if (grid > n) grid = n;
  size_t tok_bytes = (size_t)grid * kTokCap * sizeof(uint32_t);
  if (int st = ctx->d_scratch.reserve(tok_bytes + 256)) return st;
  unsigned int *queue = reinterpret_cast<unsigned int *>(ctx->d_scratch.as<uint8_t>() + tok_bytes);
  ZB_CUDA(ctx, cudaMemsetAsync(queue, 0, sizeof(unsigned int), ctx->stream));
  {
    KernelTimer kt(ctx);
    deflate_kernel<<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint32_t>(), level, d_blk_lens);
  }
  ctx->launches++;
  ZB_CUDA(ctx, cudaGetLastError());
  return ZIPC_OK;
}

}  // namespace zb

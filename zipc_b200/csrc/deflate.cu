// deflate.cu -- batched DEFLATE encoder for sm_100a: one CTA per ZIP member (or independent segment).
//
// Replaces the encode side of the reference (src/zipc_deflate.ml:742-1277): Lz77.compress and its
// hash-chain matcher (:1140-1245), the block writer and block-type choice (:873-1104) and the deflate /
// crc_32_and_deflate / zlib_compress entry points (:1247-1277).  The output is valid RFC 1951, inflates to
// the input bit-exactly through the reference's inflate, and lands within 2 % of the reference's compressed
// size per level (fast 0.994 x, default 1.018 x, best 1.001 x on the C4 members).
//
// A member is processed tile by tile (2048 input bytes), everything in shared memory (64 KiB input ring,
// 32 KiB hash heads, 64 KiB chain links).  The CTA's 32 warps are split into two groups that work on
// neighbouring tiles at the same time, synchronised with named barriers, so that the latency-bound chain
// walks of one tile are covered by the throughput-bound preparation of the next:
//
//   front end, tile t+1:  wait for the tile's bytes (the copy engine was asked for them one tile earlier:
//       cp.async.bulk into the ring, completion on an mbarrier); hash every position; partition the positions by hash
//       class so that the front warps insert them into the chains concurrently yet in exact position order (equal
//       hashes inside a 32-lane batch are resolved with ballots), which reproduces the serial chain semantics;
//       then a chain walk at every position -- the level's whole budget for fast / default (4 / 12 candidates),
//       a shallow one (2 candidates) for best.
//   back end, tile t:  warp 0 runs the one-step lazy parse (reference :1224-1241) over the lengths --
//       each lane walks a 64-position sub-range from a guessed entry, then the lanes' paths are stitched
//       together (a path that enters a sub-range elsewhere merges with the guessed one after a few tokens).
//       Level best only: the positions a parse over the shallow lengths visits (about 30 %) continue their chain
//       walk to the full depth (1024), in waves that also cover the positions a match taken there would land
//       on; then warp 0 parses again over the improved lengths: that is the final parse.
//   all warps:  visited nodes -> tokens (CTA prefix sum), literal/length and distance histograms (shared atomics).
//
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

// Optional per-phase cycle accounting (tuning builds only: make EXTRA=-DZB_DEFLATE_TIMING): one thread per role adds the
// clock64 deltas between phase boundaries to global counters, read back with zipc_b200_debug_deflate_phases().
#ifdef ZB_DEFLATE_TIMING
}  // namespace
__device__ unsigned long long g_phase[16];
__device__ unsigned long long g_dbg[16];
namespace {
#define DBG(k, v) atomicAdd(&g_dbg[k], (unsigned long long)(v))
// loop guard of the tuning build: a loop that spins past `limit` records where and traps instead of hanging the GPU
#define GUARD(var, limit, where) do { if (++(var) > (limit)) { g_dbg[15] = (where); __threadfence(); __trap(); } } while (0)
#define GUARD_DECL(var) unsigned int var = 0;
#define PH_DECL long long ph_t = clock64();
#define PH_AT(k, who) do { if (threadIdx.x == (who)) { long long ph_n = clock64(); atomicAdd(&g_phase[k], (unsigned long long)(ph_n - ph_t)); ph_t = ph_n; } } while (0)
#define PH_RESET ph_t = clock64();
#else
#define DBG(k, v) do { } while (0)
#define GUARD(var, limit, where) do { } while (0)
#define GUARD_DECL(var)
#define PH_DECL
#define PH_AT(k, who) do { } while (0)
#define PH_RESET
#endif
#define PH(k) PH_AT(k, 0)
#ifndef ZB_BACK_HELPS
#define ZB_BACK_HELPS 1
#endif
#define ZB_SMEM extern __shared__ __align__(16) uint8_t smem_raw[]
enum { PH_FRONT_WAIT = 0, PH_BACK_WAIT, PH_PARSE0, PH_DEEP, PH_PARSE1, PH_TOKENS, PH_FIN_SORT, PH_FIN_HUFF, PH_FIN_HDR, PH_FIN_PACK, PH_OTHER,
       PH_F_STAGE, PH_F_PART, PH_F_INSERT, PH_F_SHALLOW, PH_JUMP };

constexpr int THREADS = 1024;
constexpr int NWARPS = THREADS / 32;
// The split between the two groups is a template parameter BW (back-end warps):
//   BW = 1   levels fast and default: every position gets the level's whole chain budget in the front end (31 warps), there are
//            no deep walks, and the back end is warp 0 running the final parse of the previous tile underneath;
//   BW = 16  level best: shallow walks everywhere, deep walks (15 + 1 warps) where a parse goes.
constexpr int kChunk32 = kTile / 32;               // 64 batches of 32 consecutive positions, ranked by the front warps
constexpr int kMaxClasses = 31;                    // hash classes of the chain insertion: one front warp each
constexpr int PPT = kTile / THREADS;               // positions per thread per tile
constexpr int SORTN = 512;            // keys of the literal/length frequency sort
constexpr int kTokCap = 65536;        // tokens per block scratch (block source <= 61440 + 258)
constexpr int kChunk = 2048;          // items per bit-packer chunk
constexpr int IPT = kChunk / THREADS; // items per thread per chunk
constexpr int kStageWords = kChunk * 48 / 32 + 8;
constexpr uint32_t kExit = 2 * kTile; // jump codes >= kExit leave the tile

// named barriers (0 is __syncthreads)
template <int BW> __device__ __forceinline__ void bar_front() { asm volatile("bar.sync 1, %0;" ::"n"((NWARPS - BW) * 32) : "memory"); }
template <int BW> __device__ __forceinline__ void bar_back() { asm volatile("bar.sync 2, %0;" ::"n"(BW * 32) : "memory"); }
// barrier 3 (BW = 1 only): the front threads arrive, without waiting, when the chains of their tile are complete; the parse
// warp waits there before it takes shallow-walk batches of that tile
__device__ __forceinline__ void bar_walk_open_arrive() { asm volatile("bar.arrive 3, %0;" ::"n"(NWARPS * 32) : "memory"); }
__device__ __forceinline__ void bar_walk_open_wait() { asm volatile("bar.sync 3, %0;" ::"n"(NWARPS * 32) : "memory"); }

// ---- input staging by the copy engine: cp.async.bulk (global -> shared, 1-D) completing on an mbarrier ---------------------
// One elected thread asks for the next tile's bytes a whole tile ahead; nobody spends instructions on staging, the front
// group only waits on the barrier's phase when it gets to the tile (by then the bytes have long arrived).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(arrivals), "r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  GUARD_DECL(g_w)
  while (!mbar_try_wait(bar, parity)) { GUARD(g_w, 20000u, 600 + parity); }
}

// ---- shared memory map ------------------------------------------------------------------------------------
constexpr int GEN_TOK = 0;                                   // u16[kTile]: first candidate, then the resume token
constexpr int GEN_MLEN = GEN_TOK + kTile * 2;                // u16[kTile + 8]: slot 0 = the position before the tile
constexpr int GEN_MDIST = GEN_MLEN + (kTile + 8) * 2;        // u16[kTile + 8]
constexpr int GEN_BYTES = GEN_MDIST + (kTile + 8) * 2;       // 12320
constexpr int OFF_RING = 0;                                  // u32[16384]
constexpr int OFF_PREV = OFF_RING + kRing;                   // u16[32768]
constexpr int OFF_HEAD = OFF_PREV + kWindow * 2;             // u16[1 << kHashBits]
constexpr int OFF_GEN = OFF_HEAD + (2 << kHashBits);         // two generations of tile arrays (tile t, tile t + 1)
constexpr int OFF_Z = OFF_GEN + 2 * GEN_BYTES;               // jump codes, hashes, position list | bit staging + header items
constexpr int Z_JMP = 0, Z_HSH = Z_JMP + kTile * 4, Z_POSL = Z_HSH + kTile * 2, Z_BYTES = Z_POSL + kTile * 2;
constexpr int OFF_VQ = OFF_Z + Z_BYTES;                      // u16[2][kTile]: positions a parse is guessed to visit (tile t, tile t + 1)
constexpr int OFF_OQ = OFF_VQ + 2 * kTile * 2;               // u16[2 * kTile]: landings queued by the deep walks (append only) | finalize scratch
constexpr int OFF_MARK = OFF_OQ + kTile * 4;                 // u32[64] kind 0 marks, u32[64] kind 1 marks, u32[64] claims
constexpr int OFF_FIN = OFF_OQ;                              // block-finalize scratch: the landing queue is dead by then
constexpr int FIN_BYTES = 6144;
constexpr int OFF_HIST = OFF_MARK + 3 * 256;                 // u32[288] lit, u32[32] dist
constexpr int OFF_CODE = OFF_HIST + 320 * 4;                 // u32[288] lit codes, u32[32] dist codes
constexpr int OFF_MISC = OFF_CODE + 320 * 4;                 // class counters, scan scratch, scalars
constexpr int MISC_BYTES = 5120;
constexpr int kSmemBytes = OFF_MISC + MISC_BYTES;
static_assert(kStageWords * 4 + 704 * 5 + 64 <= Z_BYTES, "bit staging + header items must fit the tile scratch");
static_assert(512 * 4 * 2 + 320 * 2 + 288 + 32 + 320 + 32 + 5 * 128 <= FIN_BYTES && FIN_BYTES <= kTile * 4, "finalize scratch");
static_assert(kSmemBytes <= 232448, "227 KiB of shared memory per CTA");

struct Gen {
  uint16_t *tok, *mlen, *mdist;
};

struct Shared {
  uint32_t *ring;
  uint8_t *ringb;
  uint16_t *prev, *head;
  uint32_t *jmp;
  uint16_t *hsh, *posl, *oq, *vq;
  uint32_t *mk0, *mk1, *claim;
  // emit view of Z
  uint32_t *stage, *hdr_val;
  uint8_t *hdr_nb;
  // finalize scratch
  uint32_t *keys, *hscratch;
  uint16_t *rsyms;
  uint8_t *ll, *dl, *both, *cl;
  uint32_t *dkeys, *dscratch, *ckeys, *cscratch, *cfreq;
  uint32_t *hist_l, *hist_d, *lcode, *dcode;
  uint16_t *cnt;     // [kChunk32][classes] per-batch class counts, then offsets
  uint16_t *cstart;  // [classes + 1] class starts
  uint32_t *scan;    // [NWARPS + 1]
  uint32_t *sc;      // scalars
  uint32_t *pslot;   // [32] parse: entry hand-over between lanes
  uint64_t *ldbar;   // mbarrier of the input staging
};

// scalar slots in sh.sc
enum { SC_TASK = 0, SC_OUTW, SC_CARRY, SC_CBITS, SC_OVERFLOW, SC_M_L, SC_M_D, SC_NHDR, SC_BTYPE, SC_HLIT, SC_HDIST,
       SC_SUMDYN, SC_SUMFIX, SC_BLK_SRCLEN, SC_NRSYM, SC_HCLEN, SC_HDRBITS, SC_BATCH, SC_QHEAD, SC_OQN, SC_EXIT, SC_VQN /* [2] */, SC_VQN1, SC_LD_PHASES /* bulk copies issued so far */ };

__device__ __forceinline__ Gen gen_of(uint8_t *base, int g) {  // plain arithmetic on the shared base: the address space stays known
  uint8_t *b = base + OFF_GEN + g * GEN_BYTES;
  Gen r;
  r.tok = reinterpret_cast<uint16_t *>(b + GEN_TOK);
  r.mlen = reinterpret_cast<uint16_t *>(b + GEN_MLEN);
  r.mdist = reinterpret_cast<uint16_t *>(b + GEN_MDIST);
  return r;
}

__device__ __forceinline__ Shared carve(uint8_t *base) {
  Shared s;
  s.ring = reinterpret_cast<uint32_t *>(base + OFF_RING);
  s.ringb = base + OFF_RING;
  s.prev = reinterpret_cast<uint16_t *>(base + OFF_PREV);
  s.head = reinterpret_cast<uint16_t *>(base + OFF_HEAD);
  uint8_t *z = base + OFF_Z;
  s.jmp = reinterpret_cast<uint32_t *>(z + Z_JMP);
  s.hsh = reinterpret_cast<uint16_t *>(z + Z_HSH);
  s.posl = reinterpret_cast<uint16_t *>(z + Z_POSL);
  s.oq = reinterpret_cast<uint16_t *>(base + OFF_OQ);
  s.vq = reinterpret_cast<uint16_t *>(base + OFF_VQ);
  s.mk0 = reinterpret_cast<uint32_t *>(base + OFF_MARK);
  s.mk1 = s.mk0 + 64;
  s.claim = s.mk1 + 64;
  s.stage = reinterpret_cast<uint32_t *>(z);
  s.hdr_val = s.stage + kStageWords;
  s.hdr_nb = reinterpret_cast<uint8_t *>(s.hdr_val + 704);
  uint8_t *y = base + OFF_FIN;
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
  s.cnt = reinterpret_cast<uint16_t *>(m);                       // 64 batches * kMaxClasses * 2 bytes <= 4096
  s.cstart = reinterpret_cast<uint16_t *>(m + 4096);             // kMaxClasses + 1 entries
  s.scan = reinterpret_cast<uint32_t *>(m + 4096 + 128);         // 40 words
  s.sc = s.scan + 40;                                            // 32 words
  s.pslot = s.sc + 32;                                           // 32 words
  s.ldbar = reinterpret_cast<uint64_t *>(s.pslot + 32);          // 8 bytes, 8-byte aligned
  return s;
}
static_assert(kChunk32 * kMaxClasses * 2 <= 4096 && 4096 + 128 + (40 + 32 + 32) * 4 + 8 <= MISC_BYTES, "misc area");

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
__device__ __noinline__ void finalize_block(const uint8_t *src, uint64_t blk_src_start, uint32_t ntok,
                                            const uint32_t *toks, bool final, uint32_t *out_words, uint64_t out_cap_words) {
  ZB_SMEM;
  const Shared sh = carve(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PH_DECL
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
  PH(PH_FIN_SORT);
  // -- code lengths (serial per alphabet, two alphabets side by side)
  if (tid == 0) huff_lengths_from_sorted(sh.keys, (int)sh.sc[SC_M_L], kNumLit, 15, sh.ll, sh.hscratch);
  if (tid == 32) huff_lengths_from_sorted(sh.dkeys, (int)sh.sc[SC_M_D], kNumDist, 15, sh.dl, sh.dscratch);
  __syncthreads();
  PH(PH_FIN_HUFF);
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
  // -- code length alphabet, header cost, block type (reference :959-1043, :1071-1104).  The run-length coding of the
  //    <= 316 code lengths is done in parallel (one thread per run; it must produce exactly rle_code_lengths()'s symbols,
  //    which is what the host model emits); only the 19-symbol Huffman code over them is built by one thread.
  if (warp == 0) {  // HLIT: one past the last used literal/length symbol, at least 257
    uint32_t hi = 0;
    for (int i = lane; i < kNumLit; i += 32) if (sh.ll[i]) hi = (uint32_t)i + 1;
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) sh.sc[SC_HLIT] = hi < 257u ? 257u : hi;
  } else if (warp == 1) {  // HDIST, at least 1
    uint32_t hi = (lane < kNumDist && sh.dl[lane]) ? (uint32_t)lane + 1 : 0u;
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) sh.sc[SC_HDIST] = hi < 1u ? 1u : hi;
  } else if (warp == 2 && lane < kNumClen) sh.cfreq[lane] = 0;
  __syncthreads();
  const uint32_t hlit = sh.sc[SC_HLIT], hdist = sh.sc[SC_HDIST], nlen_codes = hlit + hdist;
  if ((uint32_t)tid < hlit) sh.both[tid] = sh.ll[tid];
  else if ((uint32_t)tid < nlen_codes) sh.both[tid] = sh.dl[tid - hlit];
  __syncthreads();
  {
    // run heads: thread i heads a run if both[i] differs from both[i - 1]; a head finds the next head in the ballot words
    const bool inside = (uint32_t)tid < nlen_codes;
    const uint32_t v = inside ? sh.both[tid] : 0xFFu;
    const bool head = inside && (tid == 0 || sh.both[tid - 1] != v);
    const uint32_t hw = __ballot_sync(0xffffffffu, head);
    if (lane == 0 && warp < 10) sh.dscratch[warp] = hw;
    __syncthreads();
    uint32_t run = 0, cnt = 0;
    if (head) {
      uint32_t w = tid >> 5, m = (tid & 31) == 31 ? 0u : (sh.dscratch[w] & (0xFFFFFFFEu << (tid & 31)));
      uint32_t nxt = nlen_codes;
      for (;;) {
        if (m) { nxt = w * 32 + (uint32_t)__ffs((int)m) - 1; break; }
        if (++w >= 10) break;
        m = sh.dscratch[w];
      }
      if (nxt > nlen_codes) nxt = nlen_codes;
      run = nxt - (uint32_t)tid;
      // symbols this run turns into (rle_code_lengths)
      if (v == 0) { uint32_t r = run; while (r >= 11) { r -= r > 138 ? 138 : r; cnt++; } if (r >= 3) { cnt++; r = 0; } cnt += r; }
      else { uint32_t r = run - 1; cnt = 1; while (r >= 3) { r -= r > 6 ? 6 : r; cnt++; } cnt += r; }
    }
    uint32_t total;
    uint32_t k = block_scan(cnt, sh.scan, &total);
    if (head) {
      if (v == 0) {
        uint32_t r = run;
        while (r >= 11) { const uint32_t t = r > 138 ? 138 : r; sh.rsyms[k++] = (uint16_t)(18 | ((t - 11) << 8)); atomicAdd(&sh.cfreq[18], 1u); r -= t; }
        if (r >= 3) { sh.rsyms[k++] = (uint16_t)(17 | ((r - 3) << 8)); atomicAdd(&sh.cfreq[17], 1u); r = 0; }
        if (r) atomicAdd(&sh.cfreq[0], r);
        while (r-- > 0) sh.rsyms[k++] = 0;
      } else {
        sh.rsyms[k++] = (uint16_t)v;
        uint32_t r = run - 1, plain = 1;
        while (r >= 3) { const uint32_t t = r > 6 ? 6 : r; sh.rsyms[k++] = (uint16_t)(16 | ((t - 3) << 8)); atomicAdd(&sh.cfreq[16], 1u); r -= t; }
        plain += r;
        atomicAdd(&sh.cfreq[v], plain);
        while (r-- > 0) sh.rsyms[k++] = (uint16_t)v;
      }
    }
    if (tid == 0) sh.sc[SC_NRSYM] = total;
  }
  __syncthreads();
  if (tid == 0) {
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
    sh.sc[SC_HCLEN] = (uint32_t)hclen; sh.sc[SC_HDRBITS] = hdr;
    // block type (exact costs) and the fixed part of the header
    const uint32_t src_len = sh.sc[SC_BLK_SRCLEN];
    uint64_t dlen = (uint64_t)hdr + sh.sc[SC_SUMDYN], flen = 3 + (uint64_t)sh.sc[SC_SUMFIX];
    uint32_t pad = (8 - ((sh.sc[SC_CBITS] + 3) & 7)) & 7;
    uint64_t nlen = 3 + pad + 32 + 8ull * src_len;
    uint32_t btype = (nlen <= dlen && nlen <= flen) ? 0u : (flen <= dlen ? 1u : 2u);
    sh.sc[SC_BTYPE] = btype;
    uint32_t k = 0;
    auto put = [&](uint32_t v, uint32_t n) { sh.hdr_val[k] = v; sh.hdr_nb[k] = (uint8_t)n; k++; };
    put((final ? 1u : 0u) | (btype << 1), 3);
    if (btype == 0) {
      if (pad) put(0, pad);
      put(src_len & 0xFFFFu, 16);
      put((~src_len) & 0xFFFFu, 16);
    } else if (btype == 2) {
      uint32_t ccode[kNumClen];
      canonical_codes(sh.cl, kNumClen, ccode);
      for (int s = 0; s < kNumClen; s++) sh.ckeys[s] = ccode[s];  // (the sort keys are dead) for the run-length symbols below
      put(hlit - 257, 5);
      put(hdist - 1, 5);
      put((uint32_t)hclen - 4, 4);
      for (int i = 0; i < hclen; i++) put(sh.cl[order[i]], 3);
    }
    sh.sc[SC_NHDR] = k;
  }
  __syncthreads();
  if (sh.sc[SC_BTYPE] == 2) {  // the run-length symbols with their extra bits, one thread each
    const uint32_t base = sh.sc[SC_NHDR], nr = sh.sc[SC_NRSYM];
    if ((uint32_t)tid < nr) {
      const uint32_t sym = sh.rsyms[tid] & 0xFFu, ex = sh.rsyms[tid] >> 8, c = sh.ckeys[sym];
      uint32_t nb = c >> 16, val = c & 0xFFFFu;
      if (sym == 16) { val |= ex << nb; nb += 2; } else if (sym == 17) { val |= ex << nb; nb += 3; } else if (sym == 18) { val |= ex << nb; nb += 7; }
      sh.hdr_val[base + tid] = val; sh.hdr_nb[base + tid] = (uint8_t)nb;
    }
    __syncthreads();
    if (tid == 0) sh.sc[SC_NHDR] = base + nr;
    __syncthreads();
  }
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
  PH(PH_FIN_HDR);
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
  PH(PH_FIN_PACK);
  // reset the per-block state
  for (int i = tid; i < 320; i += THREADS) sh.hist_l[i] = 0;  // (hist_d follows hist_l)
  if (tid == 0) sh.sc[SC_BLK_SRCLEN] = 0;
  __syncthreads();
}

// ---- the chain walk of one position, taken apart so that a warp stays together where it is expensive ---------------------
// This is match_begin / match_resume + match_step (+ match_trim) of deflate_core.h in the 16-bit position arithmetic of the
// chain tables, and it finds exactly what they find.  The cheap part of a step runs for several candidates in a row, each
// lane on its own (walk_light): is the candidate in range, and do the FOUR bytes ending at the byte that would have to
// improve the match agree (match_step tests one byte; a candidate can only beat `best` if all bytes up to there match, so
// testing more of them filters more and changes nothing).  Lanes whose candidate passes wait, and all of them do the full
// comparison together (walk_compare).
struct Walk {
  uint32_t p, p16, c, last, rng, tdist, best, best_dist, chk4, max_len, pw0, pw1, link;
  int steps;
  bool more;  // stopped only because the step budget ran out
};
__device__ __forceinline__ void walk_from(Walk &w, const MatchState &m, const RingView &ring, uint32_t trim) {
  w.p = m.p; w.p16 = m.p & 0xFFFFu; w.c = m.c; w.last = m.last_dist; w.rng = m.reach; w.best = m.best; w.best_dist = m.best_dist;
  w.max_len = m.max_len; w.pw0 = m.pw0; w.pw1 = m.pw1; w.steps = m.steps; w.more = false; w.link = 0;
  w.chk4 = ring_load32(ring, w.p + w.best - 3);
  w.tdist = trim == 0 ? 0xFFFFFFFFu : (w.p >= trim ? w.p - trim : 0u);  // candidates further back than this lie below `trim`
}
// up to `count` cheap steps; sets hit (candidate w.c passed the filter, w.link is its link) or done
__device__ __forceinline__ void walk_light(Walk &w, const Shared &sh, const RingView &ring, int count, bool &hit, bool &done) {
  const uint32_t chk = w.chk4 >> 24;  // the byte at p + best
#pragma unroll 1
  for (int k = 0; k < count; k++) {
    const uint32_t dist = (w.p16 - w.c) & 0xFFFFu;
    if (dist - w.last - 1u >= w.rng - w.last) { done = true; break; }  // empty / stale (dist <= last) / out of window
    w.last = dist;
    w.link = sh.prev[w.c & (kWindow - 1)];
    // one byte first (one load); the three before it only when that one agrees
    if (sh.ringb[(w.c + w.best) & (kRing - 1)] == chk && ring_load32(ring, w.c + w.best - 3) == w.chk4) { hit = true; break; }
    w.c = w.link;
    if (--w.steps <= 0) { done = true; w.more = true; break; }
    if (dist > w.tdist) { done = true; break; }
  }
}
// the full comparison of a candidate that passed the filter
__device__ __forceinline__ void walk_compare(Walk &w, const RingView &ring, int nice, bool &done) {
  const uint32_t len = match_compare(ring, w.p, w.c, w.pw0, w.pw1, w.max_len);  // (the ring is addressed modulo 2^16: c stands for p - dist)
  if (len > w.best) {
    w.best = len; w.best_dist = w.last;
    if (len >= (uint32_t)nice || len == w.max_len) { done = true; return; }
    w.chk4 = ring_load32(ring, w.p + len - 3);
  }
  w.c = w.link;
  if (--w.steps <= 0) { done = true; w.more = true; }
  else if (w.last > w.tdist) done = true;
}

// ---- which positions would a parse visit?  (a heuristic that feeds the deep walks) ------------------------------------
// Lane l walks the node graph from the start of its 64-position sub-range (nothing pending) to the end of that sub-range.
// The paths are not stitched together: paths merge after a few tokens, and the landings of the deep walks cover the
// stretch before they do.  The guess does not depend on where the real parse enters the tile, so the front end makes it
// right after the shallow walks.  One warp; writes the visited, walkable, unclaimed positions to vq and their number to *vqn.
__device__ __forceinline__ void guess_visits(const uint16_t *mlen, const uint32_t *claim, uint32_t ts, uint32_t te, uint32_t n, uint16_t *vq,
                                             uint32_t *vqn, int lane) {
  const uint32_t tl = te - ts, lo = 64u * lane, hi = min(lo + 64u, tl);
  uint64_t vis = 0;
  GUARD_DECL(g_g)
  for (uint32_t q = lo, k = 0; q < hi;) {
    GUARD(g_g, 1000u, 300);
    vis |= 1ull << (q - lo);
    uint32_t nk, em;
    q = lazy_next(ts + q, k, mlen[1 + q], mlen[q], nk, em) - ts;
    k = nk;
  }
  if (claim) vis &= ~((uint64_t)claim[2 * lane] | ((uint64_t)claim[2 * lane + 1] << 32));  // walked in an earlier round
  if (lo < tl && ts + lo + 64 + 4 > n)  // the member's last bytes cannot start a match
    for (uint32_t o = 0; o < 64; o++)
      if (ts + lo + o + 4 > n) vis &= ~(1ull << o);
  const uint32_t cnt = (uint32_t)__popcll(vis);
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  uint32_t at = incl - cnt;
  while (vis) {
    const uint32_t o = (uint32_t)__ffsll((long long)vis) - 1u;
    vis &= vis - 1;
    vq[at++] = (uint16_t)(lo + o);
  }
  if (lane == 31) *vqn = incl;
}

// ---- shallow walks of one tile: batches of 32 consecutive positions, handed out through a counter ---------------------
// Called by every warp that has nothing better to do (the front group; the parse warp when its tile is done).
__device__ __forceinline__ void shallow_batches(const Shared &sh, const Gen &G, uint32_t ts, uint32_t te, uint32_t n, const LevelParams &lp,
                                                uint32_t emit_from, int lane) {
  const RingView ring{sh.ring};
  const PrevView prevv{sh.prev};
  GUARD_DECL(g_s)
  for (;;) {
    GUARD(g_s, 1000u, 500);
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(&sh.sc[SC_BATCH], 1u);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= kTile / 32) break;
    const uint32_t i = b * 32 + lane, p = ts + i;
    uint32_t l = 0, d = 0, rt = p & 0xFFFFu;
    if (p < te && p + 4 <= n && p >= emit_from) {
      // every lane walks on its own: at these depths nearly every candidate passes the first test (same hash), so keeping
      // the warp together for the comparisons (as the deep walks do) only adds votes -- measured: 17.0 vs 11.4 clk/byte
      MatchState m;
      match_begin(m, ring, p, n, G.tok[i], lp.shallow);
      GUARD_DECL(g_m)
      while (!m.done) { GUARD(g_m, 100000u, 501); match_step(m, ring, prevv, lp.shallow_nice); }
      l = m.best >= (uint32_t)kMinMatch ? m.best : 0u;
      d = m.best_dist;
      rt = match_resume_token(m);
    }
    G.mlen[1 + i] = (uint16_t)l;
    G.mdist[1 + i] = (uint16_t)d;
    G.tok[i] = (uint16_t)rt;
  }
}

// ---- front end: stage, hash, partition, insert, shallow walks of one tile (FRONT_THREADS threads) -----------------
// ft: thread index inside the front group; fw: warp index inside the group.  Uses bar_front() only (and, with `helper`, one
// arrival at barrier 3 that tells the parse warp it may take shallow-walk batches of this tile).
// `state`: bits 0..31 = input bytes resident in the ring or on their way, bit 32 = a bulk copy is in flight.  Returns the new state.
// Positions below emit_from (a multiple of the tile) only prime the window: hashed and inserted, not searched.
template <int BW>
__device__ __noinline__ uint64_t front_end(int gen, const uint8_t *src, uint32_t n, uint32_t ts, uint64_t state,
                                           int level, int ft, int fw, int lane, uint32_t emit_from, bool helper) {
  constexpr int FRONT_WARPS = NWARPS - BW, FRONT_THREADS = FRONT_WARPS * 32;
  [[maybe_unused]] constexpr int BACK_THREADS = BW * 32;  // (the timing build names the first front thread by it)
  constexpr int kClasses = FRONT_WARPS, kRankRounds = (kChunk32 + FRONT_WARPS - 1) / FRONT_WARPS;
  static_assert(kClasses >= 16 && kClasses <= kMaxClasses, "a class must span at most 1024 hash values; the counters hold 31 classes");
  ZB_SMEM;
  const Shared sh = carve(smem_raw);
  const Gen G = gen_of(smem_raw, gen);
  const LevelParams lp = level_params(level);
  const RingView ring{sh.ring};
  const PrevView prevv{sh.prev};
  const uint32_t te = min(ts + (uint32_t)kTile, n), want = min(n, te + (uint32_t)kTile);
  const uint32_t lt_mask = (1u << lane) - 1u;
  PH_DECL
  // 1. stage input: this tile needs [.., want); the copy engine was asked for it one tile ago, and is now asked for the next
  uint32_t loaded = (uint32_t)state;
  bool in_flight = (state >> 32) & 1u;
  if ((((uintptr_t)src) & 15) == 0) {  // 16-byte aligned source (slots of the packed arena are): bulk copies
    uint32_t phases = sh.sc[SC_LD_PHASES];
    // issue [a, b) as one mbarrier phase (two copies when the ring wraps); the ragged tail of the member by plain stores
    auto issue = [&](uint32_t a, uint32_t b) -> bool {
      const uint32_t b16 = b == n ? (b & ~15u) : b;
      if (ft == 0) for (uint32_t i = max(a, b16); i < b; i++) sh.ringb[i & (kRing - 1)] = src[i];
      if (b16 <= a) return false;
      if (ft == 0) {
        const uint32_t len = b16 - a, ro = a & (kRing - 1), first = min(len, (uint32_t)kRing - ro);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(sh.ldbar, len);
        bulk_g2s(sh.ringb + ro, src + a, first, sh.ldbar);
        if (first < len) bulk_g2s(sh.ringb, src + a + first, len - first, sh.ldbar);
      }
      return true;
    };
    if (in_flight) { mbar_wait(sh.ldbar, (phases - 1) & 1u); in_flight = false; }
    if (loaded < want) {  // (first tile of a member: nothing was asked for yet)
      if (issue(loaded, want)) { phases++; mbar_wait(sh.ldbar, (phases - 1) & 1u); }
      loaded = want;
    }
    // A waiter tests the PARITY of a phase, so the next phase may only be opened once every thread has seen this one
    // complete: two completions behind a slow waiter's back look like none.  Hence the barrier before the prefetch.
    bar_front<BW>();
    const uint32_t next_want = min(n, want + (uint32_t)kTile);
    if (next_want > loaded) {
      if (issue(loaded, next_want)) { phases++; in_flight = true; }
      loaded = next_want;
    }
    if (ft == 0) sh.sc[SC_LD_PHASES] = phases;
  } else {
    for (uint32_t i = loaded + ft; i < want; i += FRONT_THREADS) sh.ringb[i & (kRing - 1)] = src[i];
    loaded = want;
  }
  for (int i = ft; i < kChunk32 * kClasses / 2; i += FRONT_THREADS) reinterpret_cast<uint32_t *>(sh.cnt)[i] = 0;
  if (ft == 0) sh.sc[SC_BATCH] = 0;
  bar_front<BW>();
  PH_AT(PH_F_STAGE, BACK_THREADS);
  // 2a. hashes (0xFFFF = position cannot start a match)
  for (int i = ft; i < kTile; i += FRONT_THREADS) {
    const uint32_t p = ts + i;
    sh.hsh[i] = (p < te && p + 4 <= n) ? (uint16_t)hash4(ring_load32(ring, p)) : (uint16_t)0xFFFF;
  }
  bar_front<BW>();
  // 2b. partition the tile's positions by hash class, keeping position order.  The class is the TOP of the hash
  //     ((h * classes) >> 14), so the heads of one class are neighbours in shared memory (no bank conflicts when a
  //     warp works on its class).  Batch b = positions [32 b, 32 b + 32), ranked by warp b mod classes.
  uint32_t lrank[kRankRounds];
#pragma unroll
  for (int r = 0; r < kRankRounds; r++) {
    const int b = fw + r * kClasses;
    lrank[r] = 0;
    if (b < kChunk32) {
      const uint32_t h = sh.hsh[b * 32 + lane];
      const bool valid = h != 0xFFFFu;
      const uint32_t cls = valid ? (h * (uint32_t)kClasses) >> kHashBits : 31u;
      const uint32_t m = match_any_bits<5>(cls);  // lanes of my class (invalid lanes form their own group)
      lrank[r] = __popc(m & lt_mask);
      if (valid && (m & lt_mask) == 0) sh.cnt[b * kClasses + cls] = (uint16_t)__popc(m);
    }
  }
  bar_front<BW>();
  // per class: exclusive offsets over the batches (one warp-wide scan per class: lane = pair of batches)
  {
    const int c = fw;  // kClasses == FRONT_WARPS
    const uint32_t a0 = sh.cnt[(2 * lane) * kClasses + c], a1 = sh.cnt[(2 * lane + 1) * kClasses + c];
    uint32_t incl = a0 + a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - a0 - a1;
    sh.cnt[(2 * lane) * kClasses + c] = (uint16_t)excl;
    sh.cnt[(2 * lane + 1) * kClasses + c] = (uint16_t)(excl + a0);
    if (lane == 31) sh.cstart[c + 1] = (uint16_t)incl;  // class total, turned into starts below
  }
  bar_front<BW>();
  if (fw == 0) {
    uint32_t v = lane < kClasses ? sh.cstart[lane + 1] : 0u, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncwarp();
    if (lane < kClasses) sh.cstart[lane + 1] = (uint16_t)incl;
    if (lane == 0) sh.cstart[0] = 0;
  }
  bar_front<BW>();
#pragma unroll
  for (int r = 0; r < kRankRounds; r++) {
    const int b = fw + r * kClasses;
    if (b < kChunk32) {
      const uint32_t i = b * 32 + lane, h = sh.hsh[i];
      if (h != 0xFFFFu) {
        const uint32_t cls = (h * (uint32_t)kClasses) >> kHashBits;
        sh.posl[sh.cstart[cls] + sh.cnt[b * kClasses + cls] + lrank[r]] = (uint16_t)i;
      }
    }
  }
  bar_front<BW>();
  PH_AT(PH_F_PART, BACK_THREADS);
  // 2c. chain insertion: warp w owns hash class w, walks its positions in order, 32 per step
  {
    const uint32_t c0 = sh.cstart[fw], c1 = sh.cstart[fw + 1];
    for (uint32_t base = c0; base < c1; base += 32) {
      const uint32_t k = base + lane;
      const bool valid = k < c1;
      const uint32_t i = valid ? sh.posl[k] : 0, p = ts + i;
      const uint32_t h = valid ? sh.hsh[i] : 0u;
      // lanes with my hash: inside a class the low 10 bits identify the hash; bit 10 keeps the lanes past the end of the
      // list apart
      const uint32_t m = match_any_bits<11>(valid ? (h & 1023u) : 1024u);
      const uint32_t below = m & lt_mask;
      const int srcl = below ? 31 - __clz(below) : 0;
      const uint32_t pp = __shfl_sync(0xffffffffu, p, srcl);
      uint16_t cand = 0;
      if (valid) cand = below ? (uint16_t)pp : sh.head[h];
      __syncwarp();  // every head is read before any head of this batch is replaced
      if (valid) {
        G.tok[i] = cand;
        sh.prev[p & (kWindow - 1)] = cand;
        if ((m >> lane) == 1u) sh.head[h] = (uint16_t)p;  // highest lane of the group
      }
      __syncwarp();
    }
  }
  bar_front<BW>();
  PH_AT(PH_F_INSERT, BACK_THREADS);
  // 3. shallow walk at every position; batches of 32 consecutive positions are handed out through a counter.  The back
  //    end's parse warp takes batches too once its own tile is done (ZB_BACK_HELPS): 64 batches over 31 warps are three
  //    rounds for two warps and two for the rest.
#if ZB_BACK_HELPS
  if (BW == 1 && helper) bar_walk_open_arrive();  // (behind bar_front: every front warp's links are in place)
#endif
  shallow_batches(sh, G, ts, te, n, lp, emit_from, lane);
  bar_front<BW>();
  PH_AT(PH_F_SHALLOW, BACK_THREADS);
  // 4. the positions a parse of this tile is guessed to visit: the deep walks' first queue
  if (fw == 0 && lp.rounds > 0 && ts >= emit_from) guess_visits(G.mlen, nullptr, ts, te, n, sh.vq + gen * kTile, &sh.sc[SC_VQN + gen], lane);
  return (uint64_t)loaded | ((uint64_t)in_flight << 32);
}

// ---- jump codes of the lazy parse: node v = 2 * (p - ts) + kind; codes >= kExit leave the tile -------------------
__device__ __forceinline__ void jump_codes(const Shared &sh, const Gen G, uint32_t ts, uint32_t te, int t0, int nthreads) {
  const uint32_t tl = te - ts;
  for (uint32_t i = t0; i < tl; i += nthreads) {
    const uint32_t p = ts + i, ml = G.mlen[1 + i], mp = G.mlen[i];
    uint32_t code[2];
#pragma unroll
    for (uint32_t k = 0; k < 2; k++) {
      uint32_t nk, em;
      const uint32_t np = lazy_next(p, k, ml, mp, nk, em);
      code[k] = np < te ? 2 * (np - ts) + nk : kExit + 2 * (np - te) + nk;
    }
    sh.jmp[i] = code[0] | (code[1] << 16);
  }
}

// ---- the lazy parse of one tile by one warp ------------------------------------------------------------------------
// Lane l owns positions [64 l, 64 l + 64) of the tile, i.e. nodes [128 l, 128 l + 128).  Every lane first walks its
// range from a guessed entry (its first position, nothing pending); paths through the node graph merge after a few
// tokens, so when the true path enters the range somewhere else it is followed only until it hits a guessed node.
// The lanes on the true path are found by pointer doubling over "which lane does my exit land in"; entries are
// handed over through shared memory; the loop ends when no lane's entry changed (typically 2 rounds).
// e0 < kExit: entry node.  Returns the exit code (>= kExit); m0 / m1 = this lane's marks for kind 0 / 1 nodes.
__device__ __noinline__ uint32_t warp_parse(const uint32_t *jmp, uint32_t *pslot, uint32_t e0, uint32_t tl, int lane, uint64_t *marks) {
  const uint32_t lo = 128u * lane, hi = lo + 128u;
  uint64_t m0 = 0, m1 = 0;
  auto code = [&](uint32_t v) { const uint32_t w = jmp[v >> 1]; return (v & 1u) ? (w >> 16) : (w & 0xFFFFu); };
  uint64_t g0 = 0, g1 = 0;
  uint32_t gexit = kExit;
  if (64u * lane < tl) {
    uint32_t v = lo;
    GUARD_DECL(g_p)
    while (v < hi) {
      GUARD(g_p, 1000u, 400);
      const uint64_t bit = 1ull << ((v >> 1) & 63u);
      if (v & 1u) g1 |= bit; else g0 |= bit;
      v = code(v);
    }
    gexit = v;
  }
  uint32_t ent = 0xFFFFFFFFu, xit = gexit;
  const uint32_t L0 = e0 >> 7;
  bool reached = false;
  for (int it = 0; it < 40; it++) {
    const uint32_t succ = xit < kExit ? (xit >> 7) : 32u;
    uint32_t reach = 1u << L0, jump = succ;
#pragma unroll
    for (int r = 0; r < 5; r++) {
      const uint32_t contrib = (((reach >> lane) & 1u) && jump < 32u) ? (1u << jump) : 0u;
      reach |= __reduce_or_sync(0xffffffffu, contrib);
      const uint32_t nj = __shfl_sync(0xffffffffu, jump, jump & 31u);
      jump = jump < 32u ? nj : 32u;
    }
    reached = (reach >> lane) & 1u;
    if (reached && succ < 32u) pslot[succ] = xit;
    __syncwarp();
    const uint32_t newent = lane == (int)L0 ? e0 : (reached ? pslot[lane] : 0xFFFFFFFFu);
    __syncwarp();
    const bool dirty = reached && newent != ent;
    if (!__any_sync(0xffffffffu, dirty)) break;
    if (dirty) {
      ent = newent;
      uint64_t t0 = 0, t1 = 0;
      uint32_t v = ent;
      bool merged = false;
      GUARD_DECL(g_q)
      while (v < hi) {
        GUARD(g_q, 1000u, 401);
        const uint32_t o = (v >> 1) & 63u;
        const uint64_t bit = 1ull << o;
        if (((v & 1u) ? g1 : g0) & bit) {
          const uint64_t ge = ~0ull << o;
          t0 |= g0 & ge; t1 |= g1 & ge;
          merged = true;
          break;
        }
        if (v & 1u) t1 |= bit; else t0 |= bit;
        v = code(v);
      }
      xit = merged ? gexit : v;
      m0 = t0; m1 = t1;
    }
  }
  if (!reached) { m0 = 0; m1 = 0; }
  marks[0] = m0; marks[1] = m1;
  const uint32_t succ = xit < kExit ? (xit >> 7) : 32u;
  return __reduce_max_sync(0xffffffffu, (reached && succ == 32u) ? xit : 0u);
}

// ---- deep walks of the queued positions (BACK_THREADS threads, persistent lanes) -------------------------------------
// where a match taken at tile position i would land, and the lazy look-ahead behind it: queued for the next wave
__device__ __forceinline__ void deep_landings(uint32_t i, const uint16_t *mlen, const uint32_t *claim, uint32_t tl, uint32_t ts, uint32_t n,
                                              uint16_t *oq, uint32_t *oqn) {
  const uint32_t L = mlen[1 + i];
  if (L < (uint32_t)kMinMatch) return;
#pragma unroll
  for (uint32_t k = 0; k < 2; k++) {
    const uint32_t q = i + L + k;
    if (q < tl && ts + q + 4 <= n && !((claim[q >> 5] >> (q & 31u)) & 1u)) oq[atomicAdd(oqn, 1u)] = (uint16_t)q;
  }
}

// Wave 0 walks the guessed positions vq[0, qn); every finished position appends the positions a match taken there would
// land on to oq, which the next wave walks, lp.hops waves deep.  A position is walked at most once (claim bits) and what a
// walk finds does not depend on any other walk, so the result is independent of the scheduling inside a wave.
template <int BW>
__device__ __noinline__ void deep_walks(int gen, uint32_t qn, uint32_t ts, uint32_t te, uint32_t n, uint32_t trim, int level, int bt, int lane) {
  ZB_SMEM;
  const Shared sh = carve(smem_raw);
  const Gen G = gen_of(smem_raw, gen);
  const LevelParams lp = level_params(level);
  const RingView ring{sh.ring};
  const PrevView prevv{sh.prev};
  const uint32_t tl = te - ts;
  const int extra = lp.depth - lp.shallow;
  const uint16_t *queue = sh.vq + gen * kTile;
  uint32_t qbeg = 0, qend = qn;
  for (int wave = 0;; wave++) {
    const bool landings = wave < lp.hops;
    for (uint32_t base = qbeg; base < qend; base += BW * 32) {  // passes: one position per lane
      bool have = false, hit = false, done = false;
      uint32_t i = 0;
      Walk w;
      const uint32_t idx = base + bt;
      if (idx < qend) {
        const uint32_t cand = queue[idx];
        const uint32_t bit = 1u << (cand & 31u);
        const uint32_t old = atomicOr(&sh.claim[cand >> 5], bit);
        if (!(old & bit)) {
          i = cand;
          MatchState m;
          if (match_resume(m, ring, prevv, ts + i, n, G.mlen[1 + i], G.mdist[1 + i], G.tok[i], extra, trim)) { have = true; walk_from(w, m, ring, trim); }
          else if (landings) deep_landings(i, G.mlen, sh.claim, tl, ts, n, sh.oq, &sh.sc[SC_OQN]);
        }
      }
      GUARD_DECL(g_it)
      while (__any_sync(0xffffffffu, have)) {
        GUARD(g_it, 200000u, 200 + wave);
        if (have && !hit && !done) walk_light(w, sh, ring, 16, hit, done);
        if (__any_sync(0xffffffffu, hit)) {
          if (hit) { hit = false; walk_compare(w, ring, lp.nice, done); }
        }
        if (done) {  // result, landings
          if (w.best >= (uint32_t)kMinMatch && w.best > G.mlen[1 + i]) { G.mlen[1 + i] = (uint16_t)w.best; G.mdist[1 + i] = (uint16_t)w.best_dist; }
          if (landings) deep_landings(i, G.mlen, sh.claim, tl, ts, n, sh.oq, &sh.sc[SC_OQN]);
          have = false; done = false;
        }
      }
    }
    bar_back<BW>();  // every walk of this wave has finished: lengths are final, landings are queued
    if (!landings) break;
    const uint32_t on = sh.sc[SC_OQN];
    const uint32_t obeg = wave == 0 ? 0u : qend;  // the landings of this wave are oq[obeg, on)
    if (on == obeg) break;
    bar_back<BW>();  // everybody has read SC_OQN before the next wave appends
    queue = sh.oq;
    qbeg = obeg;
    qend = on;
  }
}

// ---- one member ----------------------------------------------------------------------------------------------
template <int BW>
__device__ void encode_member(const DeflateTask t, int level, uint32_t *toks, DeflateResult *res, uint32_t *blk_lens) {
  constexpr int BACK_WARPS = BW, BACK_THREADS = BW * 32;
  ZB_SMEM;
  const Shared sh = carve(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool back = warp < BACK_WARPS;
  const int ft = tid - BACK_THREADS, fw = warp - BACK_WARPS;
  // a primed member: `prime` bytes right before t.src are run through the front end only (ring, hash chains), so that the
  // member's matches can reach back into them; positions count from the first primed byte
  const uint32_t prime = ((t.flags >> kDeflatePrimeShift) & kDeflatePrimeMask) * kDeflatePrimeTile;
  const uint8_t *src = t.src - prime;
  const uint32_t n = (uint32_t)t.src_len + prime;
  uint32_t *out_words = reinterpret_cast<uint32_t *>(t.dst);
  const uint64_t out_cap_words = t.dst_cap / 4;
  const LevelParams lp = level_params(level);
  const bool not_final = (t.flags & kDeflateNotFinal) != 0;

  for (int i = tid; i < (1 << kHashBits) / 2; i += THREADS) reinterpret_cast<uint32_t *>(sh.head)[i] = 0;
  for (int i = tid; i < kWindow / 2; i += THREADS) reinterpret_cast<uint32_t *>(sh.prev)[i] = 0;
  for (int i = tid; i < 320; i += THREADS) sh.hist_l[i] = 0;
  if (tid == 0) {
    sh.sc[SC_OUTW] = 0; sh.sc[SC_CARRY] = 0; sh.sc[SC_CBITS] = 0; sh.sc[SC_OVERFLOW] = 0; sh.sc[SC_BLK_SRCLEN] = 0;
    gen_of(smem_raw, 0).mlen[0] = 0; gen_of(smem_raw, 0).mdist[0] = 0;
  }
  __syncthreads();

  uint32_t pos = 0, kind = 0, ntok = 0, nblocks = 0;
  uint64_t loaded = 0;  // staging state of the front end (front threads only)
  uint64_t blk_src_start = prime;
  int tiles_in_block = 0, g = 0;
  PH_DECL

  // prologue: tile 0 through the front end, then its jump codes
  if (n) {
    if (!back) loaded = front_end<BW>(0, src, n, 0, loaded, level, ft, fw, lane, prime, false);
    __syncthreads();
    jump_codes(sh, gen_of(smem_raw, 0), 0, min((uint32_t)kTile, n), tid, THREADS);
    __syncthreads();
  }
  PH(PH_OTHER);

  for (uint32_t ts = 0; ts < n; ts += kTile, g ^= 1) {
    const uint32_t te = min(ts + (uint32_t)kTile, n), tl = te - ts;
    const Gen G = gen_of(smem_raw, g), G1 = gen_of(smem_raw, g ^ 1);
    const bool more = te < n;
    const bool priming = te <= prime;  // (prime is a multiple of the tile: a tile is primed as a whole or not at all)
    if (back && priming) {
      // nothing is emitted for this tile: no marks, and the parse of the next tile starts afresh at its first position
      if (warp == 0) {
        sh.mk0[2 * lane] = 0; sh.mk0[2 * lane + 1] = 0; sh.mk1[2 * lane] = 0; sh.mk1[2 * lane + 1] = 0;
        if (lane == 0) sh.sc[SC_EXIT] = kExit;
      }
    } else if (back) {
      // ---- back end: parse, deep walks, final parse of tile t ----------------------------------------------------
      const uint32_t te1 = min(te + (uint32_t)kTile, n);
      const uint32_t trim = (more && te1 > (uint32_t)kWindow) ? te1 - (uint32_t)kWindow : 0u;
      const uint32_t e0 = pos < te ? 2 * (pos - ts) + kind : kExit;  // kExit: a match of an earlier tile covers this one
      uint64_t mk[2] = {0, 0};
      for (int round = 0; round < lp.rounds; round++) {
        // round 0 walks the positions the front end guessed (vq of this generation); later rounds guess again over the
        // improved lengths
        if (round == 0) { if (tid < 64) sh.claim[tid] = 0; }
        else if (warp == 0) guess_visits(G.mlen, sh.claim, ts, te, n, sh.vq + g * kTile, &sh.sc[SC_VQN + g], lane);
        if (tid == 0) sh.sc[SC_OQN] = 0;
        bar_back<BW>();
        PH(PH_PARSE0);
        deep_walks<BW>(g, sh.sc[SC_VQN + g], ts, te, n, trim, level, tid, lane);
        PH(PH_DEEP);
        jump_codes(sh, G, ts, te, tid, BACK_THREADS);
        bar_back<BW>();
        PH(PH_JUMP);
      }
      if (warp == 0) {
        uint32_t ex = e0;
        if (e0 < kExit) ex = warp_parse(sh.jmp, sh.pslot, e0, tl, lane, mk);
        else { mk[0] = 0; mk[1] = 0; ex = kExit + 2 * (pos - te) + kind; }
        sh.mk0[2 * lane] = (uint32_t)mk[0]; sh.mk0[2 * lane + 1] = (uint32_t)(mk[0] >> 32);
        sh.mk1[2 * lane] = (uint32_t)mk[1]; sh.mk1[2 * lane + 1] = (uint32_t)(mk[1] >> 32);
        if (lane == 0) sh.sc[SC_EXIT] = ex;
      }
      PH(PH_PARSE1);
    } else if (more) {
      // ---- front end: tile t + 1 -------------------------------------------------------------------------------------
      loaded = front_end<BW>(g ^ 1, src, n, te, loaded, level, ft, fw, lane, prime, true);
    }
#if ZB_BACK_HELPS
    if (BW == 1 && back && more) {
      // the parse warp is done with tile t: it takes batches of tile t + 1's shallow walks as soon as the front end hands
      // them out (the chains are complete then); what a walk finds depends on its position alone
      bar_walk_open_wait();
      shallow_batches(sh, G1, te, min(te + (uint32_t)kTile, n), n, lp, prime, lane);
    }
#endif
    __syncthreads();
    PH(PH_BACK_WAIT);
    // ---- all warps: tokens of tile t -------------------------------------------------------------------------------------
    {
      const uint32_t ex = sh.sc[SC_EXIT] - kExit;
      pos = te + (ex >> 1);
      kind = ex & 1u;
      if (tid == 0 && more) { G1.mlen[0] = G.mlen[tl]; G1.mdist[0] = G.mdist[tl]; }
      // tokens of the visited nodes, in position order (2 * PPT consecutive nodes per thread)
      uint32_t tk[2 * PPT], cnt = 0, srcsum = 0;
      const uint32_t w0 = sh.mk0[(tid * PPT) >> 5] >> ((tid * PPT) & 31u), w1 = sh.mk1[(tid * PPT) >> 5] >> ((tid * PPT) & 31u);
#pragma unroll
      for (int j = 0; j < 2 * PPT; j++) {
        const uint32_t i = tid * PPT + (j >> 1), k = j & 1u, p = ts + i;
        const bool marked = ((k ? w1 : w0) >> (j >> 1)) & 1u;
        if (!marked || p >= te) continue;
        uint32_t nk, em;
        const uint32_t mp = G.mlen[i];
        lazy_next(p, k, G.mlen[1 + i], mp, nk, em);
        if (em == 1 || em == 2) {
          const uint32_t b = sh.ringb[(em == 1 ? p : p - 1) & (kRing - 1)];
          tk[cnt++] = tok_lit(b);
          atomicAdd(&sh.hist_l[b], 1u);
          srcsum += 1;
        } else if (em == 3) {
          uint32_t d = G.mdist[i], eb, ev;
          tk[cnt++] = tok_match(mp, d);
          atomicAdd(&sh.hist_l[len_sym_of(mp, eb, ev)], 1u);
          atomicAdd(&sh.hist_d[dist_sym_of(d, eb, ev)], 1u);
          srcsum += mp;
        }
      }
      uint32_t total;
      const uint32_t off = block_scan(cnt, sh.scan, &total);
      for (uint32_t j = 0; j < cnt; j++) toks[ntok + off + j] = tk[j];
      ntok += total;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) srcsum += __shfl_xor_sync(0xffffffffu, srcsum, o);
      if (lane == 0 && srcsum) atomicAdd(&sh.sc[SC_BLK_SRCLEN], srcsum);
    }
    __syncthreads();
    PH(PH_TOKENS);
    if (!priming) tiles_in_block++;
    const bool last = te == n;
    if (tiles_in_block == kTilesPerBlock || last) {
      const uint32_t blen = sh.sc[SC_BLK_SRCLEN];
      finalize_block(src, blk_src_start, ntok, toks, last && !not_final, out_words, out_cap_words);
      PH_RESET
      if (tid == 0 && blk_lens) blk_lens[t.blk_off + nblocks] = blen;
      blk_src_start += blen;
      ntok = 0; tiles_in_block = 0; nblocks++;
    }
    // jump codes of tile t + 1 (its slot 0 now carries position te - 1)
    if (more) jump_codes(sh, G1, te, min(te + (uint32_t)kTile, n), tid, THREADS);
    __syncthreads();
    PH(PH_OTHER);
  }
  if (n == 0) {  // (an empty member is never primed)
    finalize_block(src, 0, 0, toks, !not_final, out_words, out_cap_words);
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

template <int BW>
__global__ void __launch_bounds__(THREADS, 1)
deflate_kernel(const DeflateTask *__restrict__ tasks, uint32_t ntasks, DeflateResult *__restrict__ results,
               unsigned int *__restrict__ queue, uint32_t *__restrict__ tok_scratch, int level,
               uint32_t *__restrict__ blk_lens) {
  ZB_SMEM;
  const Shared sh = carve(smem_raw);
  uint32_t *toks = tok_scratch + (size_t)blockIdx.x * kTokCap;
  if (threadIdx.x == 0) { mbar_init(sh.ldbar, 1); sh.sc[SC_LD_PHASES] = 0; }
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh.sc[SC_TASK] = atomicAdd(queue, 1u);
    __syncthreads();
    const uint32_t task = sh.sc[SC_TASK];
    if (task >= ntasks) break;
    encode_member<BW>(tasks[task], level, toks, &results[task], blk_lens);
  }
}

// level `None (reference :1106-1116): stored blocks only.  Blocks hold 65534 bytes as in the reference (:747-750), so the
// output is byte-identical to Zipc_deflate.deflate ~level:`None.  A copy, so it should run at the speed of the memory: every
// stored block is a unit of its own (header + 65534 bytes), `per_task` CTAs share the blocks of a member (a 64 MiB member is
// 1025 blocks for the whole grid, not one CTA's), and the bytes move as aligned 32-bit words -- block b's payload lands at
// 65539 b + 5, its source starts at 65534 b, so the two are misaligned against each other by a different amount in every
// block: aligned loads, funnel shift, aligned stores (the gather kernel's method, zip_api.cu).
__device__ __forceinline__ void cta_copy_words(uint8_t *dst, const uint8_t *src, uint64_t len) {
  uint64_t head = (4 - ((uintptr_t)dst & 3)) & 3;
  if (head > len) head = len;
  for (uint64_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
  const uint8_t *sb = src + head;
  const uint32_t sh = (uint32_t)((uintptr_t)sb & 3) * 8;
  const uint32_t *s = reinterpret_cast<const uint32_t *>(sb - (sh >> 3));
  uint32_t *t = reinterpret_cast<uint32_t *>(dst + head);
  const uint64_t nw = (len - head) >> 2;
  if (sh == 0) {
    for (uint64_t i = threadIdx.x; i < nw; i += blockDim.x) t[i] = __ldg(s + i);
  } else {
    // word i of the destination = bytes [4i + sh/8, 4i + sh/8 + 4) of the aligned source words: s[i + 1] holds at least one
    // byte of the range, so it lies inside the source's last touched word
    for (uint64_t i = threadIdx.x; i < nw; i += blockDim.x) t[i] = __funnelshift_r(__ldg(s + i), __ldg(s + i + 1), sh);
  }
  for (uint64_t i = head + (nw << 2) + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(256)
stored_kernel(const DeflateTask *__restrict__ tasks, uint32_t ntasks, DeflateResult *__restrict__ results, uint32_t per_task) {
  const uint32_t task = blockIdx.x / per_task, sub = blockIdx.x % per_task;
  if (task >= ntasks) return;
  const DeflateTask t = tasks[task];
  constexpr uint64_t B = kStoredBlock;
  const uint64_t n = t.src_len, nblk = n ? (n + B - 1) / B : 1, need = n + 5 * nblk;
  if (need > t.dst_cap) {
    if (sub == 0 && threadIdx.x == 0) { results[task].out_len = 0; results[task].status = ZIPC_ERR_DST_TOO_SMALL; results[task].blocks = 0; }
    return;
  }
  for (uint64_t b = sub; b < nblk; b += per_task) {
    const uint64_t len = b + 1 < nblk ? B : n - b * B;
    uint8_t *h = t.dst + b * (B + 5);
    if (threadIdx.x == 0) {
      h[0] = (b + 1 == nblk && !(t.flags & kDeflateNotFinal)) ? 1 : 0;
      h[1] = (uint8_t)len; h[2] = (uint8_t)(len >> 8); h[3] = (uint8_t)~len; h[4] = (uint8_t)(~len >> 8);
    }
    cta_copy_words(h + 5, t.src + b * B, len);
  }
  if (sub == 0 && threadIdx.x == 0) { results[task].out_len = need; results[task].status = ZIPC_OK; results[task].blocks = (uint32_t)nblk; }
}

unsigned long long g_attr_devs = 0;  // bit d: attributes set on device d (function attributes are per device)

}  // namespace

int deflate_launch(zipc_b200_ctx *ctx, const DeflateTask *d_tasks, uint32_t n, DeflateResult *d_results, int level,
                   uint32_t *d_blk_lens, uint64_t max_src_len) {
  if (n == 0) return ZIPC_OK;
  if (level == ZIPC_LEVEL_NONE) {
    // CTAs per member: as many as its largest member has stored blocks, while the grid stays within ~64 CTAs per SM
    const uint64_t blocks = max_src_len ? (max_src_len + kStoredBlock - 1) / kStoredBlock : 1;
    const uint32_t per_task = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * 64 / n));
    {
      KernelTimer kt(ctx);
      stored_kernel<<<n * per_task, 256, 0, ctx->stream>>>(d_tasks, n, d_results, per_task);
    }
    ctx->launches++;
    ZB_CUDA(ctx, cudaGetLastError());
    return ZIPC_OK;
  }
  if (!(g_attr_devs >> (ctx->device & 63) & 1ull)) {
    ZB_CUDA(ctx, cudaFuncSetAttribute(deflate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    ZB_CUDA(ctx, cudaFuncSetAttribute(deflate_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
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
    // deep walks (rounds > 0) want half of the warps in the back end; without them one parse warp is enough
    if (dfl::level_params(level).rounds > 0)
      deflate_kernel<16><<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint32_t>(), level, d_blk_lens);
    else
      deflate_kernel<1><<<grid, THREADS, kSmemBytes, ctx->stream>>>(d_tasks, n, d_results, queue, ctx->d_scratch.as<uint32_t>(), level, d_blk_lens);
  }
  ctx->launches++;
  ZB_CUDA(ctx, cudaGetLastError());
  return ZIPC_OK;
}

}  // namespace zb

#ifdef ZB_DEFLATE_TIMING
extern "C" int zipc_b200_debug_deflate_phases(unsigned long long *out16, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, zb::g_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(zb::g_phase, z, sizeof z); }
  return 0;
}
extern "C" int zipc_b200_debug_deflate_counters(unsigned long long *out16, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, zb::g_dbg, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(zb::g_dbg, z, sizeof z); }
  return 0;
}
#endif

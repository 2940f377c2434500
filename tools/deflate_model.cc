// deflate_model.cc -- host-side model of the GPU encoder's algorithm (TEST SCAFFOLDING, not product).
// It drives the same __host__ __device__ building blocks (zipc_b200/csrc/deflate_core.h) serially, so
// its output must equal the kernel's byte for byte; tests use it to localise kernel bugs and to measure
// the compression ratio of the algorithm without a GPU.   Build: tools/build_model.sh
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../zipc_b200/csrc/deflate_core.h"

using namespace zb::dfl;

namespace {
struct Ring {
  std::vector<uint32_t> w;
  Ring() : w(kRing / 4, 0) {}
  uint32_t word(uint32_t a) const { return w[a]; }
  uint32_t byte(uint32_t i) const { return reinterpret_cast<const uint8_t *>(w.data())[i]; }
  void put(uint32_t pos, uint8_t b) {
    uint32_t i = pos & (kRing - 1);
    reinterpret_cast<uint8_t *>(w.data())[i] = b;
  }
};
struct Prev {
  std::vector<uint16_t> l;
  Prev() : l(kWindow, 0) {}
  uint32_t link(uint32_t pos) const { return l[pos & (kWindow - 1)]; }
};

struct Block {
  std::vector<uint32_t> toks;
  uint32_t fl[kNumLit], fd[kNumDist];
  uint64_t src_start = 0, src_len = 0;
  void reset(uint64_t start) { toks.clear(); memset(fl, 0, sizeof fl); memset(fd, 0, sizeof fd); src_start = start; src_len = 0; }
};

void lengths_for(const uint32_t *freq, int nsym, int max_bits, uint8_t *len) {
  std::vector<uint32_t> keys, scratch;
  for (int s = 0; s < nsym; s++) if (freq[s]) keys.push_back((freq[s] << 9) | (uint32_t)s);
  std::sort(keys.begin(), keys.end());
  scratch.resize(keys.size() + 1);
  huff_lengths_from_sorted(keys.data(), (int)keys.size(), nsym, max_bits, len, scratch.data());
}

void finalize_block(Block &b, const uint8_t *src, bool final, BitSink &out) {
  b.fl[256] += 1;
  uint8_t ll[kNumLit + kNumDist + 2], cl[kNumClen];
  lengths_for(b.fl, kNumLit, 15, ll);
  int hlit = kNumLit;
  while (hlit > 257 && ll[hlit - 1] == 0) hlit--;
  uint8_t dl[kNumDist];
  lengths_for(b.fd, kNumDist, 15, dl);
  int hdist = kNumDist;
  while (hdist > 1 && dl[hdist - 1] == 0) hdist--;
  uint8_t both[kNumLit + kNumDist];
  for (int i = 0; i < hlit; i++) both[i] = ll[i];
  for (int i = 0; i < hdist; i++) both[hlit + i] = dl[i];
  uint16_t rsyms[kNumLit + kNumDist];
  uint32_t cfreq[kNumClen];
  int nr = rle_code_lengths(both, hlit + hdist, rsyms, cfreq);
  lengths_for(cfreq, kNumClen, 7, cl);
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int hclen = 19;
  while (hclen > 4 && cl[order[hclen - 1]] == 0) hclen--;
  // exact costs (reference :1045-1079)
  uint64_t sym_dyn = 0, sym_fix = 0;
  for (int s = 0; s < kNumLit; s++) {
    uint32_t eb = s >= 257 ? len_extra_bits_of_sym((uint32_t)s) : 0;
    sym_dyn += (uint64_t)b.fl[s] * (ll[s] + eb);
    sym_fix += (uint64_t)b.fl[s] * (fixed_lit_len((uint32_t)s) + eb);
  }
  for (int s = 0; s < kNumDist; s++) {
    uint32_t eb = dist_extra_bits_of_sym((uint32_t)s);
    sym_dyn += (uint64_t)b.fd[s] * (dl[s] + eb);
    sym_fix += (uint64_t)b.fd[s] * (5 + eb);
  }
  uint64_t hdr_dyn = 3 + 5 + 5 + 4 + 3 * (uint64_t)hclen;
  for (int s = 0; s < kNumClen; s++) hdr_dyn += (uint64_t)cfreq[s] * (cl[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0));
  uint64_t dlen = hdr_dyn + sym_dyn, flen = 3 + sym_fix;
  uint64_t nlen = 3 + ((8 - ((out.bits() + 3) & 7)) & 7) + 32 + 8 * b.src_len;
  uint32_t lcode[kNumLit], dcode[kNumDist], ccode[kNumClen];
  if (nlen <= dlen && nlen <= flen) {
    out.put(final ? 1 : 0, 3);
    out.put(0, (8 - (out.bits() & 7)) & 7);
    out.put((uint32_t)b.src_len, 16);
    out.put((uint32_t)(~b.src_len) & 0xFFFF, 16);
    for (uint64_t i = 0; i < b.src_len; i++) out.put(src[b.src_start + i], 8);
    return;
  }
  if (flen <= dlen) {
    uint8_t fl8[288], fd8[32];
    for (int s = 0; s < 288; s++) fl8[s] = (uint8_t)fixed_lit_len((uint32_t)s);
    for (int s = 0; s < 32; s++) fd8[s] = 5;
    uint32_t fl_code[288], fd_code[32];
    canonical_codes(fl8, 288, fl_code);
    canonical_codes(fd8, 32, fd_code);
    memcpy(lcode, fl_code, sizeof lcode);
    memcpy(dcode, fd_code, sizeof dcode);
    out.put(final ? 3 : 2, 3);
  } else {
    canonical_codes(ll, kNumLit, lcode);
    canonical_codes(dl, kNumDist, dcode);
    canonical_codes(cl, kNumClen, ccode);
    out.put(final ? 5 : 4, 3);
    out.put((uint32_t)(hlit - 257), 5);
    out.put((uint32_t)(hdist - 1), 5);
    out.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) out.put(cl[order[i]], 3);
    for (int i = 0; i < nr; i++) {
      uint32_t s = rsyms[i] & 0xFF, ex = rsyms[i] >> 8;
      out.put(ccode[s] & 0xFFFF, ccode[s] >> 16);
      if (s == 16) out.put(ex, 2);
      else if (s == 17) out.put(ex, 3);
      else if (s == 18) out.put(ex, 7);
    }
  }
  for (uint32_t t : b.toks) {
    uint32_t dist = t >> 9, len = t & 0x1FF;
    if (!dist) { out.put(lcode[len] & 0xFFFF, lcode[len] >> 16); continue; }
    uint32_t eb, ev, s = len_sym_of(len, eb, ev);
    out.put(lcode[s] & 0xFFFF, lcode[s] >> 16);
    out.put(ev, eb);
    s = dist_sym_of(dist, eb, ev);
    out.put(dcode[s] & 0xFFFF, dcode[s] >> 16);
    out.put(ev, eb);
  }
  out.put(lcode[256] & 0xFFFF, lcode[256] >> 16);
}
}  // namespace

extern "C" {
// returns 0; *out is malloc'ed.  stats[0..2] = stored/fixed/dynamic... (optional, may be null)
int zipc_model_deflate(int level, const uint8_t *src, uint64_t n64, uint8_t **out_p, uint64_t *out_len) {
  const uint32_t n = (uint32_t)n64;
  LevelParams lp = level_params(level);
  Ring ring;
  Prev prev;
  std::vector<uint16_t> head(1u << kHashBits, 0), first(kTile, 0), mlen(kTile + 1, 0), mdist(kTile + 1, 0);
  std::vector<uint8_t> obuf((size_t)n + n / 8 + 1024);
  BitSink sink{obuf.data(), 0, 0, 0};
  Block blk;
  blk.reset(0);
  uint32_t pos = 0, kind = 0, carry_len = 0, carry_dist = 0, loaded = 0;
  int tiles_in_block = 0;
  for (uint32_t ts = 0; ts < n; ts += kTile) {
    uint32_t te = std::min(ts + (uint32_t)kTile, n);
    uint32_t want = std::min(n, te + (uint32_t)kTile);
    for (; loaded < want; loaded++) ring.put(loaded, src[loaded]);
    for (uint32_t p = ts; p < te; p++) {
      if (p + 4 > n) { first[p - ts] = 0; continue; }
      uint32_t h = hash4(ring_load32(ring, p));
      first[p - ts] = head[h];
      prev.l[p & (kWindow - 1)] = head[h];
      head[h] = (uint16_t)p;
    }
    // slot 0 of mlen/mdist holds position ts-1 (carry), slot 1+i holds ts+i
    mlen[0] = (uint16_t)carry_len; mdist[0] = (uint16_t)carry_dist;
    for (uint32_t p = ts; p < te; p++) {
      uint32_t d = 0, l = 0;
      if (p + 4 <= n) l = find_match(ring, prev, p, n, first[p - ts], lp.depth, lp.nice, d);
      mlen[1 + p - ts] = (uint16_t)l; mdist[1 + p - ts] = (uint16_t)(d & 0xFFFF);
    }
    while (pos < te) {
      uint32_t nk, em;
      uint32_t ml = mlen[1 + pos - ts], mp = mlen[pos - ts];
      uint32_t np = lazy_next(pos, kind, ml, mp, nk, em);
      if (em == 1) { uint8_t b = src[pos]; blk.toks.push_back(tok_lit(b)); blk.fl[b]++; blk.src_len += 1; }
      else if (em == 2) { uint8_t b = src[pos - 1]; blk.toks.push_back(tok_lit(b)); blk.fl[b]++; blk.src_len += 1; }
      else if (em == 3) {
        uint32_t d = mdist[pos - ts] ? mdist[pos - ts] : 65536u;
        if (d == 65536u) d = 32768;  // 32768 is stored as 0x8000, never 0; kept for clarity
        uint32_t eb, ev;
        blk.toks.push_back(tok_match(mp, d));
        blk.fl[len_sym_of(mp, eb, ev)]++; blk.fd[dist_sym_of(d, eb, ev)]++;
        blk.src_len += mp;
      }
      pos = np; kind = nk;
    }
    carry_len = mlen[te - ts]; carry_dist = mdist[te - ts];  // position te-1
    tiles_in_block++;
    bool last = te == n;
    if (tiles_in_block == kTilesPerBlock || last) {
      uint64_t next_start = blk.src_start + blk.src_len;
      finalize_block(blk, src, last, sink);
      blk.reset(next_start);
      tiles_in_block = 0;
    }
  }
  if (n == 0) finalize_block(blk, src, true, sink);
  sink.flush();
  uint8_t *o = (uint8_t *)malloc(sink.pos ? sink.pos : 1);
  memcpy(o, obuf.data(), sink.pos);
  *out_p = o; *out_len = sink.pos;
  return 0;
}
void zipc_model_free(void *p) { free(p); }
}

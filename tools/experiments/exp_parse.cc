// exp_parse.cc -- CPU experiment (not product, not test): compression ratio and chain-step counts of encoder variants
// built from deflate_core.h: shallow search everywhere + deep search only at the positions the lazy parse visits.
//   g++ -O2 -std=c++17 -o /tmp/exp_parse tools/experiments/exp_parse.cc zipc_b200/csrc/synth.cc && /tmp/exp_parse
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../zipc_b200/csrc/deflate_core.h"
extern "C" void zipc_b200_synth_text(uint64_t seed, void *out, size_t n);
extern "C" void zipc_b200_synth_rand(uint64_t seed, void *out, size_t n);
using namespace zb::dfl;

struct Ring {
  std::vector<uint32_t> w;
  Ring() : w(kRing / 4, 0) {}
  uint32_t word(uint32_t a) const { return w[a]; }
  uint32_t byte(uint32_t i) const { return reinterpret_cast<const uint8_t *>(w.data())[i]; }
  void put(uint32_t pos, uint8_t b) { reinterpret_cast<uint8_t *>(w.data())[pos & (kRing - 1)] = b; }
};
struct Prev {
  std::vector<uint16_t> l;
  Prev() : l(kWindow, 0) {}
  uint32_t link(uint32_t pos) const { return l[pos & (kWindow - 1)]; }
};

static uint64_t g_steps = 0;
template <class R, class P>
uint32_t fm(const R &ring, const P &prev, uint32_t p, uint32_t n, uint32_t first, int depth, int nice, uint32_t &dist, uint32_t min_best = kMinMatch - 1) {
  MatchState m;
  match_begin(m, ring, p, n, first, depth);
  if (min_best > m.best && min_best < m.max_len) { m.best = min_best; m.chk = ring_load8(ring, p + min_best); }
  while (!m.done) { match_step(m, ring, prev, nice); g_steps++; }
  dist = m.best_dist;
  return (m.best >= (uint32_t)kMinMatch && m.best_dist) ? m.best : 0;
}

// bit cost of a block given histograms (dynamic only, approximates the real writer well enough for comparisons)
static void lengths_for(const uint32_t *freq, int nsym, int max_bits, uint8_t *len) {
  std::vector<uint32_t> keys, scratch;
  for (int s = 0; s < nsym; s++) if (freq[s]) keys.push_back((freq[s] << 9) | (uint32_t)s);
  std::sort(keys.begin(), keys.end());
  scratch.resize(keys.size() + 1);
  huff_lengths_from_sorted(keys.data(), (int)keys.size(), nsym, max_bits, len, scratch.data());
}
static uint64_t block_bits(uint32_t *fl, uint32_t *fd, uint64_t src_len) {
  fl[256]++;
  uint8_t ll[kNumLit], dl[kNumDist], cl[kNumClen], both[kNumLit + kNumDist];
  lengths_for(fl, kNumLit, 15, ll);
  lengths_for(fd, kNumDist, 15, dl);
  int hlit = kNumLit; while (hlit > 257 && !ll[hlit - 1]) hlit--;
  int hdist = kNumDist; while (hdist > 1 && !dl[hdist - 1]) hdist--;
  memcpy(both, ll, hlit); memcpy(both + hlit, dl, hdist);
  uint16_t rs[kNumLit + kNumDist]; uint32_t cf[kNumClen];
  rle_code_lengths(both, hlit + hdist, rs, cf);
  lengths_for(cf, kNumClen, 7, cl);
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int hclen = 19; while (hclen > 4 && !cl[order[hclen - 1]]) hclen--;
  uint64_t dyn = 17 + 3 * hclen, fix = 3;
  for (int s = 0; s < kNumClen; s++) dyn += (uint64_t)cf[s] * (cl[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0));
  for (int s = 0; s < kNumLit; s++) { uint32_t eb = s >= 257 ? len_extra_bits_of_sym(s) : 0; dyn += (uint64_t)fl[s] * (ll[s] + eb); fix += (uint64_t)fl[s] * (fixed_lit_len(s) + eb); }
  for (int s = 0; s < kNumDist; s++) { uint32_t eb = dist_extra_bits_of_sym(s); dyn += (uint64_t)fd[s] * (dl[s] + eb); fix += (uint64_t)fd[s] * (5 + eb); }
  uint64_t st = 3 + 7 + 32 + 8 * src_len;
  return std::min(st, std::min(dyn, fix));
}

struct Cfg { int S, D, R, nice, niceS; bool thresh; int closure = 0; };

static uint64_t encode(const uint8_t *src, uint32_t n, const Cfg &c) {
  Ring ring; Prev prev;
  std::vector<uint16_t> head(1u << kHashBits, 0), first(kTile, 0), mlen(kTile + 1, 0), mdist(kTile + 1, 0);
  std::vector<uint8_t> deep(kTile + 1, 0);
  uint32_t fl[kNumLit] = {0}, fd[kNumDist] = {0};
  uint64_t bits = 0, blk_src = 0;
  uint32_t pos = 0, kind = 0, carry_len = 0, carry_dist = 0, loaded = 0;
  int tiles = 0;
  for (uint32_t ts = 0; ts < n; ts += kTile) {
    uint32_t te = std::min(ts + (uint32_t)kTile, n), want = std::min(n, te + (uint32_t)kTile);
    for (; loaded < want; loaded++) ring.put(loaded, src[loaded]);
    for (uint32_t p = ts; p < te; p++) {
      if (p + 4 > n) { first[p - ts] = 0; continue; }
      uint32_t h = hash4(ring_load32(ring, p));
      first[p - ts] = head[h]; prev.l[p & (kWindow - 1)] = head[h]; head[h] = (uint16_t)p;
    }
    mlen[0] = carry_len; mdist[0] = carry_dist;
    std::fill(deep.begin(), deep.end(), 0);
    for (uint32_t p = ts; p < te; p++) {
      uint32_t d = 0, l = 0;
      if (p + 4 <= n) l = fm(ring, prev, p, n, first[p - ts], c.S, c.niceS, d);
      mlen[1 + p - ts] = l; mdist[1 + p - ts] = d;
    }
    for (int r = 0; r < c.R; r++) {
      // walk the parse with the current lengths, collect visited positions that are not deep yet
      std::vector<uint32_t> todo;
      uint32_t q = pos, k = kind;
      while (q < te) {
        uint32_t nk, em;
        if (!deep[1 + q - ts] && q + 4 <= n) { todo.push_back(q); }
        uint32_t np = lazy_next(q, k, mlen[1 + q - ts], mlen[q - ts], nk, em);
        q = np; k = nk;
      }
      if (todo.empty()) break;
      for (size_t ti = 0; ti < todo.size(); ti++) {
        uint32_t p = todo[ti];
        if (deep[1 + p - ts]) continue;
        uint32_t d = 0;
        // optionally only look for something longer than what the previous position already offers (the reference's prev_match_len)
        uint32_t mb = kMinMatch - 1;
        if (c.thresh) mb = std::max<uint32_t>(mb, mlen[1 + p - ts]);
        uint32_t l = fm(ring, prev, p, n, first[p - ts], c.D, c.nice, d, mb);
        if (l > mlen[1 + p - ts]) { mlen[1 + p - ts] = l; mdist[1 + p - ts] = d; }
        deep[1 + p - ts] = 1;
        if (c.closure && mlen[1 + p - ts] >= 4) {  // where a match taken at p lands, and its lazy lookahead
          uint32_t q = p + mlen[1 + p - ts];
          for (uint32_t k = 0; k < (uint32_t)c.closure; k++)
            if (q + k < te && q + k + 4 <= n && !deep[1 + q + k - ts]) todo.push_back(q + k);
        }
      }
    }
    while (pos < te) {
      uint32_t nk, em, ml = mlen[1 + pos - ts], mp = mlen[pos - ts];
      uint32_t np = lazy_next(pos, kind, ml, mp, nk, em);
      if (em == 1) { fl[src[pos]]++; blk_src++; }
      else if (em == 2) { fl[src[pos - 1]]++; blk_src++; }
      else if (em == 3) { uint32_t d = mdist[pos - ts], eb, ev; fl[len_sym_of(mp, eb, ev)]++; fd[dist_sym_of(d, eb, ev)]++; blk_src += mp; }
      pos = np; kind = nk;
    }
    carry_len = mlen[te - ts]; carry_dist = mdist[te - ts];
    tiles++;
    if (tiles == kTilesPerBlock || te == n) {
      bits += block_bits(fl, fd, blk_src);
      memset(fl, 0, sizeof fl); memset(fd, 0, sizeof fd); blk_src = 0; tiles = 0;
    }
  }
  return (bits + 7) / 8;
}

int main(int argc, char **argv) {
  int members = argc > 1 ? atoi(argv[1]) : 100;
  std::vector<std::vector<uint8_t>> data;
  uint64_t U = 0;
  // the C4 members: sizes 4096 + r % 258049 from rand-v1(seed 3), content text-v1(1000 + i)
  std::vector<uint64_t> raw(members);
  zipc_b200_synth_rand(3, raw.data(), 8 * members);
  for (int i = 0; i < members; i++) {
    size_t sz = 4096 + raw[i] % 258049;
    data.emplace_back(sz);
    zipc_b200_synth_text(1000 + i, data.back().data(), sz);
    U += sz;
  }
  std::vector<Cfg> cfgs = {
      {2, 16, 1, 128, 32, false, 0}, {2, 16, 1, 128, 32, false, 1}, {2, 16, 1, 128, 32, false, 2}, {2, 16, 2, 128, 32, false, 0},
      {2, 48, 1, 258, 32, false, 2}, {2, 48, 2, 258, 32, false, 0},
      {2, 128, 1, 258, 32, false, 1}, {2, 128, 1, 258, 32, false, 2}, {2, 128, 2, 258, 32, false, 0}, {2, 128, 2, 258, 32, false, 2},
      {2, 1024, 1, 258, 32, false, 2}, {2, 1024, 2, 258, 32, false, 2},
      {1, 4, 1, 32, 32, false, 2}, {1, 128, 1, 258, 32, false, 2},
  };


  for (const Cfg &c : cfgs) {
    g_steps = 0;
    uint64_t C = 0;
    for (auto &d : data) C += encode(d.data(), (uint32_t)d.size(), c);
    printf("S=%-3d D=%-4d R=%d nice=%-3d niceS=%-3d clo=%d  ratio %.4f  steps/byte %.2f\n", c.S, c.D, c.R, c.nice, c.niceS, c.closure, (double)C / U,
           (double)g_steps / U);
  }
  return 0;
}

// synth.cc -- the synthetic workload generators of SURVEY.md 8d (integer only, host side): text-v1, rand-v1.
// Plain C++ without any CUDA or library dependency: it is linked into libzipc_b200.so (zipc_b200_synth_*) and the
// oracle's Makefile compiles the same file into oracle/libzipc_synth.so, so that bench.py's reference arm can
// make the identical inputs without loading the product library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t &st) {
  uint64_t z = (st += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Language {
  std::vector<std::string> vocab;
  std::vector<uint64_t> cum;
  Language() : vocab(4096), cum(4096) {
    static const char letters[] = "etaoinshrdlcumwfgypbvkjxqz";
    uint64_t vs = 0x7a6970635f623230ull;
    for (auto &w : vocab) {
      size_t len = 2 + splitmix64(vs) % 9;
      for (size_t i = 0; i < len; i++) {
        uint64_t r1 = splitmix64(vs), r2 = splitmix64(vs);
        w.push_back(letters[(r1 % 26) * (r2 % 26) / 26]);
      }
    }
    uint64_t acc = 0;
    for (size_t k = 0; k < 4096; k++) { acc += 0x100000000ull / (k + 1); cum[k] = acc; }
  }
};

}  // namespace

extern "C" {

void zipc_b200_synth_rand(uint64_t seed, void *out, size_t n) {  // rand-v1: raw splitmix64, little endian
  uint8_t *o = static_cast<uint8_t *>(out);
  uint64_t st = seed;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v = splitmix64(st); std::memcpy(o + i, &v, 8); }
  if (i < n) { uint64_t v = splitmix64(st); std::memcpy(o + i, &v, n - i); }
}

// text-v1: Zipf-distributed words from a fixed 4096-word vocabulary, sentence punctuation, lines
// wrapped once the column passes 72.  The vocabulary does not depend on `seed` (all members share
// a language); the word sequence does.
void zipc_b200_synth_text(uint64_t seed, void *out, size_t n) {
  static const Language lang;  // thread-safe initialisation (callers generate members from a thread pool)
  const std::vector<std::string> &vocab = lang.vocab;
  const std::vector<uint64_t> &cum = lang.cum;
  uint8_t *o = static_cast<uint8_t *>(out);
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  size_t pos = 0, col = 0;
  while (pos < n) {
    uint64_t r = splitmix64(st) % cum.back();
    size_t k = (size_t)(std::upper_bound(cum.begin(), cum.end(), r) - cum.begin());
    const std::string &w = vocab[k];
    for (size_t i = 0; i < w.size() && pos < n; i++) o[pos++] = (uint8_t)w[i];
    col += w.size();
    uint64_t p = splitmix64(st) % 100;
    if (p < 6) { if (pos < n) o[pos++] = '.'; col++; }
    else if (p < 12) { if (pos < n) o[pos++] = ','; col++; }
    if (col > 72) { if (pos < n) o[pos++] = '\n'; col = 0; }
    else { if (pos < n) o[pos++] = ' '; col++; }
  }
}

}  // extern "C"

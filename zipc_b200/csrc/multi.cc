// multi.cc -- box-wide entry points: one call drives every selected GPU of the node.
//
// The reference is single-threaded and in-memory (src/zipc.mli:415-416); its unit of work on this path is a
// ZIP member (Zipc.File.deflate_of_binary_string / to_binary_string, src/zipc.ml:179-185, 205-225) or one
// string (Crc_32.string, src/zipc_deflate.ml:161-163).  Members are independent, so a batch is partitioned over
// the GPUs with no data-path collective: longest-processing-time-first on the bytes that
// dominate the work, one host thread + one zipc_b200_ctx (stream, arenas) per device, results gathered into
// the caller's arena; a single buffer is cut into contiguous slices whose CRC-32s are merged with
// x^(8 len) mod P on the host.  Pure orchestration over the single-device C ABI: no kernels here.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zipc_b200.h"

struct zipc_b200_mctx {
  std::vector<int> devices;
  std::vector<zipc_b200_ctx *> ctxs;
  // layout of the last batch call, for zipc_b200_multi_fetch
  std::vector<size_t> base, need;
  size_t total = 0;
  std::string last_error;
};

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }

// LPT: heaviest first onto the least loaded device.  Returns the member indices per device (input order kept
// inside a device, so outputs of one device stay in caller order).
std::vector<std::vector<uint32_t>> partition_lpt(const size_t *weight, size_t n, size_t g) {
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
  std::vector<uint64_t> load(g, 0);
  std::vector<std::vector<uint32_t>> parts(g);
  for (uint32_t i : order) {
    size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
    parts[d].push_back(i);
    load[d] += (uint64_t)weight[i] + 4096;  // a member costs a little even when it is tiny
  }
  for (auto &p : parts) std::sort(p.begin(), p.end());
  return parts;
}

template <class F>
void for_each_device(size_t g, F f) {
  std::vector<std::thread> th;
  th.reserve(g);
  for (size_t d = 1; d < g; d++) th.emplace_back([=] { f(d); });
  f(0);
  for (auto &t : th) t.join();
}

// Shared driver of the two codec batch calls.  run(d, idx, need, off, len, ck, st) executes the single-device call
// for the members idx on device d WITHOUT an arena (results stay in that device's ctx).
template <class Run>
int batch_multi(zipc_b200_mctx *m, size_t n, const size_t *weight, void *dst, size_t dst_cap, size_t *dst_need,
                size_t *dst_off, size_t *dst_len, uint32_t *checksum, int *status, Run run) {
  const size_t g = m->ctxs.size();
  auto parts = partition_lpt(weight, n, g);
  std::vector<int> rc(g, ZIPC_OK);
  m->need.assign(g, 0);
  std::vector<std::vector<size_t>> off(g), len(g);
  std::vector<std::vector<uint32_t>> ck(g);
  std::vector<std::vector<int>> st(g);
  for_each_device(g, [&](size_t d) {
    const size_t k = parts[d].size();
    off[d].assign(k, 0); len[d].assign(k, 0); ck[d].assign(k, 0); st[d].assign(k, 0);
    if (!k) return;
    int r = run(d, parts[d], &m->need[d], off[d].data(), len[d].data(), ck[d].data(), st[d].data());
    rc[d] = r == ZIPC_ERR_DST_TOO_SMALL ? ZIPC_OK : r;
  });
  for (size_t d = 0; d < g; d++)
    if (rc[d]) { m->last_error = std::string("device ") + std::to_string(m->devices[d]) + ": " + zipc_b200_last_error(m->ctxs[d]); return rc[d]; }
  m->base.assign(g, 0);
  size_t total = 0;
  for (size_t d = 0; d < g; d++) { m->base[d] = total; total += align_up(m->need[d], 16); }
  m->total = total;
  for (size_t d = 0; d < g; d++)
    for (size_t j = 0; j < parts[d].size(); j++) {
      const uint32_t i = parts[d][j];
      dst_off[i] = m->base[d] + off[d][j]; dst_len[i] = len[d][j]; status[i] = st[d][j];
      if (checksum) checksum[i] = ck[d][j];
    }
  if (dst_need) *dst_need = total;
  if (!dst || dst_cap < total) return ZIPC_ERR_DST_TOO_SMALL;
  return zipc_b200_multi_fetch(m, dst, dst_cap);
}

}  // namespace

extern "C" {

int zipc_b200_mctx_create(uint64_t device_mask, zipc_b200_mctx **out) {
  if (!out) return ZIPC_ERR_INVALID_ARG;
  *out = nullptr;
  const int nd = zipc_b200_device_count();
  if (nd <= 0) return ZIPC_ERR_NO_DEVICE;
  zipc_b200_mctx *m = new (std::nothrow) zipc_b200_mctx();
  if (!m) return ZIPC_ERR_NOMEM;
  for (int d = 0; d < nd && d < 64; d++)
    if (device_mask == 0 || (device_mask >> d & 1ull)) m->devices.push_back(d);
  if (m->devices.empty()) { delete m; return ZIPC_ERR_INVALID_ARG; }
  for (int d : m->devices) {
    zipc_b200_ctx *c = nullptr;
    if (int st = zipc_b200_ctx_create(d, &c)) { zipc_b200_mctx_destroy(m); return st; }
    m->ctxs.push_back(c);
  }
  *out = m;
  return ZIPC_OK;
}

void zipc_b200_mctx_destroy(zipc_b200_mctx *m) {
  if (!m) return;
  for (zipc_b200_ctx *c : m->ctxs) zipc_b200_ctx_destroy(c);
  delete m;
}

int zipc_b200_mctx_device_count(const zipc_b200_mctx *m) { return m ? (int)m->ctxs.size() : 0; }
zipc_b200_ctx *zipc_b200_mctx_ctx(zipc_b200_mctx *m, int k) { return (m && k >= 0 && (size_t)k < m->ctxs.size()) ? m->ctxs[k] : nullptr; }
const char *zipc_b200_mctx_last_error(const zipc_b200_mctx *m) { return m ? m->last_error.c_str() : ""; }

int zipc_b200_multi_crc32(zipc_b200_mctx *m, const void *src, size_t len, uint32_t *crc) {
  if (!m || !crc || (!src && len)) return ZIPC_ERR_INVALID_ARG;
  const size_t g = m->ctxs.size();
  // contiguous slices, 16-byte aligned cuts (G slices, G-1 host combines)
  std::vector<size_t> cut(g + 1, 0);
  for (size_t d = 1; d < g; d++) cut[d] = std::min(len, align_up(len / g * d, 16));
  cut[g] = len;
  std::vector<uint32_t> part(g, 0);
  std::vector<int> rc(g, ZIPC_OK);
  for_each_device(g, [&](size_t d) {
    if (cut[d + 1] == cut[d] && d) return;
    rc[d] = zipc_b200_crc32(m->ctxs[d], static_cast<const uint8_t *>(src) + cut[d], cut[d + 1] - cut[d], &part[d]);
  });
  for (size_t d = 0; d < g; d++) if (rc[d]) { m->last_error = zipc_b200_last_error(m->ctxs[d]); return rc[d]; }
  uint32_t c = part[0];
  for (size_t d = 1; d < g; d++)
    if (cut[d + 1] > cut[d]) c = zipc_b200_crc32_combine(c, part[d], cut[d + 1] - cut[d]);
  *crc = c;
  return ZIPC_OK;
}

int zipc_b200_multi_inflate_batch(zipc_b200_mctx *m, int ck, int adler_mode, size_t n, const void *const *src,
                                  const size_t *src_len, const size_t *max_out, void *dst, size_t dst_cap,
                                  size_t *dst_need, size_t *dst_off, size_t *dst_len, uint32_t *checksum, int *status) {
  if (!m || (n && (!src || !src_len || !dst_off || !dst_len || !status))) return ZIPC_ERR_INVALID_ARG;
  if (dst_need) *dst_need = 0;
  if (!n) return ZIPC_OK;
  return batch_multi(m, n, src_len, dst, dst_cap, dst_need, dst_off, dst_len, checksum, status,
                     [&](size_t d, const std::vector<uint32_t> &idx, size_t *need, size_t *off, size_t *len, uint32_t *c, int *st) {
                       const size_t k = idx.size();
                       std::vector<const void *> p(k);
                       std::vector<size_t> l(k), mo(k);
                       for (size_t j = 0; j < k; j++) { p[j] = src[idx[j]]; l[j] = src_len[idx[j]]; mo[j] = max_out ? max_out[idx[j]] : ZIPC_SIZE_UNKNOWN; }
                       return zipc_b200_inflate_batch(m->ctxs[d], ck, adler_mode, k, p.data(), l.data(), mo.data(), nullptr, 0, need, off, len, c, st);
                     });
}

int zipc_b200_multi_deflate_batch(zipc_b200_mctx *m, int level, int ck, int adler_mode, size_t n, const void *const *src,
                                  const size_t *src_len, void *dst, size_t dst_cap, size_t *dst_need, size_t *dst_off,
                                  size_t *dst_len, uint32_t *checksum, int *status) {
  if (!m || (n && (!src || !src_len || !dst_off || !dst_len || !status))) return ZIPC_ERR_INVALID_ARG;
  if (dst_need) *dst_need = 0;
  if (!n) return ZIPC_OK;
  return batch_multi(m, n, src_len, dst, dst_cap, dst_need, dst_off, dst_len, checksum, status,
                     [&](size_t d, const std::vector<uint32_t> &idx, size_t *need, size_t *off, size_t *len, uint32_t *c, int *st) {
                       const size_t k = idx.size();
                       std::vector<const void *> p(k);
                       std::vector<size_t> l(k);
                       for (size_t j = 0; j < k; j++) { p[j] = src[idx[j]]; l[j] = src_len[idx[j]]; }
                       return zipc_b200_deflate_batch(m->ctxs[d], level, ck, adler_mode, k, p.data(), l.data(), nullptr, 0, need, off, len, c, st);
                     });
}

int zipc_b200_multi_fetch(zipc_b200_mctx *m, void *dst, size_t dst_cap) {
  if (!m || !dst) return ZIPC_ERR_INVALID_ARG;
  if (dst_cap < m->total) return ZIPC_ERR_DST_TOO_SMALL;
  const size_t g = m->ctxs.size();
  std::vector<int> rc(g, ZIPC_OK);
  for_each_device(g, [&](size_t d) {
    if (d < m->need.size() && m->need[d]) rc[d] = zipc_b200_fetch(m->ctxs[d], static_cast<uint8_t *>(dst) + m->base[d], m->need[d]);
  });
  for (size_t d = 0; d < g; d++) if (rc[d]) { m->last_error = zipc_b200_last_error(m->ctxs[d]); return rc[d]; }
  return ZIPC_OK;
}

}  // extern "C"

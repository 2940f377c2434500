// multi.cc -- box-wide entry points: one call drives every selected GPU of the node.
//
// The reference is single-threaded and in-memory (src/zipc.mli:415-416); its unit of work on this path is a
// ZIP member (Zipc.File.deflate_of_binary_string / to_binary_string, src/zipc.ml:179-185, 205-225) or one
// string (Crc_32.string, src/zipc_deflate.ml:161-163).  Members are independent, so a batch is partitioned over
// the GPUs with no data-path collective: contiguous runs of equal byte sums (or longest-first when
// there are only a few members), one host thread + one zipc_b200_ctx (stream, arenas) per device, results gathered into
// the caller's arena; a single buffer is cut into contiguous slices whose CRC-32s are merged with
// x^(8 len) mod P on the host.  Pure orchestration over the single-device C ABI: no kernels here.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct zipc_b200_mctx {
  std::vector<int> devices;
  std::vector<zipc_b200_ctx *> ctxs;
  bool pipelined = false;   // sub-contexts of ONE device: uploads take turns (zb::UploadGate)
  zb::UploadGate gate;
  // layout of the last batch call, for zipc_b200_multi_fetch
  std::vector<size_t> base, need;
  size_t total = 0;
  std::string last_error;
};

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }

// Members onto devices.  Many members: contiguous runs of the caller's order with equal byte sums -- members that lie
// next to each other in host memory (an in-memory archive) then travel to their device as ONE span; a device balances
// its own members over its SMs anyway.  Few members: longest-processing-time first (heaviest onto the least loaded
// device), because one large member can outweigh a whole run.  Returns the member indices per device, ascending.
std::vector<std::vector<uint32_t>> partition_members(const size_t *weight, size_t n, size_t g) {
  std::vector<std::vector<uint32_t>> parts(g);
  if (n >= 64 * g) {
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += (uint64_t)weight[i] + 4096;  // a member costs a little even when it is tiny
    uint64_t run = 0;
    size_t d = 0;
    for (size_t i = 0; i < n; i++) {
      while (d + 1 < g && run >= total * (d + 1) / g) d++;
      parts[d].push_back((uint32_t)i);
      run += (uint64_t)weight[i] + 4096;
    }
    return parts;
  }
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
  std::vector<uint64_t> load(g, 0);
  for (uint32_t i : order) {
    size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
    parts[d].push_back(i);
    load[d] += (uint64_t)weight[i] + 4096;
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

// Shared driver of the two codec batch calls.  run(d, idx, part_dst, part_cap, need, off, len, ck, st) executes the
// single-device call for the members idx on context d.  Output sizes unknown (deflate; inflate without sizes): the call
// runs WITHOUT an arena (results stay in that context); each thread then waits until the contexts before it know how much
// they produced, which fixes its place in the caller's arena, and fetches its part -- while later contexts are still
// computing.  Output sizes known (`out_size`: inflate with ?decompressed_size): every part's place is known beforehand and
// the single-device call writes there itself (and may download progressively, api.cu plan_arena).
template <class Run>
int batch_multi(zipc_b200_mctx *m, size_t n, const size_t *weight, void *dst, size_t dst_cap, size_t *dst_need,
                size_t *dst_off, size_t *dst_len, uint32_t *checksum, int *status, Run run, const size_t *out_size = nullptr,
                size_t align = 16) {
  size_t g = m->ctxs.size();
  if (m->pipelined) {  // every group gets members (the upload tickets need that) and enough bytes to be worth its fixed costs
    uint64_t bytes = 0;
    for (size_t i = 0; i < n; i++) bytes += weight[i];
    g = std::max<size_t>(1, std::min({g, n / 256, (size_t)(bytes / (32u << 20))}));
  }
  auto parts = partition_members(weight, n, g);
  std::vector<size_t> part_base(g, 0), part_size(g, 0);
  bool placed = false;
  if (out_size && dst && !m->pipelined) {
    size_t t = 0;
    for (size_t d = 0; d < g; d++) {
      part_base[d] = t;
      for (uint32_t i : parts[d]) part_size[d] += align_up(out_size[i], 16);
      t += part_size[d];
    }
    placed = t <= dst_cap;
  }
  std::vector<int> rc(g, ZIPC_OK);
  m->need.assign(m->ctxs.size(), 0);
  m->base.assign(m->ctxs.size(), 0);
  std::vector<std::vector<size_t>> off(g), len(g);
  std::vector<std::vector<uint32_t>> ck(g);
  std::vector<std::vector<int>> st(g);
  std::mutex mu;
  std::condition_variable cv;
  std::vector<char> known(g, 0);
  if (m->pipelined) { std::lock_guard<std::mutex> lk(m->gate.m); m->gate.turn = 0; m->gate.launched.assign(m->gate.done.size(), 0); }
  for_each_device(g, [&](size_t d) {
    const size_t k = parts[d].size();
    off[d].assign(k, 0); len[d].assign(k, 0); ck[d].assign(k, 0); st[d].assign(k, 0);
    zipc_b200_ctx *c = m->ctxs[d];
    if (m->pipelined) { c->gate = &m->gate; c->gate_ticket = d; c->gate_passed = false; }
    int r = ZIPC_OK;
    if (k) r = run(d, parts[d], placed ? static_cast<uint8_t *>(dst) + part_base[d] : nullptr, placed ? part_size[d] : 0, &m->need[d],
                   off[d].data(), len[d].data(), ck[d].data(), st[d].data());
    if (m->pipelined && !c->gate_passed) {  // the call never got to its upload: do not hold up the groups behind it
      std::unique_lock<std::mutex> lk(m->gate.m);
      m->gate.cv.wait(lk, [&] { return m->gate.turn == d; });
      m->gate.turn++;
      lk.unlock();
      m->gate.cv.notify_all();
    }
    if (m->pipelined && d < m->gate.launched.size()) {  // a call that never launched must not hold up the groups behind it either
      { std::lock_guard<std::mutex> lk(m->gate.m); m->gate.launched[d] = 1; }
      m->gate.cv.notify_all();
    }
    c->gate = nullptr;
    rc[d] = r == ZIPC_ERR_DST_TOO_SMALL ? ZIPC_OK : r;
    if (placed) { m->need[d] = part_size[d]; m->base[d] = part_base[d]; return; }
    size_t base = 0;
    {
      std::unique_lock<std::mutex> lk(mu);
      known[d] = 1;
      cv.notify_all();
      cv.wait(lk, [&] { for (size_t j = 0; j < d; j++) if (!known[j]) return false; return true; });
      for (size_t j = 0; j < d; j++) base += align_up(m->need[j], align);
    }
    m->base[d] = base;
    if (rc[d] == ZIPC_OK && dst && m->need[d] && base + m->need[d] <= dst_cap)
      rc[d] = zipc_b200_fetch(c, static_cast<uint8_t *>(dst) + base, m->need[d]);
    zb::pipe_mark(c, "fetched");
  });
  for (size_t d = 0; d < g; d++)
    if (rc[d]) { m->last_error = std::string("context ") + std::to_string(d) + ": " + zipc_b200_last_error(m->ctxs[d]); return rc[d]; }
  size_t total = 0;
  for (size_t d = 0; d < g; d++) total += align_up(m->need[d], align);
  m->total = total;
  for (size_t d = 0; d < g; d++)
    for (size_t j = 0; j < parts[d].size(); j++) {
      const uint32_t i = parts[d][j];
      dst_off[i] = m->base[d] + off[d][j]; dst_len[i] = len[d][j]; status[i] = st[d][j];
      if (checksum) checksum[i] = ck[d][j];
    }
  if (dst_need) *dst_need = total;
  if (!dst || dst_cap < total) return ZIPC_ERR_DST_TOO_SMALL;
  return ZIPC_OK;
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
  if (!m->gate.done.empty() && !m->devices.empty()) {
    cudaSetDevice(m->devices[0]);
    for (cudaEvent_t e : m->gate.done) cudaEventDestroy(e);
  }
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
  bool all_known = max_out != nullptr;
  for (size_t i = 0; i < n && all_known; i++) all_known = max_out[i] != ZIPC_SIZE_UNKNOWN;
  return batch_multi(m, n, src_len, dst, dst_cap, dst_need, dst_off, dst_len, checksum, status,
                     [&](size_t d, const std::vector<uint32_t> &idx, void *pdst, size_t pcap, size_t *need, size_t *off, size_t *len, uint32_t *c, int *st) {
                       const size_t k = idx.size();
                       std::vector<const void *> p(k);
                       std::vector<size_t> l(k), mo(k);
                       for (size_t j = 0; j < k; j++) { p[j] = src[idx[j]]; l[j] = src_len[idx[j]]; mo[j] = max_out ? max_out[idx[j]] : ZIPC_SIZE_UNKNOWN; }
                       return zipc_b200_inflate_batch(m->ctxs[d], ck, adler_mode, k, p.data(), l.data(), mo.data(), pdst, pcap, need, off, len, c, st);
                     },
                     all_known ? max_out : nullptr);
}

int zipc_b200_multi_deflate_batch(zipc_b200_mctx *m, int level, int ck, int adler_mode, size_t n, const void *const *src,
                                  const size_t *src_len, void *dst, size_t dst_cap, size_t *dst_need, size_t *dst_off,
                                  size_t *dst_len, uint32_t *checksum, int *status) {
  if (!m || (n && (!src || !src_len || !dst_off || !dst_len || !status))) return ZIPC_ERR_INVALID_ARG;
  if (dst_need) *dst_need = 0;
  if (!n) return ZIPC_OK;
  return batch_multi(m, n, src_len, dst, dst_cap, dst_need, dst_off, dst_len, checksum, status,
                     [&](size_t d, const std::vector<uint32_t> &idx, void *, size_t, size_t *need, size_t *off, size_t *len, uint32_t *c, int *st) {
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

namespace zb {

int pipeline_create(int device, int depth, zipc_b200_mctx **out) {
  *out = nullptr;
  zipc_b200_mctx *m = new (std::nothrow) zipc_b200_mctx();
  if (!m) return ZIPC_ERR_NOMEM;
  m->pipelined = true;
  for (int k = 0; k < depth; k++) {
    zipc_b200_ctx *c = nullptr;
    if (int st = zipc_b200_ctx_create(device, &c)) { zipc_b200_mctx_destroy(m); return st; }
    c->is_sub = true;
    m->devices.push_back(device);
    m->ctxs.push_back(c);
  }
  cudaSetDevice(device);
  for (int k = 0; k < depth; k++) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { zipc_b200_mctx_destroy(m); return ZIPC_ERR_CUDA; }
    m->gate.done.push_back(e);
  }
  m->gate.launched.assign(depth, 0);
  *out = m;
  return ZIPC_OK;
}

// Payloads of a ZIP archive through the pipeline: like zipc_b200_multi_deflate_batch, but every output is preceded by
// gap[i] free bytes (its local file header) and nothing is padded, so dst_off[] are the payloads' offsets in the archive.
int pipeline_deflate_gapped(zipc_b200_mctx *m, int level, size_t n, const void *const *src, const size_t *src_len, const uint32_t *gap,
                            void *dst, size_t dst_cap, size_t *dst_need, size_t *dst_off, size_t *dst_len, uint32_t *crc, int *status) {
  return batch_multi(m, n, src_len, dst, dst_cap, dst_need, dst_off, dst_len, crc, status,
                     [&](size_t d, const std::vector<uint32_t> &idx, void *, size_t, size_t *need, size_t *off, size_t *len, uint32_t *c, int *st) {
                       const size_t k = idx.size();
                       std::vector<const void *> p(k);
                       std::vector<size_t> l(k);
                       std::vector<uint32_t> gp(k);
                       for (size_t j = 0; j < k; j++) { p[j] = src[idx[j]]; l[j] = src_len[idx[j]]; gp[j] = gap[idx[j]]; }
                       return deflate_batch_layout(m->ctxs[d], level, ZIPC_CK_CRC32, 0, k, p.data(), l.data(), nullptr, 0, need, off, len, c, st, gp.data(), 1);
                     },
                     nullptr, 1);
}

uint64_t pipeline_launches(const zipc_b200_mctx *m) {
  uint64_t n = 0;
  for (const zipc_b200_ctx *c : m->ctxs) n += c->launches;
  return n;
}

}  // namespace zb

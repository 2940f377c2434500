// de_yardstick.cu -- a YARDSTICK, not a product path: Blackwell's hardware decompression engine on the same deflate
// members (cuMemBatchDecompressAsync, CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE; CUDA 12.8+ driver API).  It is the only
// on-box competitor to inflate_kernel; bench.py reports its throughput as de_yardstick_GBps next to ours, or says
// that the box does not support it.  Nothing in zipc_b200/ calls this.
//
// Build: tools/build_tools.sh  ->  tools/libde_yardstick.so  (driver entry points are fetched at run time, no -lcuda)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {
template <class F>
bool entry(const char *name, F *fn) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}
}  // namespace

// Decompresses n raw-deflate members (host pointers) `reps` times; *ms_out = average milliseconds per batch (device
// resident, CUDA events), *n_ok = members whose output length and first/last bytes matched `expect` on the last rep.
// Returns 0 on success, 1 = unsupported on this box (msg says why), 2 = failure.
extern "C" int de_yardstick(int device, size_t n, const void *const *src, const size_t *src_len, const void *const *expect,
                            const size_t *dst_len, int reps, double *ms_out, size_t *n_ok, char *msg, size_t msg_cap) {
  auto say = [&](const char *m) { if (msg && msg_cap) { std::snprintf(msg, msg_cap, "%s", m); } };
  if (cudaSetDevice(device) != cudaSuccess) { say("cudaSetDevice failed"); return 2; }
  cudaFree(0);
  CUresult (*getattr)(int *, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*batch)(CUmemDecompressParams *, size_t, unsigned int, size_t *, CUstream) = nullptr;
  if (!entry("cuDeviceGetAttribute", &getattr) || !entry("cuMemBatchDecompressAsync", &batch)) {
    say("driver has no cuMemBatchDecompressAsync");
    return 1;
  }
  int mask = 0, maxlen = 0;
  if (getattr(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, device) != CUDA_SUCCESS) mask = 0;
  getattr(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, device);
  if (!(mask & CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE)) {
    char b[160];
    std::snprintf(b, sizeof b, "CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK = %#x: deflate not supported by the decompression engine of this device/driver", mask);
    say(b);
    return 1;
  }
  size_t ctot = 0, utot = 0;
  std::vector<size_t> coff(n), uoff(n);
  for (size_t i = 0; i < n; i++) {
    coff[i] = ctot; uoff[i] = utot;
    ctot += (src_len[i] + 63) & ~(size_t)63; utot += (dst_len[i] + 63) & ~(size_t)63;
    if (maxlen > 0 && dst_len[i] > (size_t)maxlen) { say("a member exceeds CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH"); return 1; }
  }
  uint8_t *dc = nullptr, *du = nullptr;
  uint32_t *dact = nullptr;
  if (cudaMalloc(&dc, ctot + 64) != cudaSuccess || cudaMalloc(&du, utot + 64) != cudaSuccess || cudaMalloc(&dact, n * 4 + 64) != cudaSuccess) {
    say("cudaMalloc failed"); cudaFree(dc); cudaFree(du); cudaFree(dact); return 2;
  }
  for (size_t i = 0; i < n; i++) cudaMemcpy(dc + coff[i], src[i], src_len[i], cudaMemcpyHostToDevice);
  std::vector<CUmemDecompressParams> ps(n);
  for (size_t i = 0; i < n; i++) {
    std::memset(&ps[i], 0, sizeof ps[i]);
    ps[i].srcNumBytes = src_len[i]; ps[i].dstNumBytes = dst_len[i]; ps[i].dstActBytes = dact + i;
    ps[i].src = dc + coff[i]; ps[i].dst = du + uoff[i]; ps[i].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
  }
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  size_t bad = (size_t)-1;
  for (int w = 0; w < 2 && !rc; w++) {
    CUresult r = batch(ps.data(), n, 0, &bad, (CUstream)st);
    if (r != CUDA_SUCCESS) { char b[128]; std::snprintf(b, sizeof b, "cuMemBatchDecompressAsync failed: CUresult %d at index %zu", (int)r, bad); say(b); rc = r == CUDA_ERROR_NOT_SUPPORTED ? 1 : 2; }
  }
  if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { say(cudaGetErrorString(cudaGetLastError())); rc = 2; }
  if (!rc) {
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; r++) batch(ps.data(), n, 0, &bad, (CUstream)st);
    cudaEventRecord(e1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { say(cudaGetErrorString(cudaGetLastError())); rc = 2; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / (reps > 0 ? reps : 1);
  }
  if (!rc) {
    std::vector<uint32_t> act(n);
    cudaMemcpy(act.data(), dact, n * 4, cudaMemcpyDeviceToHost);
    size_t ok = 0;
    std::vector<uint8_t> head(64), tail(64);
    for (size_t i = 0; i < n; i++) {
      if (act[i] != dst_len[i]) continue;
      size_t k = dst_len[i] < 64 ? dst_len[i] : 64;
      cudaMemcpy(head.data(), du + uoff[i], k, cudaMemcpyDeviceToHost);
      cudaMemcpy(tail.data(), du + uoff[i] + dst_len[i] - k, k, cudaMemcpyDeviceToHost);
      const uint8_t *e = static_cast<const uint8_t *>(expect[i]);
      if (!std::memcmp(head.data(), e, k) && !std::memcmp(tail.data(), e + dst_len[i] - k, k)) ok++;
    }
    *n_ok = ok;
    say("ok");
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
  cudaFree(dc); cudaFree(du); cudaFree(dact);
  return rc;
}

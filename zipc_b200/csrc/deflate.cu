// deflate.cu -- placeholder until the encoder kernels land (next milestone).
#include "common.cuh"
extern "C" {
size_t zipc_b200_deflate_bound(size_t src_len) { return src_len + 5 * (src_len / 65534 + 1) + 64; }
int zipc_b200_deflate_batch(zipc_b200_ctx *, int, int, int, size_t, const void *const *, const size_t *, void *, size_t,
                            size_t *, size_t *, size_t *, uint32_t *, int *) { return ZIPC_ERR_INVALID_ARG; }
int zipc_b200_deflate_batch_dev(zipc_b200_ctx *, int, int, int, size_t, const void *, const size_t *, const size_t *,
                                void *, const size_t *, const size_t *, size_t *, uint32_t *, int *) { return ZIPC_ERR_INVALID_ARG; }
int zipc_b200_zlib_compress_batch(zipc_b200_ctx *, int, int, size_t, const void *const *, const size_t *, void *, size_t,
                                  size_t *, size_t *, size_t *, uint32_t *, int *) { return ZIPC_ERR_INVALID_ARG; }
}

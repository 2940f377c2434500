// zip_api.cu -- GPU-backed archive entry points (placeholder: extract/deflate-archive land with deflate).
#include "common.cuh"
extern "C" {
int zipc_b200_zip_extract_batch(zipc_b200_ctx *, const zipc_b200_member *, size_t, void *, size_t, size_t *, size_t *,
                                size_t *, uint32_t *, int *) { return ZIPC_ERR_INVALID_ARG; }
int zipc_b200_zip_deflate_archive(zipc_b200_ctx *, int, size_t, const char *const *, const uint32_t *, const void *const *,
                                  const size_t *, const int32_t *, const int64_t *, const char *, void *, size_t, size_t *) {
  return ZIPC_ERR_INVALID_ARG;
}
}

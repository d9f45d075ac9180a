// Tensor-core (tcgen05) dissimilarity map -- placeholder until the GEMM kernels land.
#include "common.cuh"

namespace cas {

size_t whd_gemm_workspace_bytes(int64_t, int32_t, int32_t, int32_t) { return 256; }

int whd_gemm_map(const int32_t *, int64_t, int32_t, int32_t, const double *, int32_t, void *,
                 int64_t, int32_t, int64_t, int64_t, void *, size_t, cudaStream_t) {
  set_error("CAS_MAP_TENSOR is not built into this library yet");
  return CAS_ERR_UNSUPPORTED;
}

}  // namespace cas

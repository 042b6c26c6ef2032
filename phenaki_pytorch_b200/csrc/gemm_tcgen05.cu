// tcgen05 / TMEM / TMA bf16 GEMM -- placeholder until the tensor-core path lands (returns UNSUPPORTED,
// loudly; nothing falls back to another implementation).
#include "phk_common.cuh"
using namespace phk;
extern "C" int phk_gemm_bf16(const void*, int64_t, const void*, int64_t, void*, int64_t, int64_t, int32_t, int32_t,
                             const float*, const float*, int64_t, int64_t, int64_t, int32_t, phk_stream_t) {
  PHK_REQUIRE(false, PHK_E_UNSUPPORTED, "phk_gemm_bf16: tcgen05 path not built yet");
  return 0;
}

// bf16 tensor-core path (placeholder until the tcgen05 kernels land: reports "unsupported" loudly).
#include "bf16_path.cuh"
namespace aaconv {
size_t bf16_saved_bytes(const Dims&) { return 256; }
size_t bf16_scratch_bytes(const Dims&) { return 256; }
int64_t bf16_saved_offset(const Dims&, const char*) { return -1; }
int bf16_forward(const Dims&, const float*, const aaconv_params*, float*, float*, void*, void*, cudaStream_t) {
  return fail(AACONV_E_UNSUPPORTED, "bf16 path not built yet");
}
int bf16_backward(const Dims&, const float*, const aaconv_params*, const float*, const void*, void*, float*,
                  const aaconv_param_grads*, cudaStream_t) {
  return fail(AACONV_E_UNSUPPORTED, "bf16 path not built yet");
}
}  // namespace aaconv

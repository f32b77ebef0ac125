// extern "C" surface of libaaconv_b200.so (see include/aaconv_b200.h).
#include "fp32_path.cuh"
#include "bf16_path.cuh"

namespace aaconv {
int bce_launch(const float* z, const float* targets, int ld, const int32_t* cols, int B, int C, float* el,
               float* loss, float* dz, const float* grad_scale, cudaStream_t st);

int ensemble_mean_launch(const float* logits, int M, size_t n, float* mean, cudaStream_t st);
int auroc_launch(const float* z, const float* t, int N, int C, float* auroc, void* workspace, cudaStream_t st);

static int validate(const aaconv_dims* dd, int precision) {
  if (!dd) return fail(AACONV_E_ARG, "dims is NULL");
  const aaconv_dims& d = *dd;
  if (precision != AACONV_FP32 && precision != AACONV_BF16) return fail(AACONV_E_ARG, "unknown precision %d", precision);
  if (d.B <= 0 || d.Cin <= 0 || d.Hin <= 0 || d.Win <= 0 || d.Cout <= 0 || d.nh <= 0 || d.dk <= 0 || d.dv <= 0)
    return fail(AACONV_E_ARG, "non-positive dimension");
  if (d.dk % d.nh) return fail(AACONV_E_ARG, "nh must divide dk");       // attn_aug_conv.py:27
  if (d.dv % d.nh) return fail(AACONV_E_ARG, "nh must divide dv");       // attn_aug_conv.py:28
  if (d.stride <= 0 || d.ksize <= 0 || d.dil <= 0 || d.pad < 0) return fail(AACONV_E_ARG, "bad conv geometry");
  if (d.ksize > 8) return fail(AACONV_E_UNSUPPORTED, "kernel_size > 8");
  const int H = (d.Hin - 1) / d.stride + 1, W = (d.Win - 1) / d.stride + 1;   // 1x1 strided projection
  if (H != d.H || W != d.W)
    return fail(AACONV_E_ARG, "input_dims (%d,%d) do not match the strided map (%d,%d)", d.H, d.W, H, W);
  if (d.Cout > d.dv) {
    const int Hc = (d.Hin + 2 * d.pad - d.dil * (d.ksize - 1) - 1) / d.stride + 1;
    const int Wc = (d.Win + 2 * d.pad - d.dil * (d.ksize - 1) - 1) / d.stride + 1;
    if (Hc != H || Wc != W)
      return fail(AACONV_E_ARG, "conv branch map (%d,%d) != attention map (%d,%d): cannot concatenate", Hc, Wc, H, W);
  }
  // bf16 mode has no fallback to the fp32 kernels: report WHY a shape is outside the tensor-core kernels here, before any
  // buffer is sized (the *_bytes() queries return 0 for an invalid call)
  if (precision == AACONV_BF16) AACONV_TRY(aug_supported(Dims(d)));
  return 0;
}
static int validate_io(const aaconv_dims* d, const aaconv_io* io) {
  if (!io) return 0;
  if ((io->x_dtype != AACONV_FP32 && io->x_dtype != AACONV_BF16) || (io->y_dtype != AACONV_FP32 && io->y_dtype != AACONV_BF16))
    return fail(AACONV_E_ARG, "aaconv_io: element types must be AACONV_FP32 or AACONV_BF16");
  if (io->y_batch_stride != 0 && io->y_batch_stride < (int64_t)d->Cout * d->H * d->W)
    return fail(AACONV_E_ARG, "aaconv_io: y_batch_stride %lld is smaller than one sample of y (%lld)", (long long)io->y_batch_stride,
                (long long)d->Cout * d->H * d->W);
  return 0;
}
}  // namespace aaconv

using namespace aaconv;

extern "C" {

int aaconv_abi_version(void) { return AACONV_ABI_VERSION; }
const char* aaconv_last_error(void) { return last_error_ref().c_str(); }
int aaconv_validate(const aaconv_dims* d, int precision) { return validate(d, precision); }

size_t aaconv_saved_bytes_io(const aaconv_dims* d, int precision, const aaconv_io* io) {
  if (validate(d, precision) || validate_io(d, io)) return 0;
  return precision == AACONV_FP32 ? f32_saved_bytes_io(Dims(*d, io)) : bf16_saved_bytes(Dims(*d, io));
}
size_t aaconv_scratch_bytes_io(const aaconv_dims* d, int precision, const aaconv_io* io) {
  if (validate(d, precision) || validate_io(d, io)) return 0;
  return precision == AACONV_FP32 ? f32_scratch_bytes(Dims(*d, io)) : bf16_scratch_bytes(Dims(*d, io), 0);
}
size_t aaconv_saved_bytes(const aaconv_dims* d, int precision) { return aaconv_saved_bytes_io(d, precision, nullptr); }
size_t aaconv_scratch_bytes(const aaconv_dims* d, int precision, int) { return aaconv_scratch_bytes_io(d, precision, nullptr); }
int64_t aaconv_saved_offset(const aaconv_dims* d, int precision, const char* name) {
  if (validate(d, precision) || !name) return -1;
  return precision == AACONV_FP32 ? f32_saved_offset(Dims(*d), name) : bf16_saved_offset(Dims(*d), name);
}

int aaconv_forward_io(const aaconv_dims* d, int precision, const aaconv_io* io, const void* x, const aaconv_params* p, void* y,
                      float* weights, void* saved, void* scratch, void* stream) {
  AACONV_TRY(validate(d, precision));
  AACONV_TRY(validate_io(d, io));
  if (!x || !p || !y || !saved || !scratch) return fail(AACONV_E_ARG, "NULL buffer");
  if (!p->qkv_w || !p->out_w || (d->Cout > d->dv && !p->conv_w) || (d->relative && (!p->key_rel_h || !p->key_rel_w)))
    return fail(AACONV_E_ARG, "NULL parameter");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return precision == AACONV_FP32 ? f32_forward(Dims(*d, io), x, p, y, weights, saved, scratch, st)
                                  : bf16_forward(Dims(*d, io), x, p, y, weights, saved, scratch, st);
}
int aaconv_forward(const aaconv_dims* d, int precision, const float* x, const aaconv_params* p, float* y,
                   float* weights, void* saved, void* scratch, void* stream) {
  return aaconv_forward_io(d, precision, nullptr, x, p, y, weights, saved, scratch, stream);
}

int aaconv_backward_io(const aaconv_dims* d, int precision, const aaconv_io* io, const void* x, const aaconv_params* p, const float* dy,
                       void* saved, void* scratch, void* dx, const aaconv_param_grads* g, void* stream) {
  AACONV_TRY(validate(d, precision));
  AACONV_TRY(validate_io(d, io));
  if (!x || !p || !dy || !saved || !scratch || !g) return fail(AACONV_E_ARG, "NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return precision == AACONV_FP32 ? f32_backward(Dims(*d, io), x, p, dy, saved, scratch, dx, g, st)
                                  : bf16_backward(Dims(*d, io), x, p, dy, saved, scratch, dx, g, st);
}
int aaconv_backward(const aaconv_dims* d, int precision, const float* x, const aaconv_params* p, const float* dy,
                    void* saved, void* scratch, float* dx, const aaconv_param_grads* g, void* stream) {
  return aaconv_backward_io(d, precision, nullptr, x, p, dy, saved, scratch, dx, g, stream);
}

int aaconv_bce_forward_backward(const float* z, const float* targets, int ld, const int32_t* cols, int B, int C,
                                float* element_loss, float* loss, float* dz, const float* grad_scale, void* stream) {
  if (!z || !targets || B <= 0 || C <= 0 || ld < (cols ? 1 : C)) return fail(AACONV_E_ARG, "bad BCE arguments");
  return bce_launch(z, targets, ld, cols, B, C, element_loss, loss, dz, grad_scale, static_cast<cudaStream_t>(stream));
}

int aaconv_ensemble_mean(const float* logits, int n_models, int N, int C, float* mean, void* stream) {
  if (!logits || !mean || n_models <= 0 || N <= 0 || C <= 0) return fail(AACONV_E_ARG, "bad ensemble_mean arguments");
  return ensemble_mean_launch(logits, n_models, (size_t)N * C, mean, static_cast<cudaStream_t>(stream));
}

size_t aaconv_auroc_workspace_bytes(int C) { return C > 0 ? align256(sizeof(unsigned long long) * 3 * (size_t)C) : 0; }

int aaconv_auroc(const float* logits, const float* targets, int N, int C, float* auroc, void* workspace, void* stream) {
  if (!logits || !targets || !auroc || !workspace || N <= 0 || C <= 0 || C > 1024) return fail(AACONV_E_ARG, "bad auroc arguments");
  return auroc_launch(logits, targets, N, C, auroc, workspace, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

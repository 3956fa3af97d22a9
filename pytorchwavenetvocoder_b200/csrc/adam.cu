// adam.cu -- wnb_adam_flat: the optimizer step of the training loop (reference bin/train.py:457-460 torch.optim.Adam,
// :539 optimizer.step()) as ONE streaming kernel over flat buffers.
//
// The backward already leaves every parameter gradient in one contiguous buffer (pack.cu, `module._wnb_flat_grad`); with
// the parameters and both moment estimates laid out the same way (pytorchwavenetvocoder_b200/optim.py) the whole update
// is 7 floats of traffic per parameter in one launch -- torch's multi-tensor fused Adam over the 184 reference-shaped
// tensors of the 30-block model is 11 launches / 170 us, this is one launch at HBM speed.
// Arithmetic follows torch's fused Adam (fp32): g += wd*p; m = m + (1-b1)(g - m); v = b2 v + (1-b2) g^2;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps), bias corrections bc1 = 1 - b1^t, bc2 = 1 - b2^t from the host.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/wnb200.h"
#include "common.cuh"

namespace wnb {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step_size, float omb1, float b2, float omb2,
                                         float eps, float wd, float bc2_sqrt) {
  if (wd != 0.f) g += wd * p;
  m = m + omb1 * (g - m);
  v = b2 * v + omb2 * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float step_size, float omb1, float b2,
                                                        float omb2, float eps, float wd, float bc2_sqrt) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt);
    adam_one(pp.y, gg.y, mm.y, vv.y, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt);
    adam_one(pp.z, gg.z, mm.z, vv.z, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt);
    adam_one(pp.w, gg.w, mm.w, vv.w, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i], m[i], v[i], step_size, omb1, b2, omb2, eps, wd, bc2_sqrt);
}

}  // namespace wnb

using namespace wnb;

extern "C" {

WNB_API int wnb_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                          double eps, double weight_decay, double bias_correction1, double bias_correction2, void* stream) {
  WNB_REQUIRE(p && g && m && v && n > 0, "adam_flat: null pointer / empty buffer");
  WNB_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_flat: buffers must be 16-byte aligned");
  WNB_REQUIRE(bias_correction1 > 0.0 && bias_correction2 > 0.0, "adam_flat: bias corrections must be positive (step >= 1)");
  const int64_t blocks = cdiv64(n >> 2 ? n >> 2 : n, 256);
  const int sms = device_sms();
  const unsigned grid = (unsigned)(blocks < (int64_t)sms * 8 ? blocks : (int64_t)sms * 8);
  // scalars are formed in double and rounded once, as torch does (1 - beta2 in float would already be off by 1.3e-5)
  adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)(lr / bias_correction1), (float)(1.0 - beta1),
                                                           (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay,
                                                           (float)sqrt(bias_correction2));
  WNB_CHECK_LAUNCH("adam_flat");
  return WNB_OK;
}

}  // extern "C"

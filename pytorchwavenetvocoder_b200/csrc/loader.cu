// loader.cu -- the arithmetic of the reference's train_generator (bin/train.py:67-312) on the device.
//
// The reference prepares every mini-batch on the host: slice a window out of the concatenated utterances, mu-law
// encode the WHOLE window (receptive field included, so every sample is encoded ~1.15 times), StandardScaler the
// features, transpose, stack, and copy three tensors to the GPU (train.py:157-185).  Here the concatenated waveform
// (float32) and frame-rate features stay resident in device ring buffers (uploaded once per utterance, asynchronously,
// from pinned memory) and ONE kernel cuts a whole batch out of them:
//   s0_b = s0 + b * hop                                                       (windows hop by batch_length, :178, :226)
//   x[b][i] = encode_mu_law(wave[s0_b + i])          t[b][i] = encode_mu_law(wave[s0_b + i + 1])      (train.py:167-169)
//   h[b][d][j] = float32((feat[row][d] - mean[d]) / scale[d])     row = s0_b / U + j   (up-sampling layer, :202-224)
//                                                                  row = frame_of_sample[s0_b + j]   (extend_time, :172-176)
// Positions are absolute stream positions; the buffers are rings (index = position mod capacity).
// in the dtypes numpy / sklearn use (see mulaw_encode_f32_kernel in elementwise.cu; StandardScaler.transform on a
// float32 array rounds to float32 after the subtraction and after the division -- in float64 arithmetic with
// scikit-learn 0.22, in float32 arithmetic with scikit-learn >= 1.x, selectable -- on float64 once at the final .float()).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/wnb200.h"
#include "common.cuh"

namespace wnb {

__device__ __forceinline__ int64_t mulaw_f32(float v, float mu, double log1pmu) {
  const float sgn = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
  const float arg = __fadd_rn(1.0f, __fmul_rn(mu, fabsf(v)));
  const float lg = (float)log((double)arg);
  const float num = __fmul_rn(sgn, lg);
  const double fx = __ddiv_rn((double)num, log1pmu);
  const double q = __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(fx, 1.0), 2.0), (double)mu), 0.5);
  return (int64_t)floor(q);
}

__global__ void __launch_bounds__(256) make_train_batch_kernel(
    const float* __restrict__ wave, const void* __restrict__ feat, const int32_t* __restrict__ frame_of_sample,
    int64_t s0, int64_t hop, int U, int64_t cap_s, int64_t cap_f, const double* __restrict__ mean,
    const double* __restrict__ scale, int64_t* __restrict__ x, int64_t* __restrict__ t, float* __restrict__ h, int B, int T,
    int Tf, int D, int feat_f64, float mu, double log1pmu) {
  const int b = blockIdx.y;
  const int64_t so = s0 + (int64_t)b * hop;
  // ---- waveform: x and the next-sample targets (each sample encoded once, written to both) ----
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= T; i += gridDim.x * blockDim.x) {
    const int64_t q = mulaw_f32(__ldg(wave + (so + i) % cap_s), mu, log1pmu);
    if (i < T) x[(int64_t)b * T + i] = q;
    if (i > 0) t[(int64_t)b * T + i - 1] = q;
  }
  // ---- aux features: (frames, D) rows -> (D, Tf) with the scaler applied ----
  const int64_t fo = so / U;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < D * Tf; e += gridDim.x * blockDim.x) {
    const int d = e / Tf, j = e - d * Tf;
    const int64_t row = (frame_of_sample ? (int64_t)frame_of_sample[(so + j) % cap_s] : fo + j) % cap_f;
    float out;
    if (feat_f64 == 1) {
      const double v = reinterpret_cast<const double*>(feat)[row * D + d];
      out = mean ? (float)__ddiv_rn(__dadd_rn(v, -mean[d]), scale[d]) : (float)v;
    } else {
      const float v = reinterpret_cast<const float*>(feat)[row * D + d];
      if (mean && feat_f64 == 2) {
        // scikit-learn >= 1.x: X -= astype(mean_, float32); X /= astype(scale_, float32)  (float32 arithmetic)
        out = __fdiv_rn(__fsub_rn(v, (float)mean[d]), (float)scale[d]);
      } else if (mean) {
        // scikit-learn 0.22 (the reference's pin): in-place ops with float64 operands, rounded to float32 after each
        const float c = (float)__dadd_rn((double)v, -mean[d]);   // X -= mean_
        out = (float)__ddiv_rn((double)c, scale[d]);             // X /= scale_
      } else {
        out = v;
      }
    }
    h[((int64_t)b * D + d) * Tf + j] = out;
  }
}

}  // namespace wnb

using namespace wnb;

extern "C" {

WNB_API int wnb_make_train_batch(const float* wave, const void* feat, const int32_t* frame_of_sample, int64_t s0,
                                 int64_t hop, int U, int64_t cap_s, int64_t cap_f, const double* mean, const double* scale,
                                 int64_t* x, int64_t* t, float* h, int B, int T, int Tf, int D, int feat_f64, int mu,
                                 void* stream) {
  WNB_REQUIRE(wave && feat && x && t && h, "make_train_batch: null pointer");
  WNB_REQUIRE(B > 0 && T > 0 && Tf > 0 && D > 0 && mu > 1 && (!mean == !scale) && s0 >= 0 && hop >= 0 && U >= 1 &&
                  cap_s > T && cap_f >= Tf,
              "make_train_batch: bad arguments");
  const int work = (T + 1 > D * Tf) ? T + 1 : D * Tf;
  dim3 grid((unsigned)((work + 255) / 256 < 148 ? (work + 255) / 256 : 148), (unsigned)B);
  make_train_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(wave, feat, frame_of_sample, s0, hop, U, cap_s, cap_f, mean,
                                                                 scale, x, t, h, B, T, Tf, D, feat_f64, (float)(mu - 1),
                                                                 log(1.0 + (double)(mu - 1)));   // wavenet.py:27 mu = mu - 1
  WNB_CHECK_LAUNCH("make_train_batch");
  return WNB_OK;
}

}  // extern "C"

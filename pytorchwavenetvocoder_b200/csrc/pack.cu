// pack.cu -- wnb_pack_weights: the state_dict layout <-> kernel layout transform as ONE launch per direction.
//
// The reference keeps every convolution as its own nn.Parameter (wavenet.py:188-210: 6 modules x L blocks + 6);
// the kernels want stacked / concatenated / transposed matrices (include/wnb200.h "Packed weight layouts").  Doing
// that with torch ops costs ~100 cat/stack/permute/pad launches per step plus their autograd mirror.  Here the whole
// transform is a table of strided 3-D copies executed by one kernel:
//   forward  (pack):    parameter tensors (absolute pointers)  ->  sections of one packed buffer
//   backward (unpack):  sections of the packed-gradient buffer  ->  one flat gradient buffer whose slices are the
//                       .grad tensors (scaled by the upstream gradient, a device scalar)
// The table is built once per model on the host (pytorchwavenetvocoder_b200/nets/packing.py) and lives on the device.
#include <cuda_runtime.h>

#include "../../include/wnb200.h"
#include "common.cuh"

namespace wnb {

__global__ void __launch_bounds__(256) pack_kernel(const WnbPackDesc* __restrict__ descs, const float* __restrict__ src_base,
                                                   float* __restrict__ dst_base, const float* __restrict__ scale) {
  const WnbPackDesc d = descs[blockIdx.x];
  const int total = d.n0 * d.n1 * d.n2;
  const float sc = scale ? __ldg(scale) : 1.f;
  float* dst = dst_base + d.dst;
  const float* src = (d.flags & WNB_PACK_SRC_ABS) ? reinterpret_cast<const float*>(d.src) : src_base + d.src;
  const float* src2 = (d.flags & WNB_PACK_SRC_ABS) ? reinterpret_cast<const float*>(d.src2) : src_base + d.src2;
  const int n12 = d.n1 * d.n2;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const int i0 = i / n12, r = i - i0 * n12, i1 = r / d.n2, i2 = r - i1 * d.n2;
    const int64_t so = (int64_t)i0 * d.ss0 + (int64_t)i1 * d.ss1 + (int64_t)i2 * d.ss2;
    float v;
    if (d.op == WNB_PACK_COPY) {
      v = __ldg(src + so);
    } else if (d.op == WNB_PACK_ADD2) {
      v = __ldg(src + so) + __ldg(src2 + so);
    } else {   // WNB_PACK_SUMPTR: src is a device table of nsum tensors of identical shape, summed in table order
      const float* const* tab = reinterpret_cast<const float* const*>(d.src);
      v = 0.f;
      for (int k = 0; k < d.nsum; k++) v += __ldg(tab[k] + so);
    }
    dst[(int64_t)i0 * d.ds0 + (int64_t)i1 * d.ds1 + (int64_t)i2 * d.ds2] = v * sc;
  }
}

}  // namespace wnb

using namespace wnb;

extern "C" {

WNB_API int wnb_pack_weights(const WnbPackDesc* descs, int ndesc, const float* src_base, float* dst_base,
                             const float* scale, void* stream) {
  WNB_REQUIRE(descs && ndesc > 0 && dst_base, "pack_weights: null pointer / empty table");
  dim3 grid((unsigned)ndesc, 8);
  pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(descs, src_base, dst_base, scale);
  WNB_CHECK_LAUNCH("pack_weights");
  return WNB_OK;
}

// cudaMemsetAsync on the caller's stream (gradient accumulators that the kernels add into)
WNB_API int wnb_zero(void* p, size_t bytes, void* stream) {
  WNB_REQUIRE(p || bytes == 0, "zero: null pointer");
  if (bytes) WNB_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return WNB_OK;
}

}  // extern "C"

// common.cuh -- error plumbing and small device helpers shared by every kernel file of libwnb200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/wnb200.h"

namespace wnb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// Optional per-launch CUDA-event timing by kernel kind (WNB_PROF_* in wnb200.h), see wnb_profile_enable().
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
struct ProfScope {
  int kind; cudaStream_t st;
  ProfScope(int k, cudaStream_t s) : kind(k), st(s) { prof_begin(k, s); }
  ~ProfScope() { prof_end(kind, st); }
};

#define WNB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::wnb::set_error(__VA_ARGS__);      \
      return WNB_ERR_INVALID;             \
    }                                     \
  } while (0)

#define WNB_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      ::wnb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return WNB_ERR_CUDA;                                                         \
    }                                                                              \
    ::wnb::count_launch();                                                         \
  } while (0)

#define WNB_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ::wnb::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return WNB_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace wnb

// gemm_simt.cuh -- fp32 FFMA register-tiled GEMM main loop used by the WNB_MATH_FP32 (parity) kernels.
//
// One CTA of 256 threads computes a (16*TM) x 64 output tile; thread (ty, tx) = (tid/16, tid%16) owns
// rows ty*TM.. and columns tx*4...  Operands are fetched through functors so that the callers can
// express the causal/dilated row gathers, the sigmoid/tanh row interleave, ReLU-on-load etc. without
// materialising anything in HBM.  K is streamed in chunks of 16 through shared memory with a
// register-staged software prefetch of the next chunk.
#pragma once
#include "common.cuh"

namespace wnb {

constexpr int kBN = 64;
constexpr int kBK = 16;
constexpr int kThreads = 256;
constexpr int kTN = 4;
constexpr int kBsLd = kBN + 4;

template <int TM>
struct TileSmem {
  static constexpr int BM = 16 * TM;
  static constexpr int AsLd = BM + 4;
  float As[kBK][AsLd];
  float Bs[kBK][kBsLd];
};

// fa(m, k) -> A element of tile row m, reduction index k (callers guard m; k < K guaranteed)
// fb(n, k) -> B element of tile column n
// AKC / BKC: operand is contiguous along k (else contiguous along m / n): picks the coalesced mapping.
template <int TM, bool AKC, bool BKC, class FA, class FB>
__device__ __forceinline__ void tile_mainloop(float (&acc)[TM][kTN], FA fa, FB fb, int K,
                                              TileSmem<TM>& sm) {
  constexpr int BM = 16 * TM;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float ra[TM], rb[kTN];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < TM; i++) {
      const int e = tid + i * kThreads;
      int m, k;
      if (AKC) { k = e & (kBK - 1); m = e >> 4; } else { m = e % BM; k = e / BM; }
      ra[i] = (k0 + k < K) ? fa(m, k0 + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kTN; i++) {
      const int e = tid + i * kThreads;
      int n, k;
      if (BKC) { k = e & (kBK - 1); n = e >> 4; } else { n = e & (kBN - 1); k = e >> 6; }
      rb[i] = (k0 + k < K) ? fb(n, k0 + k) : 0.f;
    }
  };

  gload(0);
  for (int k0 = 0; k0 < K; k0 += kBK) {
    __syncthreads();  // everybody finished reading the previous chunk
#pragma unroll
    for (int i = 0; i < TM; i++) {
      const int e = tid + i * kThreads;
      int m, k;
      if (AKC) { k = e & (kBK - 1); m = e >> 4; } else { m = e % BM; k = e / BM; }
      sm.As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < kTN; i++) {
      const int e = tid + i * kThreads;
      int n, k;
      if (BKC) { k = e & (kBK - 1); n = e >> 4; } else { n = e & (kBN - 1); k = e >> 6; }
      sm.Bs[k][n] = rb[i];
    }
    __syncthreads();
    if (k0 + kBK < K) gload(k0 + kBK);
#pragma unroll
    for (int k = 0; k < kBK; k++) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.As[k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
      const float4 bv = *reinterpret_cast<const float4*>(&sm.Bs[k][tx * 4]);
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < kTN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  __syncthreads();  // smem may be reused by the caller right away
}

template <int TM>
__device__ __forceinline__ void zero_acc(float (&acc)[TM][kTN]) {
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < kTN; j++) acc[i][j] = 0.f;
}

}  // namespace wnb

// resblock_tc.cu -- WNB_MATH_TF32: the fused residual block (reference wavenet.py:525-536) as ONE
// persistent, warp-specialised tcgen05 + TMA kernel for the BASELINE shape family
// (n_resch = 64, kernel_size = 2, n_aux <= 32, n_skipch % 64 == 0).
//
// Per 128-sample time tile (M = 128 rows = time, channels-last activations are K-major operands):
//   GEMM-1  D1[128 x 128] = [x(t-d) | x(t) | aux(t)] [128 x 160] * W1^T          (tcgen05.mma kind::tf32, SS)
//           A tiles arrive by TMA (zero fill for t-d < 0 = the causal left padding), W1 stays resident.
//   gate    z = sigmoid(D1[:, :64] + b) * tanh(D1[:, 64:] + b)   TMEM -> registers -> TMEM (never smem/HBM)
//   GEMM-2  D2[128 x 64] = z[128 x 64] * W2_chunk^T for the res chunk and S/64 skip chunks
//           (A operand straight from TMEM, W2 chunks streamed through a 2-deep TMA ring from L2)
//   out     xout = D2 + b2 + x(t) (x re-used from the A stage in smem) -> swizzled staging -> TMA store
//           skip += D2 + b2                                            -> staging -> TMA reduce-add (.add.f32)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 =
// epilogue (each owns the 32 TMEM lanes / time rows of its quarter and its own staging + bulk groups, so
// the epilogue warps never synchronise with each other).  All hand-offs are mbarriers; every wait is
// bounded (trap instead of hang).  fp32 storage in HBM, tf32 multiplies, fp32 accumulate.
//
// Shared memory (1 CTA / SM): W1 80 KB | A stage 80 KB | W2 ring 2 x 16 KB | staging 4 x 2 x 4 KB.
// TMEM: D1 128 cols | z 64 | D2 ping-pong 2 x 64  (512 allocated).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace wnb {

struct FwdParams {
  const float* xin; const float* haux; const float* w1; const float* b1; const float* w2; const float* b2;
  float* xout; float* skip; float* zsave;
  int B, T, R, S, Ap, ks, d, skip_init;
};

namespace tc {

constexpr int kTM = 128;                 // time rows per tile
constexpr int kR = 64;
constexpr int kAp = 32;
constexpr int kK1 = 2 * kR + kAp;        // 160
constexpr int kSubBytes = kTM * 32 * 4;  // one [128 x 32 fp32] swizzled sub-tile
constexpr int kNSubA = kK1 / 32;         // 5
constexpr int kW2SubBytes = 64 * 32 * 4;
constexpr int kW2StageBytes = 2 * kW2SubBytes;
constexpr int kStgBytes = 32 * 32 * 4;   // one [32 rows x 32 fp32] staging box
constexpr int kOffW1 = 0;
constexpr int kOffA = kOffW1 + kNSubA * kSubBytes;
constexpr int kOffW2 = kOffA + kNSubA * kSubBytes;
constexpr int kOffStg = kOffW2 + 2 * kW2StageBytes;
constexpr int kOffBar = kOffStg + 4 * 2 * kStgBytes;
constexpr int kSmemBytes = kOffBar + 256 + 1024;  // + alignment slack
constexpr int kThreadsTc = 192;
constexpr uint32_t kColD1 = 0, kColZ = 128, kColD2 = 192;

struct alignas(64) Maps {
  CUtensorMap x, haux, w1, w2, xout, skip;
};

enum { B_W1 = 0, B_AFULL, B_AEMPTY, B_W2F0, B_W2F1, B_W2E0, B_W2E1, B_D1F, B_D1E, B_ZF, B_D2F0, B_D2F1, B_D2E0,
       B_D2E1, B_COUNT };

__global__ void __launch_bounds__(kThreadsTc, 1)
resblock_fwd_tc_kernel(const __grid_constant__ Maps maps, const float* __restrict__ b1, const float* __restrict__ b2,
                       int B, int T, int S, int d, int has_xout, int skip_init) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_b = (T + kTM - 1) / kTM;
  const int ntiles = B * tiles_per_b;
  const int nchunks = (has_xout ? 1 : 0) + S / 64;
  const int row0 = has_xout ? 0 : kR;  // first W2 row of chunk 0

  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[B_W1], 1);
    ptx::mbar_init(&bars[B_AFULL], 1);
    ptx::mbar_init(&bars[B_AEMPTY], 128);
    ptx::mbar_init(&bars[B_W2F0], 1); ptx::mbar_init(&bars[B_W2F1], 1);
    ptx::mbar_init(&bars[B_W2E0], 1); ptx::mbar_init(&bars[B_W2E1], 1);
    ptx::mbar_init(&bars[B_D1F], 1);
    ptx::mbar_init(&bars[B_D1E], 128);
    ptx::mbar_init(&bars[B_ZF], 128);
    ptx::mbar_init(&bars[B_D2F0], 1); ptx::mbar_init(&bars[B_D2F1], 1);
    ptx::mbar_init(&bars[B_D2E0], 128); ptx::mbar_init(&bars[B_D2E1], 128);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      ptx::prefetch_tmap(&maps.x); ptx::prefetch_tmap(&maps.haux); ptx::prefetch_tmap(&maps.w1);
      ptx::prefetch_tmap(&maps.w2);
      ptx::mbar_arrive_expect_tx(&bars[B_W1], kNSubA * kSubBytes);
      for (int j = 0; j < kNSubA; j++) ptx::tma_load_2d(smem + kOffW1 + j * kSubBytes, &maps.w1, &bars[B_W1], j * 32, 0);
      uint32_t it = 0, gchunk = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int b = tile / tiles_per_b, t0 = (tile - b * tiles_per_b) * kTM;
        ptx::mbar_wait(&bars[B_AEMPTY], (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars[B_AFULL], kNSubA * kSubBytes);
        unsigned char* sA = smem + kOffA;
        ptx::tma_load_3d(sA + 0 * kSubBytes, &maps.x, &bars[B_AFULL], 0, t0 - d, b);
        ptx::tma_load_3d(sA + 1 * kSubBytes, &maps.x, &bars[B_AFULL], 32, t0 - d, b);
        ptx::tma_load_3d(sA + 2 * kSubBytes, &maps.x, &bars[B_AFULL], 0, t0, b);
        ptx::tma_load_3d(sA + 3 * kSubBytes, &maps.x, &bars[B_AFULL], 32, t0, b);
        ptx::tma_load_3d(sA + 4 * kSubBytes, &maps.haux, &bars[B_AFULL], 0, t0, b);
        for (int c = 0; c < nchunks; c++, gchunk++) {
          const uint32_t s = gchunk & 1;
          ptx::mbar_wait(&bars[B_W2E0 + s], ((gchunk >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&bars[B_W2F0 + s], kW2StageBytes);
          unsigned char* dst = smem + kOffW2 + s * kW2StageBytes;
          ptx::tma_load_2d(dst, &maps.w2, &bars[B_W2F0 + s], 0, row0 + c * 64);
          ptx::tma_load_2d(dst + kW2SubBytes, &maps.w2, &bars[B_W2F0 + s], 32, row0 + c * 64);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc1 = ptx::idesc_tf32(128, 128);
      constexpr uint32_t idesc2 = ptx::idesc_tf32(128, 64);
      const uint32_t sA = ptx::smem_u32(smem + kOffA), sW1 = ptx::smem_u32(smem + kOffW1),
                     sW2 = ptx::smem_u32(smem + kOffW2);
      ptx::mbar_wait(&bars[B_W1], 0);
      uint32_t it = 0, gchunk = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        ptx::mbar_wait(&bars[B_AFULL], it & 1);
        ptx::mbar_wait(&bars[B_D1E], (it & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int j = 0; j < kNSubA; j++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            ptx::mma_tf32_ss(tmem + kColD1, ptx::smem_desc_k_sw128(sA + j * kSubBytes + k * 32),
                             ptx::smem_desc_k_sw128(sW1 + j * kSubBytes + k * 32), idesc1, (j | k) != 0);
        ptx::tc_commit(&bars[B_D1F]);
        ptx::mbar_wait(&bars[B_ZF], it & 1);
        ptx::tc_fence_after();
        for (int c = 0; c < nchunks; c++, gchunk++) {
          const uint32_t s = gchunk & 1, ph = (gchunk >> 1) & 1;
          ptx::mbar_wait(&bars[B_W2F0 + s], ph);
          ptx::mbar_wait(&bars[B_D2E0 + s], ph ^ 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 2; kk++)
#pragma unroll
            for (int k = 0; k < 4; k++)
              ptx::mma_tf32_ts(tmem + kColD2 + s * 64, tmem + kColZ + kk * 32 + k * 8,
                               ptx::smem_desc_k_sw128(sW2 + s * kW2StageBytes + kk * kW2SubBytes + k * 32), idesc2,
                               (kk | k) != 0);
          ptx::tc_commit(&bars[B_W2E0 + s]);
          ptx::tc_commit(&bars[B_D2F0 + s]);
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;                  // time row inside the tile
    unsigned char* stg = smem + kOffStg + (warp - 2) * 2 * kStgBytes;
    const unsigned char* sX = smem + kOffA + 2 * kSubBytes;  // x(t) sub-tiles (channels 0-31, 32-63)
    uint32_t it = 0, gchunk = 0, nstore = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      const int b = tile / tiles_per_b, t0 = (tile - b * tiles_per_b) * kTM;
      ptx::mbar_wait(&bars[B_D1F], it & 1);
      ptx::tc_fence_after();
      // ---- gate: z = sigmoid(a) * tanh(g), 16 channels at a time, TMEM -> regs -> TMEM ----
#pragma unroll
      for (int g = 0; g < 4; g++) {
        float a[16], t[16], z[16];
        ptx::tmem_ld16(tmem + lane_base + kColD1 + g * 16, a);
        ptx::tmem_ld16(tmem + lane_base + kColD1 + 64 + g * 16, t);
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float av = a[i] + __ldg(b1 + g * 16 + i), tv = t[i] + __ldg(b1 + 64 + g * 16 + i);
          z[i] = (0.5f * ptx::tanh_approx(0.5f * av) + 0.5f) * ptx::tanh_approx(tv);
        }
        ptx::tmem_st16(tmem + lane_base + kColZ + g * 16, z);
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[B_ZF]);
      ptx::mbar_arrive(&bars[B_D1E]);
      if (!has_xout) ptx::mbar_arrive(&bars[B_AEMPTY]);  // last block: nothing else reads the A stage
      // ---- res / skip chunks ----
      for (int c = 0; c < nchunks; c++, gchunk++) {
        const uint32_t s = gchunk & 1, ph = (gchunk >> 1) & 1;
        const bool is_res = has_xout && c == 0;
        const int col0 = is_res ? 0 : (c - (has_xout ? 1 : 0)) * 64;  // channel offset inside xout / skip
        const float* bias = b2 + (is_res ? 0 : kR + col0);
        ptx::mbar_wait(&bars[B_D2F0 + s], ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; half++) {
          float v[32];
          {
            float lo[16], hi[16];
            ptx::tmem_ld16(tmem + lane_base + kColD2 + s * 64 + half * 32, lo);
            ptx::tmem_ld16(tmem + lane_base + kColD2 + s * 64 + half * 32 + 16, hi);
            ptx::tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i++) { v[i] = lo[i]; v[16 + i] = hi[i]; }
          }
          if (half == 1) {  // both halves of this D2 buffer are in registers: hand it back to the MMA warp
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars[B_D2E0 + s]);
          }
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += __ldg(bias + half * 32 + i);
          if (is_res) {
            const float4* xr = reinterpret_cast<const float4*>(sX + half * kSubBytes + row * 128);
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 xv = xr[j ^ (row & 7)];
              v[4 * j] += xv.x; v[4 * j + 1] += xv.y; v[4 * j + 2] += xv.z; v[4 * j + 3] += xv.w;
            }
            if (half == 1) ptx::mbar_arrive(&bars[B_AEMPTY]);  // x(t) consumed: the A stage may be refilled
          }
          unsigned char* sb = stg + (nstore & 1) * kStgBytes;
          if (lane == 0) ptx::bulk_wait_read<1>();  // the store that used this staging buffer has drained
          __syncwarp();
          float4* dstrow = reinterpret_cast<float4*>(sb + lane * 128);
#pragma unroll
          for (int j = 0; j < 8; j++)
            dstrow[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int tc = t0 + q * 32;
            if (is_res) ptx::tma_store_3d(&maps.xout, sb, half * 32, tc, b);
            else if (skip_init) ptx::tma_store_3d(&maps.skip, sb, col0 + half * 32, tc, b);
            else ptx::tma_reduce_add_3d(&maps.skip, sb, col0 + half * 32, tc, b);
            ptx::bulk_commit();
          }
          nstore++;
        }
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 tensor, dims innermost-first, 128B swizzle, zero fill out of bounds
static bool make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  uint64_t stride = sizeof(float);
  for (int i = 0; i < rank; i++) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    stride *= dims[i];
    if (i + 1 < rank) gstr[i] = stride;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc

bool resblock_fwd_tc_supported(int R, int S, int Ap, int ks) {
  return R == tc::kR && ks == 2 && Ap == tc::kAp && S >= 64 && S % 64 == 0;
}

int resblock_fwd_tc(const FwdParams& p, cudaStream_t st) {
  using namespace tc;
  if (p.zsave) {
    set_error("resblock_fwd_tc: zsave is not supported on the tcgen05 path");
    return WNB_ERR_UNSUPPORTED;
  }
  Maps maps;
  const uint64_t dx[3] = {(uint64_t)kR, (uint64_t)p.T, (uint64_t)p.B};
  const uint64_t dh[3] = {(uint64_t)kAp, (uint64_t)p.T, (uint64_t)p.B};
  const uint64_t ds[3] = {(uint64_t)p.S, (uint64_t)p.T, (uint64_t)p.B};
  const uint64_t dw1[2] = {(uint64_t)kK1, (uint64_t)(2 * kR)};
  const uint64_t dw2[2] = {(uint64_t)kR, (uint64_t)(kR + p.S)};
  const uint32_t box_ld[3] = {32, 128, 1}, box_st[3] = {32, 32, 1}, box_w1[2] = {32, 128}, box_w2[2] = {32, 64};
  bool ok = make_map(&maps.x, p.xin, 3, dx, box_ld) && make_map(&maps.haux, p.haux, 3, dh, box_ld) &&
            make_map(&maps.w1, p.w1, 2, dw1, box_w1) && make_map(&maps.w2, p.w2, 2, dw2, box_w2) &&
            make_map(&maps.xout, p.xout ? p.xout : p.xin, 3, dx, box_st) && make_map(&maps.skip, p.skip, 3, ds, box_st);
  if (!ok) {
    set_error("resblock_fwd_tc: cuTensorMapEncodeTiled failed or is unavailable");
    return WNB_ERR_CUDA;
  }
  static bool configured = false;
  if (!configured) {
    WNB_CUDA(cudaFuncSetAttribute(resblock_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    WNB_CUDA(cudaGetDevice(&dev));
    WNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int ntiles = p.B * ((p.T + kTM - 1) / kTM);
  const int grid = ntiles < sms ? ntiles : sms;
  resblock_fwd_tc_kernel<<<grid, kThreadsTc, kSmemBytes, st>>>(maps, p.b1, p.b2, p.B, p.T, p.S, p.d, p.xout ? 1 : 0,
                                                             p.skip_init);
  WNB_CHECK_LAUNCH("resblock_fwd_tc");
  return WNB_OK;
}

}  // namespace wnb

// resblock_tc.cu -- WNB_MATH_TF32: the fused residual block (reference wavenet.py:525-536) as ONE
// persistent, warp-specialised tcgen05 + TMA kernel for the BASELINE shape family
// (n_resch = 64, kernel_size = 2, n_aux <= 32, n_skipch % 64 == 0).
//
// Per 128-sample time tile (M = 128 rows = time, channels-last activations are K-major operands):
//   GEMM-1  D1[128 x 128] = [x(t-d) | x(t) | aux(t)] [128 x 160] * W1^T          (tcgen05.mma kind::tf32, SS)
//           A tiles arrive by TMA (zero fill for t-d < 0 = the causal left padding); W1 k-chunks and
//           W2 row-chunks stream from L2 through one 7-deep TMA ring (prefetched across tile boundaries).
//   gate    z = sigmoid(D1[:, :64] + b) * tanh(D1[:, 64:] + b)   TMEM -> registers -> TMEM (never smem/HBM)
//   GEMM-2  D2[128 x 64] = z[128 x 64] * W2_chunk^T for the res chunk and S/64 skip chunks
//           (A operand straight from TMEM, B operand = the ring slot)
//   out     xout = D2 + b2 + x(t) (x re-used from the A stage in smem) -> swizzled staging -> TMA store
//           skip += D2 + b2                                            -> staging -> TMA reduce-add (.add.f32)
//
// Warp roles (352 threads): warp 0 = TMA producer for the activation tiles, warp 2 = TMA producer for the
// weights (two independent producers: the next tile's A stage is requested as soon as the epilogue frees it,
// not after the current tile's W2 chunks), warp 1 = MMA issuer (one elected lane), warps 3-10 =
// epilogue: two per SM sub-partition, each owning (32 TMEM lanes / time rows) x (32 of the 64 columns of a
// chunk) with its own staging box + bulk group, so the epilogue warps never synchronise with each other.  All hand-offs are mbarriers; every wait is
// bounded (trap instead of hang).  fp32 storage in HBM, tf32 multiplies, fp32 accumulate.
//
// Shared memory (1 CTA / SM): weight ring 7 x 16 KB | A stage 80 KB | staging 8 x 4 KB.
// TMEM: D1 128 cols | z 64 | D2 4 x 64  (512 allocated).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

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
constexpr int kEpiWarps = 8;     // two per SM sub-partition: (TMEM lane quarter) x (column half)
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kNSlot = 7;                // weight ring depth: W1 k-chunks and W2 row-chunks, 16 KB each
constexpr int kSlotBytes = kSubBytes;
static_assert(kW2StageBytes == kSlotBytes, "W1 and W2 chunks share the ring slots");
constexpr int kND2 = 4;                  // D2 accumulator buffers in TMEM
constexpr int kOffRing = 0;
constexpr int kOffA = kOffRing + kNSlot * kSlotBytes;
constexpr int kOffStg = kOffA + kNSubA * kSubBytes;
constexpr int kOffBar = kOffStg + kEpiWarps * kStgBytes;
constexpr int kSmemBytes = kOffBar + 256 + 1024;  // + alignment slack
constexpr int kThreadsTc = 32 * (3 + kEpiWarps);  // warp 0 A-tile producer, 1 MMA issuer, 2 weight producer, 3.. epilogue
constexpr uint32_t kColD1 = 0, kColZ = 128, kColD2 = 192;

struct alignas(64) Maps {
  CUtensorMap x, haux, w1, w2, xout, skip;
};

enum { B_AFULL = 0, B_AEMPTY, B_D1F, B_D1E, B_ZF, B_WF0, B_WE0 = B_WF0 + kNSlot, B_D2F0 = B_WE0 + kNSlot,
       B_D2E0 = B_D2F0 + kND2, B_TILE0 = B_D2E0 + kND2, B_COUNT = B_TILE0 + 2 };
static_assert(8 * B_COUNT + 16 <= 256, "barrier block");
static_assert(kColD2 + kND2 * 64 <= 512, "TMEM columns");

// Tuning aid (WNB_FWD_PROF=1): cycles each role spends blocked on each hand-off, summed over CTAs.
enum { P_A_AEMPTY = 0, P_W_W2E, P_M_AFULL, P_M_D1E, P_M_ZF, P_M_W2F, P_M_D2E, P_M_TOTAL, P_E_D1F, P_E_D2F, P_E_BULK,
       P_E_GATE, P_E_TOTAL, P_E_MIN, P_E_MAX, P_COUNT };
__device__ unsigned long long g_fwd_prof[P_COUNT];

// Dynamic tile scheduler: CTA c starts on tile c and then takes tiles from this counter, so SMs that get a
// smaller share of the memory system simply process fewer tiles (with a static stride the slowest SM set the
// kernel time: ncu showed SMs active for only 80 % of the elapsed cycles).  The last CTA to finish resets both
// counters.  They are two device words owned by the launching (device, stream) (common.cuh sched_counters), so launches
// on different streams never share them.

template <bool PROF>
__global__ void __launch_bounds__(kThreadsTc, 1)
resblock_fwd_tc_kernel(const __grid_constant__ Maps maps, const float* __restrict__ b1, const float* __restrict__ b2,
                       int B, int T, int S, int d, int has_xout, int skip_init, unsigned int* __restrict__ sched) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * B_COUNT);
  volatile int* tile_ring = reinterpret_cast<volatile int*>(smem + kOffBar + 8 * B_COUNT + 8);  // 2 entries

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_b = (T + kTM - 1) / kTM;
  const int ntiles = B * tiles_per_b;
  const int nchunks = (has_xout ? 1 : 0) + S / 64;
  const int row0 = has_xout ? 0 : kR;  // first W2 row of chunk 0
  long long acc[P_COUNT] = {};
  auto wait = [&](uint64_t* bar, uint32_t parity, int k) {
    if constexpr (PROF) {
      const long long t = clock64();
      ptx::mbar_wait(bar, parity);
      acc[k] += clock64() - t;
    } else {
      ptx::mbar_wait(bar, parity);
    }
  };
  const long long t_begin = PROF ? clock64() : 0;

  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[B_AFULL], 1);
    ptx::mbar_init(&bars[B_AEMPTY], kEpiThreads);
    for (int i = 0; i < kNSlot; i++) { ptx::mbar_init(&bars[B_WF0 + i], 1); ptx::mbar_init(&bars[B_WE0 + i], 1); }
    ptx::mbar_init(&bars[B_D1F], 1);
    ptx::mbar_init(&bars[B_D1E], kEpiThreads);
    ptx::mbar_init(&bars[B_ZF], kEpiThreads);
    for (int i = 0; i < kND2; i++) { ptx::mbar_init(&bars[B_D2F0 + i], 1); ptx::mbar_init(&bars[B_D2E0 + i], kEpiThreads); }
    ptx::mbar_init(&bars[B_TILE0], 1); ptx::mbar_init(&bars[B_TILE0 + 1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::tmem_base_must_be_zero(*tmem_slot);
  constexpr uint32_t tmem = 0;

  // (single-thread roles are entered through elect.sync, not `lane == 0`: the compiler then knows exactly one thread
  //  is active and issues the uniform-datapath TMA / tcgen05 instructions without a per-thread ELECT loop)
  if (warp == 0) {
    // =============================== TMA producer: activation tiles ===============================
    // (its own warp, so that the next tile's A stage is requested the moment the epilogue releases it instead of
    //  queueing behind the W2 chunk stream of the current tile)
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&maps.x); ptx::prefetch_tmap(&maps.haux);
      int tile = blockIdx.x;
      for (uint32_t it = 0;; it++) {
        tile_ring[it & 1] = tile;                  // publish the tile (or the stop mark) to the other roles
        ptx::mbar_arrive(&bars[B_TILE0 + (it & 1)]);
        if (tile < 0) break;
        const int b = tile / tiles_per_b, t0 = (tile - b * tiles_per_b) * kTM;
        wait(&bars[B_AEMPTY], (it & 1) ^ 1, P_A_AEMPTY);
        ptx::mbar_arrive_expect_tx(&bars[B_AFULL], kNSubA * kSubBytes);
        unsigned char* sA = smem + kOffA;
        ptx::tma_load_3d(sA + 0 * kSubBytes, &maps.x, &bars[B_AFULL], 0, t0 - d, b);
        ptx::tma_load_3d(sA + 1 * kSubBytes, &maps.x, &bars[B_AFULL], 32, t0 - d, b);
        ptx::tma_load_3d(sA + 2 * kSubBytes, &maps.x, &bars[B_AFULL], 0, t0, b);
        ptx::tma_load_3d(sA + 3 * kSubBytes, &maps.x, &bars[B_AFULL], 32, t0, b);
        ptx::tma_load_3d(sA + 4 * kSubBytes, &maps.haux, &bars[B_AFULL], 0, t0, b);
        tile = (int)(atomicAdd(&sched[0], 1u) + gridDim.x);
        if (tile >= ntiles) tile = -1;
      }
    }
  } else if (warp == 2) {
    // =============================== TMA producer: weight ring ===============================
    // Every tile streams W1 (5 k-chunks of [128 x 32]) and then W2 (row chunks of [64 x 64]) from L2 through one
    // kNSlot-deep ring, so up to 112 KB of weights are in flight / prefetched across tile boundaries and the
    // MMA warp does not see the L2 latency (it did with W1 resident and only a 2-deep W2 ring).
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&maps.w1); ptx::prefetch_tmap(&maps.w2);
      uint32_t slot = 0, ph = 0;
      for (uint32_t it = 0;; it++) {
        ptx::mbar_wait(&bars[B_TILE0 + (it & 1)], (it >> 1) & 1);
        if (tile_ring[it & 1] < 0) break;
        for (int c = 0; c < kNSubA + nchunks; c++) {
          wait(&bars[B_WE0 + slot], ph ^ 1, P_W_W2E);
          ptx::mbar_arrive_expect_tx(&bars[B_WF0 + slot], kSlotBytes);
          unsigned char* dst = smem + kOffRing + slot * kSlotBytes;
          if (c < kNSubA) {
            ptx::tma_load_2d(dst, &maps.w1, &bars[B_WF0 + slot], c * 32, 0);
          } else {
            const int r = row0 + (c - kNSubA) * 64;
            ptx::tma_load_2d(dst, &maps.w2, &bars[B_WF0 + slot], 0, r);
            ptx::tma_load_2d(dst + kW2SubBytes, &maps.w2, &bars[B_WF0 + slot], 32, r);
          }
          if (++slot == kNSlot) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::idesc_tf32(128, 128);
      constexpr uint32_t idesc2 = ptx::idesc_tf32(128, 64);
      const uint32_t a_lo0 = ptx::desc_lo(ptx::smem_u32(smem + kOffA), 16);
      const uint32_t w_lo0 = ptx::desc_lo(ptx::smem_u32(smem + kOffRing), 16);
      constexpr uint32_t hi = ptx::kDescHiKSw128;
      uint32_t gchunk = 0, slot = 0, wph = 0;
      for (uint32_t it = 0;; it++) {
        ptx::mbar_wait(&bars[B_TILE0 + (it & 1)], (it >> 1) & 1);
        if (tile_ring[it & 1] < 0) break;
        wait(&bars[B_AFULL], it & 1, P_M_AFULL);
        wait(&bars[B_D1E], (it & 1) ^ 1, P_M_D1E);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < kNSubA; j++) {
          wait(&bars[B_WF0 + slot], wph, P_M_W2F);
          ptx::tc_fence_after();
          const uint32_t w_lo = w_lo0 + slot * (kSlotBytes >> 4), a_lo = a_lo0 + j * (kSubBytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; k++)
            ptx::mma_tf32_ss(tmem + kColD1, ptx::desc64(a_lo + 2 * k, hi), ptx::desc64(w_lo + 2 * k, hi), idesc1,
                             (j | k) != 0);
          ptx::tc_commit(&bars[B_WE0 + slot]);
          if (++slot == kNSlot) { slot = 0; wph ^= 1; }
        }
        ptx::tc_commit(&bars[B_D1F]);
        wait(&bars[B_ZF], it & 1, P_M_ZF);
        ptx::tc_fence_after();
        for (int c = 0; c < nchunks; c++, gchunk++) {
          const uint32_t s = gchunk % kND2, ph = (gchunk / kND2) & 1;
          wait(&bars[B_WF0 + slot], wph, P_M_W2F);
          wait(&bars[B_D2E0 + s], ph ^ 1, P_M_D2E);
          ptx::tc_fence_after();
          const uint32_t w_lo = w_lo0 + slot * (kSlotBytes >> 4);
#pragma unroll
          for (int kk = 0; kk < 2; kk++)
#pragma unroll
            for (int k = 0; k < 4; k++)
              ptx::mma_tf32_ts(tmem + kColD2 + s * 64, tmem + kColZ + kk * 32 + k * 8,
                               ptx::desc64(w_lo + kk * (kW2SubBytes >> 4) + 2 * k, hi), idesc2, (kk | k) != 0);
          ptx::tc_commit(&bars[B_WE0 + slot]);
          ptx::tc_commit(&bars[B_D2F0 + s]);
          if (++slot == kNSlot) { slot = 0; wph ^= 1; }
        }
      }
      if constexpr (PROF) acc[P_M_TOTAL] = clock64() - t_begin;
    }
  } else {
    // =============================== epilogue warps (3..10) ===============================
    // Warp e = warp - 3 owns TMEM lane quarter (warp & 3) -- the hardware restriction -- and column half (e >> 2):
    // two epilogue warps per SM sub-partition, so TMEM-load / bias / fence / bulk-wait latencies of one overlap
    // the arithmetic of the other.  Each warp has one 4 KB staging box and its own bulk-async group.
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int hf = (warp - 3) >> 2;                 // which 32 of the 64 columns of every chunk
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;                  // time row inside the tile
    unsigned char* sb = smem + kOffStg + (warp - 3) * kStgBytes;
    const uint32_t sb_row = ptx::smem_u32(sb + lane * 128);
    const unsigned char* sX = smem + kOffA + (2 + hf) * kSubBytes;  // x(t) sub-tile holding this warp's channels
    const float4* b1v = reinterpret_cast<const float4*>(b1);
    uint32_t gchunk = 0;
    for (uint32_t it = 0;; it++) {
      ptx::mbar_wait(&bars[B_TILE0 + (it & 1)], (it >> 1) & 1);
      const int tile = tile_ring[it & 1];
      if (tile < 0) break;
      const int b = tile / tiles_per_b, t0 = (tile - b * tiles_per_b) * kTM;
      wait(&bars[B_D1F], it & 1, P_E_D1F);
      const long long t_gate = PROF ? clock64() : 0;
      ptx::tc_fence_after();
      // ---- gate: z = sigmoid(a) * tanh(g), 16 channels at a time, TMEM -> regs -> TMEM ----
#pragma unroll
      for (int gg = 0; gg < 2; gg++) {
        const int g = hf * 2 + gg;
        float a[16], t[16], z[16];
        ptx::tmem_ld16(tmem + lane_base + kColD1 + g * 16, a);
        ptx::tmem_ld16(tmem + lane_base + kColD1 + 64 + g * 16, t);
        float4 ba[4], bt[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { ba[i] = __ldg(b1v + g * 4 + i); bt[i] = __ldg(b1v + 16 + g * 4 + i); }
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float bav[4] = {ba[i].x, ba[i].y, ba[i].z, ba[i].w}, btv[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float av = a[4 * i + k] + bav[k], tv = t[4 * i + k] + btv[k];
            z[4 * i + k] = (0.5f * ptx::tanh_approx(0.5f * av) + 0.5f) * ptx::tanh_approx(tv);
          }
        }
        ptx::tmem_st16(tmem + lane_base + kColZ + g * 16, z);
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[B_ZF]);
      ptx::mbar_arrive(&bars[B_D1E]);
      if constexpr (PROF) acc[P_E_GATE] += clock64() - t_gate;
      if (!has_xout) ptx::mbar_arrive(&bars[B_AEMPTY]);  // last block: nothing else reads the A stage
      // ---- res / skip chunks ----
      for (int c = 0; c < nchunks; c++, gchunk++) {
        const uint32_t s = gchunk % kND2, ph = (gchunk / kND2) & 1;
        const bool is_res = has_xout && c == 0;
        const int col0 = is_res ? 0 : (c - (has_xout ? 1 : 0)) * 64;  // channel offset inside xout / skip
        const float4* bias = reinterpret_cast<const float4*>(b2 + (is_res ? 0 : kR + col0) + hf * 32);
        float4 bv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) bv[j] = __ldg(bias + j);
        wait(&bars[B_D2F0 + s], ph, P_E_D2F);
        ptx::tc_fence_after();
        float v[32];
        {
          float lo[16], hi[16];
          ptx::tmem_ld16(tmem + lane_base + kColD2 + s * 64 + hf * 32, lo);
          ptx::tmem_ld16(tmem + lane_base + kColD2 + s * 64 + hf * 32 + 16, hi);
          ptx::tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; i++) { v[i] = lo[i]; v[16 + i] = hi[i]; }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars[B_D2E0 + s]);   // this warp's part of the D2 buffer is in registers
#pragma unroll
        for (int j = 0; j < 8; j++) {
          v[4 * j] += bv[j].x; v[4 * j + 1] += bv[j].y; v[4 * j + 2] += bv[j].z; v[4 * j + 3] += bv[j].w;
        }
        if (is_res) {
          const float4* xr = reinterpret_cast<const float4*>(sX + row * 128);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 xv = xr[j ^ (row & 7)];
            v[4 * j] += xv.x; v[4 * j + 1] += xv.y; v[4 * j + 2] += xv.z; v[4 * j + 3] += xv.w;
          }
          ptx::mbar_arrive(&bars[B_AEMPTY]);  // x(t) consumed: the A stage may be refilled
        }
        {
          const long long t = PROF ? clock64() : 0;
          if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store out of this staging box has drained
          __syncwarp();
          if constexpr (PROF) acc[P_E_BULK] += clock64() - t;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
          ptx::st_shared_v4(sb_row + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const int tc = t0 + q * 32;
          if (is_res) ptx::tma_store_3d(&maps.xout, sb, hf * 32, tc, b);
          else if (skip_init) ptx::tma_store_3d(&maps.skip, sb, col0 + hf * 32, tc, b);
          else ptx::tma_reduce_add_3d(&maps.skip, sb, col0 + hf * 32, tc, b);
          ptx::bulk_commit();
        }
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }
  if constexpr (PROF) {
    if (warp <= 2 || (warp == 3 && lane == 0)) {   // single-thread roles: only the elected lane has counts
      if (warp == 3) {
        acc[P_E_TOTAL] = clock64() - t_begin;
        atomicMin(&g_fwd_prof[P_E_MIN], (unsigned long long)acc[P_E_TOTAL]);
        atomicMax(&g_fwd_prof[P_E_MAX], (unsigned long long)acc[P_E_TOTAL]);
      }
      for (int k = 0; k < P_E_MIN; k++)
        if (acc[k]) atomicAdd(&g_fwd_prof[k], (unsigned long long)acc[k]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem);
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {  // last CTA out: re-arm the scheduler for the next launch
      sched[0] = 0;
      sched[1] = 0;
    }
  }
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
  if ((reinterpret_cast<uintptr_t>(p.b1) | reinterpret_cast<uintptr_t>(p.b2)) & 15) {
    set_error("resblock_fwd_tc: b1 / b2 must be 16-byte aligned");
    return WNB_ERR_INVALID;
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
  static const bool prof = [] { const char* e = getenv("WNB_FWD_PROF"); return e && e[0] == '1'; }();
  WNB_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(resblock_fwd_tc_kernel<false>), kSmemBytes));
  WNB_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(resblock_fwd_tc_kernel<true>), kSmemBytes));
  const int sms = device_sms();
  unsigned int* sched = sched_counters(st);
  if (!sched) { set_error("resblock_fwd_tc: cannot allocate the scheduler words"); return WNB_ERR_CUDA; }
  const int ntiles = p.B * ((p.T + kTM - 1) / kTM);
  const int grid = ntiles < sms ? ntiles : sms;
  if (prof) {  // tuning aid: synchronous, prints where each role of the pipeline waited
    unsigned long long zero[P_COUNT] = {}, h[P_COUNT];
    zero[P_E_MIN] = ~0ull;
    WNB_CUDA(cudaMemcpyToSymbol(g_fwd_prof, zero, sizeof(zero)));
    resblock_fwd_tc_kernel<true><<<grid, kThreadsTc, kSmemBytes, st>>>(maps, p.b1, p.b2, p.B, p.T, p.S, p.d,
                                                                     p.xout ? 1 : 0, p.skip_init, sched);
    WNB_CHECK_LAUNCH("resblock_fwd_tc");
    WNB_CUDA(cudaStreamSynchronize(st));
    WNB_CUDA(cudaMemcpyFromSymbol(h, g_fwd_prof, sizeof(h)));
    static const char* names[P_COUNT] = {"A:aempty", "W:w2e", "M:afull", "M:d1e", "M:zf", "M:w2f", "M:d2e", "M:total",
                                         "E:d1f", "E:d2f", "E:bulk", "E:gate", "E:total", "E:min", "E:max"};
    fprintf(stderr, "wnb200 fwd prof d=%d tiles/cta=%.2f (kcycles per CTA):", p.d, (double)ntiles / grid);
    for (int k = 0; k < P_COUNT; k++) fprintf(stderr, " %s=%.1f", names[k], h[k] / 1e3 / (k >= P_E_MIN ? 1 : grid));
    fprintf(stderr, "\n");
    return WNB_OK;
  }
  resblock_fwd_tc_kernel<false><<<grid, kThreadsTc, kSmemBytes, st>>>(maps, p.b1, p.b2, p.B, p.T, p.S, p.d,
                                                                    p.xout ? 1 : 0, p.skip_init, sched);
  WNB_CHECK_LAUNCH("resblock_fwd_tc");
  return WNB_OK;
}

}  // namespace wnb

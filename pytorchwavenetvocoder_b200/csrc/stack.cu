// stack.cu -- the whole residual stack of WaveNet.forward (reference wavenet.py:229-236 over :525-536) and its
// backward in "deferred skip" form, one C-ABI call each (WNB_MATH_TF32, n_resch = 64, kernel_size = 2, n_aux <= 32).
//
// The per-block formulation moves the (B,T,S) skip tensor through HBM in every block: a read-modify-write in the
// forward (skip += ...), and two reads in the backward (dz = dskip W2skip, dW2skip = dskip^T z) -- about 60 % of all
// bytes of a training step at 64 res / 512 skip channels.  Both directions are linear in the concatenated channel
// axis, so they are hoisted out of the block loop:
//
//   forward    z_l -> Z_all[:, :, l*R:(l+1)*R]     (resblock_z.cu writes the slice)
//              skip  = Z_all Wskip^T + bskip        one NT GEMM, K = L*R          (Wskip[s][l*R+c] = W2_l[R+s][c])
//   backward   dZ_all = dskip Wskip                  one NT GEMM, N = L*R in column blocks (A tile shared via L2)
//              per block: gate recompute + dz_l = dZ_all[l] + dout W2res_l + gate backward (one GEMM) ; dx
//              dW1_l, dW2res_l of ALL blocks: one segmented weight-gradient launch each (dpre_l, dx_l are kept)
//              dWskip = dskip^T Z_all                one weight-gradient GEMM, column groups of Z_all across CTAs
//
// Z_all is written once and read twice per step (it replaces the z recompute output of the per-block backward).
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/wnb200.h"
#include "common.cuh"
#include "tc_host.h"

namespace wnb {

static int pick_block(int total, const int* cands, int n) {
  for (int i = 0; i < n; i++)
    if (total % cands[i] == 0) return cands[i];
  return 0;
}

// WNB_STAGE_EPI=1 stages the epilogue operands (gate-backward dz slice, dX residual) through TMA tiles loaded by the producer
// warp instead of per-lane global row loads.  Built on the ncu finding that the epilogue warps stall on L1TEX, measured
// on B200 -- and SLOWER: gate backward 59.3 -> 66.2 us, dX 51.6 -> 58.8 us per block (one pipeline stage less and one more
// serial job for the single producer thread cost more than the row loads).  Off by default; kept as a measured negative.
static int stage_epi() {
  static const int on = [] { const char* e = getenv("WNB_STAGE_EPI"); return (e && e[0] == '1') ? 1 : 0; }();
  return on;
}

// The gate-backward and dX GEMMs of a block keep their weights (96 / 80 KB) in shared memory for the whole launch instead
// of streaming them from L2 for every 128-row tile; WNB_NT_WRES=0 restores the streaming ring (A/B runs).
static int resident_w() {
  static const int on = [] { const char* e = getenv("WNB_NT_WRES"); return (e && e[0] == '0') ? 0 : 1; }();
  return on && !stage_epi();
}

// Consecutive kernels of the block chain walk the time tiles in opposite directions (forward: odd blocks backwards;
// backward: gate backward front to back, dX back to front), so each starts on the rows its producer wrote last -- still
// in the 126 MB L2 -- instead of on the ones written a tensor ago.  WNB_ALT_DIR=0: every kernel front to back.
static int alternate_dir() {
  static const int on = [] { const char* e = getenv("WNB_ALT_DIR"); return (e && e[0] == '0') ? 0 : 1; }();
  return on;
}

static bool stack_supported(int R, int S, int Ap, int ks, int L) {
  if (!resblock_fwd_z_supported(R, Ap, ks)) return false;
  if (L < 1 || L > 64 /* wnb_stack_bwd's segment table, wgrad kMaxSeg */) return false;
  if (S % 32 != 0 || S < 32 || (S > 256 && S != 512)) return false;
  return true;
}

}  // namespace wnb

using namespace wnb;

WNB_API int wnb_stack_supported(int R, int S, int Ap, int ks, int L, int math_mode) {
  return (math_mode == WNB_MATH_TF32 && stack_supported(R, S, Ap, ks, L)) ? 1 : 0;
}

WNB_API int wnb_resblock_fwd_z(const float* xin, const float* haux, const float* w1, const float* b1, const float* w2res,
                               const float* b2res, float* xout, float* zall, int ldz, int zcol0, int B, int T, int R,
                               int Ap, int ks, int dilation, void* stream) {
  WNB_REQUIRE(B > 0 && T > 0 && dilation >= 1 && ldz >= R && zcol0 >= 0 && zcol0 + R <= ldz && ldz % 4 == 0,
              "resblock_fwd_z: bad shape");
  WNB_REQUIRE(xin && haux && w1 && b1 && zall && (!xout || (w2res && b2res)), "resblock_fwd_z: null pointer");
  if (!resblock_fwd_z_supported(R, Ap, ks)) {
    set_error("resblock_fwd_z: unsupported shape (R=%d Ap=%d ks=%d); use wnb_resblock_fwd", R, Ap, ks);
    return WNB_ERR_UNSUPPORTED;
  }
  return resblock_fwd_z(xin, haux, w1, b1, w2res, b2res, xout, zall, ldz, zcol0, B, T, dilation, (cudaStream_t)stream);
}

WNB_API int wnb_skip_gemm(const float* zall, const float* wskip, const float* bskip, float* skip, int B, int T, int K,
                          int S, int relu, void* stream) {
  WNB_REQUIRE(zall && wskip && skip && B > 0 && T > 0 && K % 32 == 0 && S % 32 == 0 && (S <= 256 || S == 512),
              "skip_gemm: bad arguments");
  const NtTcSeg seg[1] = {{zall, K, 0, K, wskip, S, K, 0, 0}};
  // 256 rows x (<=256 columns) per CTA tile: every Wskip chunk is fetched from L2 once per 256 rows (K = L*R is long,
  // the GEMM is bound by the L2 -> SM path: DESIGN.md section 3)
  const int nblk = S > 256 ? S / 256 : 1;
  const NtTcOpts o{nblk, 0, 0, 0, nt_default_m_tiles()};
  return gemm_nt_tc(seg, 1, S / nblk, skip, S, bskip, nullptr, 0, nullptr, 0, relu ? 1 : 0, 0, B, T, (cudaStream_t)stream,
                    nullptr, nullptr, nullptr, 0, 0, nullptr, &o);
}

WNB_API int wnb_stack_fwd(float* xs, int nxs, const float* haux, const float* w1, const float* b1, const float* w2res,
                          const float* b2res, const float* wskip, const float* bskip, float* zall, float* skip,
                          const int* dilations, int L, int B, int T, int R, int S, int Ap, int ks, int skip_relu,
                          void* stream) {
  WNB_REQUIRE(xs && haux && w1 && b1 && w2res && b2res && wskip && bskip && zall && skip && dilations,
              "stack_fwd: null pointer");
  WNB_REQUIRE(B > 0 && T > 0 && L >= 1 && nxs >= (L > 1 ? 2 : 1), "stack_fwd: bad shape");
  if (!stack_supported(R, S, Ap, ks, L)) {
    set_error("stack_fwd: unsupported shape (R=%d S=%d Ap=%d ks=%d)", R, S, Ap, ks);
    return WNB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t xsz = (size_t)B * T * R;
  const int K1 = ks * R + Ap, ldz = L * R;
  int rc;
  for (int l = 0; l < L; l++) {
    WNB_REQUIRE(dilations[l] >= 1, "stack_fwd: bad dilation");
    const float* xin = xs + (size_t)(l % nxs) * xsz;
    float* xout = (l + 1 < L) ? xs + (size_t)((l + 1) % nxs) * xsz : nullptr;
    ProfScope ps(WNB_PROF_FWD_BLOCK, st);
    if ((rc = resblock_fwd_z(xin, haux, w1 + (size_t)l * 2 * R * K1, b1 + (size_t)l * 2 * R, w2res + (size_t)l * R * R,
                             b2res + (size_t)l * R, xout, zall, ldz, l * R, B, T, dilations[l], st,
                             alternate_dir() ? (l & 1) : 0)) != WNB_OK)
      return rc;
  }
  ProfScope ps(WNB_PROF_SKIP_GEMM, st);
  return wnb_skip_gemm(zall, wskip, bskip, skip, B, T, ldz, S, skip_relu, stream);
}

// workspace: dZ_all (B,T,L*R) [+ one row of slack] | dpre of every block (L,B,T,2R) | dx of every block (L,B,T,R).
// dpre / dx of ALL blocks are kept so that the weight gradients dW1_l, dW2res_l of the whole stack are two launches at
// the end instead of 2L small ones inside the block loop (each of those spent a third of its time on launch, set-up
// and the atomic flush).
WNB_API size_t wnb_stack_bwd_workspace(int L, int B, int T, int R, int S, int Ap, int ks) {
  (void)S; (void)Ap; (void)ks;
  const size_t bt = (size_t)B * T;
  return sizeof(float) * (bt * L * R + (size_t)L * R + (size_t)L * bt * 2 * R + (size_t)L * bt * R);
}

WNB_API int wnb_stack_bwd(const float* xs, const float* haux, const float* zall, const float* dskip, const float* w1,
                          const float* b1, const float* w1t, const float* wgate, const float* wskip_t, float* dx0,
                          float* dhaux, float* dw1, float* db1, float* dw2res, float* db2res, float* dwskip,
                          float* dbskip, void* workspace, const int* dilations, int L, int B, int T, int R, int S,
                          int Ap, int ks, void* stream) {
  WNB_REQUIRE(xs && haux && zall && dskip && w1 && b1 && w1t && wgate && wskip_t && dx0 && dw1 && db1 && dw2res &&
                  db2res && dwskip && dbskip && workspace && dilations,
              "stack_bwd: null pointer");
  WNB_REQUIRE(B > 0 && T > 0 && L >= 1, "stack_bwd: bad shape");
  if (!stack_supported(R, S, Ap, ks, L)) {
    set_error("stack_bwd: unsupported shape (R=%d S=%d Ap=%d ks=%d)", R, S, Ap, ks);
    return WNB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bt = (size_t)B * T, xsz = bt * R;
  const int K1 = ks * R + Ap, ldz = L * R;
  float* dzall = (float*)workspace;
  float* dpre_all = dzall + bt * ldz + ldz;                 // (L, B, T, 2R)
  float* dx_all = dpre_all + (size_t)L * bt * 2 * R;        // (L, B, T, R); slot 0 unused (block 0 writes dx0)
  WNB_REQUIRE(L <= 64, "stack_bwd: at most 64 blocks");
  int rc;

  {  // dZ_all = dskip Wskip   (wskip_t: rows n = l*R + c, K = S contiguous)
    static const int cands[] = {256, 192, 128, 64, 32};
    const int nb = pick_block(ldz, cands, 5);
    WNB_REQUIRE(nb > 0, "stack_bwd: L*R must be a multiple of 32");
    const NtTcSeg seg[1] = {{dskip, S, 0, S, wskip_t, ldz, S, 0, 0}};
    const NtTcOpts o{ldz / nb, 0, 0, 0, nt_default_m_tiles()};
    ProfScope ps(WNB_PROF_DZALL_GEMM, st);
    if ((rc = gemm_nt_tc(seg, 1, nb, dzall, ldz, nullptr, nullptr, 0, nullptr, 0, 0, 0, B, T, st, nullptr, nullptr, nullptr,
                         0, 0, nullptr, &o)) != WNB_OK)
      return rc;
  }

  const float* dout = nullptr;   // gradient w.r.t. the block's residual output (none for the last block)
  for (int l = L - 1; l >= 0; l--) {
    const int d = dilations[l];
    const float* xin = xs + (size_t)l * xsz;
    const float* w1tl = w1t + (size_t)l * K1 * 2 * R;
    float* dzl = dzall + (size_t)l * R;          // (B,T,R) view with row pitch ldz
    float* dxin = (l == 0) ? dx0 : dx_all + (size_t)l * xsz;
    float* dpre = dpre_all + (size_t)l * bt * 2 * R;
    {  // gate recompute (pre = W1 [x(t-d) | x(t) | aux]) and dz_l = dZ_all[l] + dout W2res_l in ONE GEMM over the block
       // matrix wgate_l = [[W1, 0], [0, W2res^T]] (3R x (K1+R)): accumulator columns 0..2R-1 are the gate
       // pre-activations, columns 2R..3R-1 the residual part of dz; the epilogue turns both into dpre.
      const float* wg = wgate + (size_t)l * 3 * R * (K1 + R);
      // (resident weights: the zero blocks are skipped -- the first three segments touch accumulator columns 0..2R-1, the
      //  dout segment columns 2R..3R-1)
      const NtTcSeg sg[4] = {{xin, R, -d, R, wg, 3 * R, K1 + R, 0, 0, 0, 2 * R}, {xin, R, 0, R, wg, 3 * R, K1 + R, R, 0, 0, 2 * R},
                             {haux, Ap, 0, Ap, wg, 3 * R, K1 + R, 2 * R, 0, 0, 2 * R},
                             {dout, R, 0, R, wg, 3 * R, K1 + R, K1, 0, 2 * R, R}};
      const NtTcOpts o{1, ldz, 1, dout ? 1 : 0, 1, stage_epi(), resident_w()};
      ProfScope ps(WNB_PROF_GATE_BWD, st);
      if ((rc = gemm_nt_tc(sg, dout ? 4 : 3, dout ? 3 * R : 2 * R, dxin /* z output suppressed */, R,
                           b1 + (size_t)l * 2 * R, nullptr, 0, nullptr, 0, 0, 0, B, T, st, dzl, dpre, nullptr, 0, 0,
                           nullptr, &o)) != WNB_OK)
        return rc;
    }
    // dxin = dout + dpre(t+d) W1[:, tap0] + dpre(t) W1[:, tap1]   and   dhaux += dpre(t) W1[:, aux]
    if (dhaux) {
      ProfScope ps(WNB_PROF_DX_GEMM, st);
      // (the unshifted segment covers all R + Ap columns and goes first; the shifted one only adds to the R dx columns)
      const NtTcSeg sx[2] = {{dpre, 2 * R, 0, 2 * R, w1tl, K1, 2 * R, 0, R, 0, R + Ap}, {dpre, 2 * R, d, 2 * R, w1tl, R, 2 * R, 0, 0, 0, R}};
      const NtTcOpts ox{1, 0, 0, 0, 1, stage_epi(), resident_w(), alternate_dir()};
      if ((rc = gemm_nt_tc(sx, 2, R + Ap, dxin, R, nullptr, nullptr, 0, dout, R, 0, 0, B, T, st, nullptr, nullptr, dhaux,
                           Ap, R, nullptr, &ox)) != WNB_OK)
        return rc;
    } else {
      ProfScope ps(WNB_PROF_DX_GEMM, st);
      const NtTcSeg sx[2] = {{dpre, 2 * R, d, 2 * R, w1tl, K1, 2 * R, 0, 0}, {dpre, 2 * R, 0, 2 * R, w1tl, K1, 2 * R, 0, R}};
      const NtTcOpts ox{1, 0, 0, 0, 1, stage_epi(), resident_w(), alternate_dir()};
      if ((rc = gemm_nt_tc(sx, 2, R, dxin, R, nullptr, nullptr, 0, dout, R, 0, 0, B, T, st, nullptr, nullptr, nullptr, 0, 0,
                           nullptr, &ox)) != WNB_OK)
        return rc;
    }
    dout = dxin;
  }

  int seg_shift[64];
  for (int l = 0; l < L; l++) seg_shift[l] = -dilations[l];
  {  // every block's dW1 (128 x 160) += dpre_l^T [x_l(t-d_l) | x_l(t) | aux(t)], db1 = column sums of dpre_l: one launch
    WgBlock blk;
    blk.nops = 1;
    blk.ops[0] = WgOperand{dpre_all, 2 * R, 0, 4, 0, WG_LAYERED, 0};
    blk.c = dw1; blk.m_valid = 128; blk.db = db1;
    const WgOperand b[3] = {{xs, R, 0, 2, 0, WG_LAYERED | WG_SEG_SHIFT, 0}, {xs, R, 0, 2, 0, WG_LAYERED, 0},
                            {haux, Ap, 0, 1, 0, 0, 0}};
    const WgOpts o{1, L, seg_shift, 2 * R * K1, 2 * R};
    ProfScope ps(WNB_PROF_DW1, st);
    if ((rc = wgrad_tc_blocks(&blk, 1, b, 3, K1, B, T, st, &o)) != WNB_OK) return rc;
  }
  if (L > 1) {  // every block's dW2res (R x R) += dout_l^T z_l with dout_l = dx_{l+1} (rows 64..127 of the M-block: TMA
                // zero fill), db2res = column sums of dout_l; the last block has no residual output, hence L-1 segments
    WgBlock blk;
    blk.nops = 1;
    blk.ops[0] = WgOperand{dx_all + xsz, R, 0, 4, 0, WG_LAYERED, 0};
    blk.c = dw2res; blk.m_valid = R; blk.db = db2res;
    const WgOperand bz[1] = {{zall, ldz, 0, R / 32, 0, 0, R}};
    const WgOpts o{1, L - 1, seg_shift, R * R, R};
    ProfScope ps(WNB_PROF_DW2RES, st);
    if ((rc = wgrad_tc_blocks(&blk, 1, bz, 1, R, B, T, st, &o)) != WNB_OK) return rc;
  }

  {  // dWskip (S x L*R) += dskip^T Z_all, dbskip = column sums of dskip.  M-blocks per launch x columns per CTA is a
     // trade: fewer M-blocks leave TMEM for wider column groups (fewer CTAs re-reading the same dskip tile through L2,
     // fewer wasted bias columns per MMA) at the price of one launch per group of M-blocks.  Measured at the bench
     // shape: 4 M-blocks x 96 columns = one 827 us launch, 2 M-blocks x 192 columns = two 313 us launches.
     // WNB_DWSKIP_MB overrides (1..4).
    static int mb_per_launch = 0;
    if (!mb_per_launch) {
      const char* e = getenv("WNB_DWSKIP_MB");
      mb_per_launch = e ? atoi(e) : 2;
      if (mb_per_launch < 1 || mb_per_launch > 4) mb_per_launch = 2;
    }
    const int groups = ldz / 32;
    const int nmb = (S + 127) / 128 < mb_per_launch ? (S + 127) / 128 : mb_per_launch;
    int nB = 0;
    for (int c = 512 / nmb / 32 - 1 < 7 ? 512 / nmb / 32 - 1 : 7; c >= 1 && !nB; c--)
      if (groups % c == 0 && groups / c <= 128) nB = c;
    WNB_REQUIRE(nB > 0, "stack_bwd: cannot split L*R into column groups");
    for (int r0 = 0; r0 < S; r0 += 128 * nmb) {
      WgBlock blk[4];
      int nblk = 0;
      for (int r = r0; r < S && nblk < nmb; r += 128, nblk++) {
        blk[nblk].nops = 1;
        blk[nblk].ops[0] = WgOperand{dskip, S, r, 4, 0};
        blk[nblk].c = dwskip + (size_t)r * ldz;
        blk[nblk].m_valid = (S - r) < 128 ? (S - r) : 128;
        blk[nblk].db = dbskip + r;
      }
      const WgOperand bz[1] = {{zall, ldz, 0, nB, 0}};
      const WgOpts o{groups / nB};
      ProfScope ps(WNB_PROF_DWSKIP, st);
      if ((rc = wgrad_tc_blocks(blk, nblk, bz, 1, ldz, B, T, st, &o)) != WNB_OK) return rc;
    }
  }
  return WNB_OK;
}

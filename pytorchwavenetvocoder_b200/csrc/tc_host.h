// tc_host.h -- host-side entry points of the tcgen05 kernels (internal; the C ABI is include/wnb200.h).
#pragma once
#include <cuda_runtime.h>

namespace wnb {

// ---- weight gradients (wgrad_tc.cu):  C_i[128 x N] += A_i^T B  over all (b, t) --------------------------------
// One operand = a channels-last tensor (B,T,C); `groups` 32-channel groups starting at channel c0, shifted in time.
enum { WG_LAYERED = 1,     // segment mode: the tensor is (nseg*B, T, C); segment i reads batches i*B ..
       WG_SEG_SHIFT = 2 }; // segment mode: time shift = WgOpts::seg_shift[i] instead of `shift`
struct WgOperand { const float* base; int C; int c0; int groups; int shift; int flags; int seg_cstride; };
// (seg_cstride: segment mode, channel offset added per segment -- a block's slice of a concatenated tensor)
// One M-block: 128 output rows = 4 groups taken from up to 2 operands (TMA zero fill past a tensor's channels).
struct WgBlock { WgOperand ops[2]; int nops; float* c; int m_valid; float* db; };
// n_split > 1: the B operand holds n_split consecutive column groups of 32*sum(groups_b) channels each; CTA
// (g, s) = (blockIdx % n_split, blockIdx / n_split) accumulates column group g over time split s, so that the A
// operand (read by every group) is shared through L2.  Bias gradients are flushed by group 0 only.
// nseg > 1: nseg independent problems of identical structure in one launch (every residual block's dW1, say):
// operands pick their per-segment batch / shift / channel offsets with the WG_* flags, block i's C and db are at
// c + i*c_seg_stride, db + i*db_seg_stride; CTAs split the flattened (segment, time tile) space evenly.
struct WgOpts { int n_split; int nseg; const int* seg_shift; int c_seg_stride; int db_seg_stride; };
int wgrad_tc_blocks(const WgBlock* blocks, int nblocks, const WgOperand* b_ops, int nb_ops, int ldc, int B, int T,
                    cudaStream_t st, const WgOpts* opts = nullptr);
int wgrad_tc(const WgOperand* a_ops, int na_ops, const WgOperand* b_ops, int nb_ops, float* c, int ldc, int m_valid,
             float* db, int B, int T, cudaStream_t st);

// ---- NT GEMM (gemm_nt_tc.cu):  out[b,t,n] = sum_seg sum_k A_seg[b,t+shift,k] W_seg[n0+n][k0+k]  (+ epilogues) ----
// One segment: activation tensor (B,T,CA) read at rows t+shift, channels [0,K); weight matrix w (rows x ldw,
// K-contiguous) rows [n0, n0+N), columns [k0, k0+K).
struct NtTcSeg { const float* a; int CA; int shift; int K; const float* w; int w_rows; int w_cols; int k0; int n0;
                 // NtTcOpts::resident_weights only: this segment touches output columns [n_lo, n_lo + n_cnt) (weight rows
                 // n0 + n_lo ..); n_cnt == 0 means all N.  Lets a block matrix skip its zero blocks.
                 int n_lo = 0; int n_cnt = 0; };
// gate epilogue request: mode 1 = forward (z out), 2 = backward (dz in, z + dpre out); channels c0..c0+63 of R
struct NtTcGate { int mode; int c0; int R; const float* bias_sig; const float* bias_tanh; const float* dz; float* dpre; };
struct NtTcOpts {
  int n_blocks;      // > 1: N is one column block of n_blocks; block j uses weight rows n0 + j*N, output / bias / add /
                     //      mask columns j*N.. (tiles are ordered block-fastest, so the A tile is shared through L2)
  int gate_ld_dz;    // row pitch (floats) of the gate-backward dz input (0: gate_R)
  int gate_skip_z;   // gate-backward: do not write z
  int gate_fused_dz; // gate-backward shorthand with N == 192: accumulator columns 128..191 are added to dz
  int m_tiles;       // 2: a CTA tile is two 128-row time tiles sharing every weight chunk (plain epilogue, N <= 256):
                     //    halves the L2 -> SM weight traffic of the K >= 512 GEMMs (skip, dZ_all, post network)
  int stage_epilogue_operand;   // 1: the gate-backward dz slice / a 64-column residual `add` reaches the epilogue as a TMA
                                //    tile loaded by the producer warp instead of per-lane row loads (gate backward, dX)
  int resident_weights;   // 1: all weight chunks live in shared memory for the whole launch (loaded once per CTA), the ring
                          //    carries activations only; needs sum_seg K * n_cnt * 4 bytes (<= ~140 KB) and m_tiles == 1
  int reverse;            // 1: time tiles are processed from the end of the tensor backwards (a kernel that consumes what the
                          //    previous launch wrote front to back finds the rows written last still in L2)
};
// default of NtTcOpts::m_tiles for the K >= 512 GEMMs: 2, WNB_NT_MT=1 in the environment restores one tile per CTA
int nt_default_m_tiles();
int gemm_nt_tc(const NtTcSeg* segs, int nseg, int N, float* out, int ld_out, const float* bias, const float* mask,
               int ldmask, const float* add, int ldadd, int relu_out, int accumulate, int B, int T, cudaStream_t st,
               const float* gate_dz = nullptr, float* gate_dpre = nullptr, float* out2 = nullptr, int ld_out2 = 0,
               int out2_col0 = 0, const NtTcGate* gate = nullptr, const NtTcOpts* opts = nullptr);

// ---- fused residual block, deferred-skip form (resblock_z.cu) --------------------------------------------------
bool resblock_fwd_z_supported(int R, int Ap, int ks);
int resblock_fwd_z(const float* xin, const float* haux, const float* w1, const float* b1, const float* w2res,
                   const float* b2res, float* xout, float* zall, int ldz, int zcol0, int B, int T, int d,
                   cudaStream_t st, int reverse = 0);   // reverse: time tiles from the end of the tensor backwards

}  // namespace wnb

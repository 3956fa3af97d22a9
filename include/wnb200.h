/*
 * wnb200.h -- C ABI of libwnb200.so, the B200 (sm_100a) kernels behind the WaveNet vocoder hot paths.
 *
 * The reference (kan-bayashi/PytorchWaveNetVocoder v0.1.1) is pure Python on top of torch: it has no
 * FFI layer of its own.  This header is therefore OUR drop-in boundary: each entry
 * point replaces one piece of reference/wavenet_vocoder/nets/wavenet.py (cited per function) and is
 * what the host-side mirror `pytorchwavenetvocoder_b200.nets.WaveNet` binds through ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller (PyTorch) owns every buffer; nothing is allocated or freed in here;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and never synchronise;
 *   - return value: 0 = WNB_OK, negative = error; wnb_last_error() gives a thread-local message;
 *   - activations are CHANNELS-LAST fp32: (B, T, C) with C contiguous (the reference is (B, C, T));
 *     `Ap` is n_aux rounded up to a multiple of 32 (zero padded);
 *   - weights arrive in the PACKED layouts documented below (built from the reference state_dict by
 *     pytorchwavenetvocoder_b200/nets/packing.py).
 *
 * Packed weight layouts (R=n_resch, S=n_skipch, Q=n_quantize, ks=kernel_size, K1 = ks*R + Ap)
 *   wf   (ks, Q, R)   wf[k][q][r]     = causal.conv.weight[r][q][k]                 (wavenet.py:189)
 *   w1   (2R, K1)     rows 0..R-1 sigmoid branch, R..2R-1 tanh branch; columns j*R+c = tap j
 *                     (j=0 oldest, x[t-(ks-1-j)d]) of dil_{sigmoid,tanh}[l].conv.weight[o][c][j],
 *                     columns ks*R+a = aux_1x1_{sigmoid,tanh}[l].weight[o][a][0] (zero for a>=A)
 *   b1   (2R)         dil bias + aux bias                                            (wavenet.py:527-532)
 *   w2   (R+S, R)     rows 0..R-1 res_1x1[l].weight, rows R.. skip_1x1[l].weight     (wavenet.py:533-534)
 *   b2   (R+S)
 *   wp1  (S, S), bp1 (S), wp2 (Q, S), bp2 (Q)                                        (wavenet.py:209-210)
 *   *_t  the same matrices transposed (K-major), used by the backward and decode kernels.
 */
#ifndef WNB200_H_
#define WNB200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define WNB_API __attribute__((visibility("default")))
#else
#define WNB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define WNB_OK 0
#define WNB_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define WNB_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define WNB_ERR_UNSUPPORTED (-3)

#define WNB_MATH_FP32 0        /* SIMT FFMA, fp32 throughout (parity mode) */
#define WNB_MATH_TF32 1        /* tcgen05 kind::tf32 contractions, fp32 accumulate in TMEM */

#define WNB_MODE_ARGMAX 0
#define WNB_MODE_SAMPLING 1

/* ---- library / device info -------------------------------------------------------------- */
WNB_API int wnb_version(void);
WNB_API const char* wnb_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
WNB_API int64_t wnb_launch_count(void);

/* ---- a1/a2: mu-law codec (wavenet.py:17-30, 33-47) -- bit exact with numpy ----------------- */
WNB_API int wnb_mulaw_encode_f32(const float* x, int64_t* y, int64_t n, int mu, void* stream);
WNB_API int wnb_mulaw_encode_f64(const double* x, int64_t* y, int64_t n, int mu, void* stream);
WNB_API int wnb_mulaw_decode_f64(const int64_t* y, double* x, int64_t n, int mu, void* stream);

/* a17 (decode driver, bin/decode.py:318-319): generated codes -> PCM_16 samples for a whole batch on the device.  The
 * input domain is the mu codes, so `table` (ntab int16, device) holds decode_mu_law + the writer's quantisation per code
 * (built by nets.mulaw_pcm16_table); out[i] = table[idx[i] mod ntab]. */
WNB_API int wnb_lut_i16(const int32_t* idx, const int16_t* table, int16_t* out, int64_t n, int ntab, void* stream);

/* ---- a4/a5/a9: OneHot + causal conv as an embedding gather (wavenet.py:78-92, 513-516) ----- */
WNB_API int wnb_front_embed_fwd(const int64_t* x /*(B,T)*/, const float* wf, const float* bias /*(R)*/,
                        float* out /*(B,T,R)*/, int B, int T, int Q, int R, int ks, void* stream);
/* dwf (ks,Q,R) and dbias (R) are ACCUMULATED into (caller zeroes) */
WNB_API int wnb_front_embed_bwd(const int64_t* x, const float* dout /*(B,T,R)*/, float* dwf, float* dbias,
                        int B, int T, int Q, int R, int ks, void* stream);

/* ---- a6: UpSampling (wavenet.py:124-154) fused with the (B,A,Tf)->(B,T,Ap) layout change ----
 * U > 0: haux[b][tU+j][a] = h[b][a][t]*w[j] + bias[0];  U == 0: plain transpose (Tf == T). */
WNB_API int wnb_aux_upsample_fwd(const float* h /*(B,A,Tf)*/, const float* w /*(U)*/, const float* bias /*(1)*/,
                         float* haux /*(B,T,Ap)*/, int B, int A, int Ap, int Tf, int U, void* stream);
/* dw (U), dbias (1) accumulated into */
WNB_API int wnb_aux_upsample_bwd(const float* h, const float* dhaux /*(B,T,Ap)*/, float* dw, float* dbias,
                         int B, int A, int Ap, int Tf, int U, void* stream);

/* ---- a10: fused residual block forward (wavenet.py:525-536), one launch per layer ----------
 * xout = res_1x1(z) + xin, skip (+)= skip_1x1(z), z = sigmoid(..)*tanh(..).
 * xout may be NULL (last layer: its residual output is discarded, wavenet.py:230-238).
 * skip_init != 0: skip = value (first layer; python `0 + s0`), else skip += value.
 * zsave (B,T,R) may be NULL; when given, z is also stored (not used by the default backward). */
WNB_API int wnb_resblock_fwd(const float* xin, const float* haux, const float* w1, const float* b1,
                     const float* w2, const float* b2, float* xout, float* skip, float* zsave,
                     int B, int T, int R, int S, int Ap, int ks, int dilation, int skip_init,
                     int math_mode, void* stream);

/* 1 = fused kernel, 2 = composed tcgen05 path (pass the zsave scratch), 0 = not covered by math_mode */
WNB_API int wnb_resblock_fwd_supported(int R, int S, int Ap, int ks, int math_mode);

/* ---- a5 standalone: CausalConv1d.forward (wavenet.py:95-121), channels-last, fp32 FFMA, forward only.
 * x (B,T,Cin), w (Cout, ks*Cin) with w[o][j*Cin+c] = conv.weight[o][c][j], bias (Cout) or NULL, out (B,T,Cout). */
WNB_API int wnb_causal_conv1d_fwd(const float* x, const float* w, const float* bias, float* out, int B, int T,
                                  int Cin, int Cout, int ks, int dilation, void* stream);

/* ---- a18 / f1: train_generator on the device (bin/train.py:67-312) ------------------------------------------------
 * One launch cuts a whole mini-batch out of device-resident RING buffers: `wave` (cap_s floats) = the concatenated
 * float32 waveforms, `feat` (cap_f rows of D, float32 or float64 as read from the feature file) = the concatenated
 * feature frames; positions are absolute stream positions, ring index = position mod capacity.  Window b starts at
 * sample s0 + b*hop: x[b][i] = encode_mu_law(wave[..+i]), t[b][i] = encode_mu_law(wave[..+i+1]) (bit exact with
 * wnb_mulaw_encode_f32); h[b][d][j] = float32 StandardScaler transform ((v - mean[d]) / scale[d] with sklearn's dtype
 * behaviour; mean == NULL: none) of frame (s0 + b*hop)/U + j (up-sampling layer: Tf = T/U, frame_of_sample NULL) or of
 * frame frame_of_sample[sample] (extend_time mode: Tf = T; an int32 ring parallel to `wave`).
 * feat_f64: 1 = float64 rows; 0 = float32 rows, scaler in float64 arithmetic rounded to float32 after each step
 * (scikit-learn 0.22, the reference's pin); 2 = float32 rows, scaler in float32 arithmetic (scikit-learn >= 1.x).
 * x, t (B,T) int64; h (B,D,Tf) fp32. */
WNB_API int wnb_make_train_batch(const float* wave, const void* feat, const int32_t* frame_of_sample, int64_t s0,
                                 int64_t hop, int U, int64_t cap_s, int64_t cap_f, const double* mean, const double* scale,
                                 int64_t* x, int64_t* t, float* h, int B, int T, int Tf, int D, int feat_f64, int mu,
                                 void* stream);

/* ---- a12: the optimizer step (bin/train.py:457-460, :539 torch.optim.Adam) as one kernel over flat buffers ----------
 * p, g, m, v: parameters, gradients (the flat buffer wnb_pack_weights' backward direction fills), first and second moment
 * estimates, n floats each, 16-byte aligned, identical layout.  torch's (fused) Adam arithmetic in fp32: g += weight_decay*p;
 * m += (1-beta1)(g-m); v = beta2 v + (1-beta2) g^2; p -= lr/bias_correction1 * m / (sqrt(v)/sqrt(bias_correction2) + eps)
 * with bias_correction_i = 1 - beta_i^step supplied by the caller (step counted from 1); the scalar hyper-parameters are
 * doubles, combined in double and rounded to float once, as torch does. */
WNB_API int wnb_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                          double eps, double weight_decay, double bias_correction1, double bias_correction2, void* stream);

/* ---- f4: MLSA noise-shaping filter for a batch of utterances (bin/noise_shaping.py:46-87: pysptk
 * Synthesizer(MLSADF(order, alpha), hopsize).synthesis(x, tiled coefficients), i.e. per sample SPTK
 * mlsadf(x * exp(coef[0]), coef, order, alpha, pd) with a time-invariant coefficient vector) -------------------------
 * x: the utterances concatenated, int16 (x_is_i16 = 1: the wav samples, converted like np.float64(x)) or float64;
 * offsets (n_utts + 1 int64, device): utterance u is [offsets[u], offsets[u+1]); coef (order + 1 float64, device) =
 * pysptk.mc2b output (pass -coef for the inverse filter, noise_shaping.py:55-56); pd = Pade order 4 (pysptk's default) or
 * 5; gain = exp(coef[0]) computed by the caller; y: same layout as x, float64 or int16 (y_is_i16 = 1: truncation toward
 * zero like np.int16(y), noise_shaping.py:87).  fp64 arithmetic in SPTK's operation order without FMA contraction;
 * every utterance starts from a zero filter state (the reference lets the state of one file run on into the next file of
 * the same worker process). */
WNB_API int wnb_mlsa_filter(const void* x, int x_is_i16, const long long* offsets, int n_utts, const double* coef, int order,
                            double alpha, int pd, double gain, void* y, int y_is_i16, void* stream);

/* ---- pack_weights: state_dict layout <-> kernel layout as one launch per direction ----------
 * A table of strided 3-D copies dst[i0*ds0 + i1*ds1 + i2*ds2] = scale * f(src[i0*ss0 + i1*ss1 + i2*ss2]) executed by
 * one kernel (one block row per entry).  `src` / `src2` are absolute device pointers (WNB_PACK_SRC_ABS: the
 * reference-shaped nn.Parameters of wavenet.py:188-210) or element offsets from `src_base`; `dst` is an element
 * offset from `dst_base`.  Ops: COPY; ADD2 = src + src2 (dilated-conv bias + aux bias, wavenet.py:527-532);
 * SUMPTR = sum over a device table (at `src`) of `nsum` equally shaped tensors (the summed skip biases).
 * Forward: parameters -> packed buffer (layouts above, plus the transposes the backward wants).  Backward: packed
 * gradient buffer -> one flat buffer holding every parameter's .grad, scaled by the device scalar `scale` (NULL = 1). */
#define WNB_PACK_COPY 0
#define WNB_PACK_ADD2 1
#define WNB_PACK_SUMPTR 2
#define WNB_PACK_SRC_ABS 1
typedef struct WnbPackDesc {
  int64_t src, src2, dst;
  int32_t n0, n1, n2;
  int32_t ss0, ss1, ss2;
  int32_t ds0, ds1, ds2;
  int32_t op, flags, nsum;
} WnbPackDesc;
WNB_API int wnb_pack_weights(const WnbPackDesc* descs /*device*/, int ndesc, const float* src_base, float* dst_base,
                             const float* scale /*device scalar or NULL*/, void* stream);
/* cudaMemsetAsync(p, 0, bytes) on `stream`: gradient accumulators the kernels add into */
WNB_API int wnb_zero(void* p, size_t bytes, void* stream);

/* ---- per-launch timing of the stack kernels (measurement aid for bench.py's roofline block) ----
 * While enabled, wnb_stack_fwd / wnb_stack_bwd bracket every kernel launch with CUDA events on the launching stream;
 * wnb_profile_read() waits for them and returns the summed duration and launch count of one kind since the last
 * wnb_profile_enable().  Single-threaded use. */
#define WNB_PROF_FWD_BLOCK 0   /* resblock_fwd_z: one residual block, deferred-skip form */
#define WNB_PROF_SKIP_GEMM 1   /* skip = Z_all Wskip^T */
#define WNB_PROF_DZALL_GEMM 2  /* dZ_all = dskip Wskip */
#define WNB_PROF_GATE_BWD 3    /* gate recompute + dz + dpre */
#define WNB_PROF_DX_GEMM 4     /* dx (+ dhaux) */
#define WNB_PROF_DW1 5         /* dW1, db1 */
#define WNB_PROF_DW2RES 6      /* dW2res, db2res */
#define WNB_PROF_DWSKIP 7      /* dWskip, dbskip */
#define WNB_PROF_KINDS 8
WNB_API int wnb_profile_enable(int on);
WNB_API int wnb_profile_read(int kind, double* total_ms, int* launches);

/* ---- residual stack in deferred-skip form (training path of WaveNet.forward, wavenet.py:229-236 over :525-536) ----
 * The skip sum over all L blocks is one GEMM over the concatenated gate outputs,
 *     skip = [z_0 | ... | z_{L-1}] Wskip^T + bskip,   Wskip[s][l*R+c] = skip_1x1_l.weight[s][c],  bskip = sum_l skip bias_l
 * so a block only writes z_l into Z_all (B,T,L*R) and its residual output (no read-modify-write of skip per block),
 * and the backward hoists dZ_all = dskip Wskip and dWskip = dskip^T Z_all out of the block loop the same way.
 * WNB_MATH_TF32 only; wnb_stack_supported() says whether the shape is covered (else use the per-block entries). */
WNB_API int wnb_stack_supported(int R, int S, int Ap, int ks, int L, int math_mode);

/* one block: xout (B,T,R) = xin + res_1x1(z) (NULL for the last block), z -> zall[:, :, zcol0:zcol0+R] (row pitch ldz).
 * w1 (2R,K1), b1 (2R) as for wnb_resblock_fwd; w2res (R,R) = res_1x1.weight[o][c], b2res (R). */
WNB_API int wnb_resblock_fwd_z(const float* xin, const float* haux, const float* w1, const float* b1,
                               const float* w2res, const float* b2res, float* xout, float* zall, int ldz, int zcol0,
                               int B, int T, int R, int Ap, int ks, int dilation, void* stream);

/* skip (B,T,S) = zall (B,T,K) wskip^T + bskip (relu != 0: rectified, the form wnb_post_fwd consumes);
 * wskip (S,K) row-major, bskip (S) or NULL */
WNB_API int wnb_skip_gemm(const float* zall, const float* wskip, const float* bskip, float* skip, int B, int T,
                          int K, int S, int relu, void* stream);

/* whole stack: xs (nxs,B,T,R) holds the block inputs, xs[0] filled by the caller (wnb_front_embed_fwd); block l reads
 * xs[l % nxs] and writes xs[(l+1) % nxs] (nxs = L keeps every input for the backward, nxs = 2 ping-pongs).
 * w1 (L,2R,K1), b1 (L,2R), w2res (L,R,R), b2res (L,R), wskip (S,L*R), bskip (S); dilations: L host ints.
 * Outputs zall (B,T,L*R) and skip (B,T,S) (rectified when skip_relu != 0). */
WNB_API int wnb_stack_fwd(float* xs, int nxs, const float* haux, const float* w1, const float* b1,
                          const float* w2res, const float* b2res, const float* wskip, const float* bskip,
                          float* zall, float* skip, const int* dilations, int L, int B, int T, int R, int S,
                          int Ap, int ks, int skip_relu, void* stream);

/* backward of wnb_stack_fwd.  xs (L,B,T,R) and zall from the forward; dskip (B,T,S) = d loss / d skip.
 * w1t (L,K1,2R) as for wnb_resblock_bwd; wgate (L,3R,K1+R) = block matrix [[w1, 0], [0, w2res^T]] per block (gate
 * recompute and the residual part of dz share one GEMM); wskip_t (L*R,S) = wskip^T.
 * dx0 (B,T,R) = gradient of xs[0]; dhaux (B,T,Ap) accumulated into (or NULL); dw1, db1, dw2res, db2res, dwskip
 * (S,L*R), dbskip (S) are ACCUMULATED into (the last block's dw2res / db2res are left untouched: that conv has no
 * gradient, like the reference).  workspace: wnb_stack_bwd_workspace() bytes. */
WNB_API size_t wnb_stack_bwd_workspace(int L, int B, int T, int R, int S, int Ap, int ks);
WNB_API int wnb_stack_bwd(const float* xs, const float* haux, const float* zall, const float* dskip, const float* w1,
                          const float* b1, const float* w1t, const float* wgate, const float* wskip_t, float* dx0,
                          float* dhaux, float* dw1, float* db1, float* dw2res, float* db2res, float* dwskip,
                          float* dbskip, void* workspace, const int* dilations, int L, int B, int T, int R, int S,
                          int Ap, int ks, void* stream);

/* ---- a10 backward --------------------------------------------------------------------------
 * Recomputes the gate from xin/haux, then produces dxin and accumulates weight gradients.
 * dout = d(loss)/d(xout) (NULL for the last layer), dskip = d(loss)/d(skip_sum) (same for all layers).
 * dhaux (B,T,Ap) accumulated into (NULL when the aux path has no trainable producer).
 * dw1 (2R,K1), db1 (2R), dw2 (R+S,R), db2 (R+S) are ACCUMULATED into.
 * w1t = (ks+1 blocks) transposed W1: block j<ks is (R, 2R): w1t[j][c][o] = w1[o][j*R+c];
 *        block ks is (Ap, 2R): w1[o][ks*R+a];  w2t (R, R+S): w2t[c][o] = w2[o][c].
 * workspace: wnb_resblock_bwd_workspace() bytes. */
WNB_API size_t wnb_resblock_bwd_workspace(int B, int T, int R, int S, int Ap, int ks);
WNB_API int wnb_resblock_bwd(const float* xin, const float* haux, const float* dout, const float* dskip,
                     const float* w1, const float* b1, const float* w1t, const float* w2t,
                     float* dxin, float* dhaux, float* dw1, float* db1, float* dw2, float* db2,
                     void* workspace, int B, int T, int R, int S, int Ap, int ks, int dilation,
                     int math_mode, void* stream);

/* ---- a11: post network (wavenet.py:518-523) -------------------------------------------------
 * logits (B,T,Q) = wp2 * relu(wp1 * relu(skip) + bp1) + bp2 ; r1 (B,T,S) = relu(h1) is kept (needed by the
 * backward).  With WNB_MATH_TF32 `skip` is rectified IN PLACE first (its sign pattern -- all that
 * wnb_post_bwd needs from it -- is unchanged; wnb_post_bwd in tf32 mode relies on it); skip_rectified != 0
 * (tf32 only) says the producer already did that (wnb_skip_gemm / wnb_stack_fwd with relu) and skips the pass. */
WNB_API int wnb_post_fwd(float* skip, const float* wp1, const float* bp1, const float* wp2,
                 const float* bp2, float* r1, float* logits, int B, int T, int S, int Q,
                 int math_mode, int skip_rectified, void* stream);
/* dskip (B,T,S) out; dwp1,dbp1,dwp2,dbp2 accumulated into; workspace (B,T,S) floats */
WNB_API int wnb_post_bwd(const float* skip, const float* r1, const float* dlogits, const float* wp1t /*(S,S) [c][o]*/,
                 const float* wp2t /*(S,Q) [c][o]*/, float* dskip, float* dwp1, float* dbp1, float* dwp2,
                 float* dbp2, float* workspace, int B, int T, int S, int Q, int math_mode, void* stream);

/* ---- a12: CrossEntropyLoss(mean) on [:, start:] fused with its gradient (bin/train.py:534-536)
 * loss_sum (1 double, accumulated into; caller zeroes) receives the MEAN loss;
 * dlogits (B,T,Q) may be NULL (loss only); rows t < start get zero gradient. */
WNB_API int wnb_cross_entropy(const float* logits, const int64_t* target /*(B,T)*/, double* loss_sum,
                      float* dlogits, int B, int T, int Q, int start, void* stream);

/* ---- a14/a15/a16: persistent fast-generate kernel (wavenet.py:309-395, 397-511, 538-549) ----
 * One launch generates every sample of every utterance.  Queues live in `queues` (caller-allocated,
 * wnb_decode_workspace() bytes, need not be zeroed).
 *   xs      (B, P + max_n) int32: [0,P) = the padded/seed prefix (wavenet.py:330-334), the rest is
 *           written by the kernel (xs[b][P+i] = i-th generated sample);
 *   h       (B, A, Th) aux features BEFORE upsampling (U>0) or at sample rate (U==0);
 *   n_pad   number of left-replicated aux columns (wavenet.py:334): aux at padded position p is
 *           h_up[max(p - n_pad, 0)];
 *   n_samples (B) int32 per-utterance lengths; mode WNB_MODE_*; uniforms (B,max_n) optional
 *           externally supplied U[0,1) draws (tests), else Philox4x32-10(seed, utterance, step);
 *   logits_out (B, max_n, Q) optional teacher-check output (NULL in production).
 * Decode weights are K-major: wf (ks,Q,R), w1d (L,K1,2R), b1 (L,2R), w2d (L,R,R+S), b2 (L,R+S),
 * wp1d (S,S) [c][o], wp2d (S,Q) [c][o]. */
WNB_API size_t wnb_decode_workspace(int B, int R, int ks, const int32_t* host_dilations, int L);
WNB_API int wnb_decode(int32_t* xs, const float* h, const float* up_w, const float* up_b,
               const float* wf, const float* bf, const float* w1d, const float* b1, const float* w2d,
               const float* b2, const float* wp1d, const float* bp1, const float* wp2d, const float* bp2,
               const int32_t* host_dilations, int L, void* queues, const int32_t* n_samples,
               const float* uniforms, float* logits_out, int B, int P, int max_n, int n_pad, int Th,
               int Q, int A, int Ap, int R, int S, int ks, int U, int mode, uint64_t seed,
               void* stream);

/* ---- a14-a16, streaming variant: same contract as wnb_decode, but the K-major matrices arrive as ONE
 * packed stream consumed in order every step (a producer warp pushes it through a shared-memory ring
 * with cp.async.bulk while the consumer warps run the recurrence):
 *   per layer l: W1d [K1][O1] | W2res [R][Or] | W2skip [R][Os]   then  wp1d [S][Sp] | wp2d [S][Qp]
 * with O1 = 2R, Or = R, Os = Sp = S, Qp = Q each rounded up to a multiple of 4 (zero padded).
 * Returns WNB_ERR_UNSUPPORTED when the shape does not fit (caller uses wnb_decode). */
WNB_API size_t wnb_decode_stream_floats(int Q, int Ap, int R, int S, int ks, int L);
WNB_API int wnb_decode_stream(int32_t* xs, const float* h, const float* up_w, const float* up_b,
                              const float* wf, const float* bf, const float* stream, const float* b1,
                              const float* b2, const float* bp1, const float* bp2,
                              const int32_t* host_dilations, int L, void* queues, const int32_t* n_samples,
                              const float* uniforms, float* logits_out, int B, int P, int max_n, int n_pad,
                              int Th, int Q, int A, int Ap, int R, int S, int ks, int U, int mode,
                              uint64_t seed, void* stream_handle);

/* ---- a14-a16, warp-tiled variant for the BASELINE shape (R 64, S 512, Q 256, Ap 32, ks 2): same contract,
 * weights as ONE stream in warp-tile order (layout documented in csrc/decode_warp.cu; built by
 * nets/wavenet.py::_decode_warp_pack).  Batches of <= 74 utterances run ONE UTTERANCE PER 2-CTA CLUSTER (CL = 2):
 * each CTA streams its half of every matrix and the two exchange their halves of every activation vector through
 * distributed shared memory; the stream then holds the two per-CTA streams back to back. */
WNB_API size_t wnb_decode_warp_floats(int L, int CL);
/* debug aid: per-phase cycle counters of the free-running steps (16 int64 per CTA; NULL = off) */
WNB_API void wnb_decode_warp_set_timing(long long* device_buf);
WNB_API int wnb_decode_warp_supported(int Q, int Ap, int R, int S, int ks, int L);
WNB_API int wnb_decode_warp(int32_t* xs, const float* h, const float* up_w, const float* up_b, const float* wf,
                            const float* bf, const float* stream, const float* b1, const float* b2,
                            const float* bp1, const float* bp2, const int32_t* host_dilations, int L,
                            void* queues, const int32_t* n_samples, const float* uniforms, float* logits_out,
                            int B, int P, int max_n, int n_pad, int Th, int A, int U, int mode, uint64_t seed,
                            int W, int CL, void* stream_handle);
/* consumer warps (8 or 16) / CTAs per utterance (1 or 2) the launcher uses for B utterances; the stream layout depends
 * on both */
WNB_API int wnb_decode_warp_plan(int B);
WNB_API int wnb_decode_warp_cluster(int B);

#ifdef __cplusplus
}
#endif
#endif /* WNB200_H_ */

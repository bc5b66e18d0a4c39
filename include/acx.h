/*
 * libacx -- C ABI of the B200-native audioset-convnext inference hot path.
 *
 * The reference (topel/audioset-convnext-inf) is pure Python/PyTorch and has NO FFI, plugin
 * or operator interface: its boundary is the Python class API of
 *   src/audioset_convnext_inf/pytorch/convnext.py  (ConvNeXt.forward :287-331,
 *   forward_scene_embeddings :333-366, forward_frame_embeddings :369-402).
 * This header is therefore the interface a maintainer of the reference would bind with
 * ctypes underneath those three methods; every entry point cites the
 * reference lines whose computation it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless suffixed _host.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises, never allocates persistent device memory; the caller (torch) owns all
 *    buffers, including the workspace.
 *  - return 0 on success, non-zero ACX_ERR_* otherwise; acx_last_error() gives the message
 *    (thread-local).  No C++ exceptions cross this boundary.
 *  - activations are channels-last (N, H, W, C) with H = time, W = mel;
 *    `act_dtype` = ACX_BF16 (tensor-core path) or ACX_F32 (fp32-accurate SIMT path).
 */
#ifndef ACX_H_
#define ACX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACX_VERSION 100
#define ACX_API __attribute__((visibility("default")))

enum { ACX_OK = 0, ACX_ERR_ARG = 1, ACX_ERR_CUDA = 2, ACX_ERR_UNSUPPORTED = 3, ACX_ERR_WORKSPACE = 4 };
enum { ACX_BF16 = 0, ACX_F32 = 1 };

/* GEMM epilogues (acx_gemm). */
enum {
  ACX_EPI_BIAS = 0,            /* out = acc + bias[n]                                  (CX:231-234 conv) */
  ACX_EPI_BIAS_GELU = 1,       /* out = gelu_erf(acc + bias[n])                         (CX:79-80)        */
  ACX_EPI_BIAS_SCALE_RESID = 2 /* out = resid[m,n] + gamma[n] * (acc + bias[n])         (CX:81-86)        */
};

ACX_API int acx_version(void);
ACX_API const char* acx_last_error(void);
/* 1 if the current device is compute capability 10.x (B200), else 0 (and sets last error). */
ACX_API int acx_device_ok(void);

/* ---- front end: torchlibrosa Spectrogram + LogmelFilterBank + bn0 (CX:298-306) ------------- */

/* Split-precision operands of the tensor-core DFT are FP16 pairs of the value scaled by 2^ACX_FE_SCALE_LOG2
 * (hi = fp16(s x), lo = fp16(s x - hi): 22 mantissa bits, lo stays a normal fp16 down to |x| ~ 5e-4). */
#define ACX_FE_SCALE_LOG2 8

/* Reflect-pad n_fft/2 both sides (torchlibrosa STFT.forward: F.pad(..., mode='reflect')) and
 * split each fp32 sample into the scaled fp16 hi + lo pair above.  wave (B, L) fp32 ->
 * hi, lo (B, ld_pad) fp16 with ld_pad >= L + n_fft, ld_pad % 8 == 0.  If act_dtype == ACX_F32
 * writes one fp32 padded (unscaled) array to `hi` and ignores `lo`. */
ACX_API int acx_wave_prep(const float* wave, void* hi, void* lo, int B, int L, int n_fft, int ld_pad,
                  int act_dtype, void* stream);

/* Same, reading int16 PCM (the AudioSet HDF5 storage format, reference utils/data_generator.py:70-74) and
 * converting x / 32767. on the fly (utils/utilities.py:226-227): halves the H2D bytes of the eval loop. */
ACX_API int acx_wave_prep_pcm16(const int16_t* pcm, void* hi, void* lo, int B, int L, int n_fft, int ld_pad,
                        int act_dtype, void* stream);

/* fp32-accurate path: spec (B*T, 2*n_bins) fp32 (re | im, from acx_gemm_f32 on the padded wave)
 * -> power -> mel (sparse band form) -> 10*log10(max(.,1e-10)) -> bn0 affine.
 * mel_lo/mel_hi (n_mels) int32 give each filter's non-zero bin range [lo, hi); melT is
 * (n_mels, n_bins) fp32 (transposed melW); bn_scale/bn_shift (n_mels) fold bn0 (eps 1e-5). */
ACX_API int acx_power_mel_log(const float* spec, int ld_spec, int n_bins, const float* melT, const int32_t* mel_lo,
                      const int32_t* mel_hi, const float* bn_scale, const float* bn_shift, float* out,
                      int rows, int n_mels, void* stream);

/* tensor-core path: one fused kernel, frames x DFT (split-fp16 x3, fp32 accumulate in TMEM) ->
 * power -> x mel (split-bf16 x3) -> log10 -> bn0; never writes the spectrogram to HBM.
 * hi/lo: padded split waveform from acx_wave_prep.  dft_hi/lo: (n_chunks*128, n_fft) fp16 pairs of the
 * 2^ACX_FE_SCALE_LOG2-scaled windowed-DFT rows,
 * chunk c rows [0,64) = real rows of bins [64c, 64c+64), rows [64,128) = imag rows.
 * mel_hi/lo: (n_chunks*256, 64) bf16, chunk c rows m<224 = melW[64c + k, m] (K-major), rest 0.
 * out (B, T, n_mels) fp32. */
/* The same front end on FOLDED frames (frontend_folded.cu): for STFT rows with the real-input symmetry of a periodic-Hann
 * windowed DFT (W_re even, W_im odd about n = n_fft/2, zero at n = 0; torchlibrosa Spectrogram, CX:179-187) a frame's
 * spectrum is  re = E . Wre'^T, im = O . Wim'^T  with E[0] = x[512], E[j] = x[j] + x[1024-j], O[0] = 0, O[j] = x[j] - x[1024-j]:
 * half the tensor-core work of the dense product.
 *   acx_frame_fold(_pcm16): waveform (B, L) fp32 / int16 PCM -> f_hi / f_lo (B, T, n_fft) fp16 pairs of the
 *     2^ACX_FE_SCALE_LOG2-scaled folded frames [E (n_fft/2) | O (n_fft/2)], reflect padding (center=True), T = L/hop + 1.
 *   acx_frontend_folded: w_hi / w_lo (ceil(n_chunks/2)*256, n_fft/2) fp16 pairs, per pair of 64-bin chunks the rows
 *     [re chunk 2p | re chunk 2p+1 | im chunk 2p | im chunk 2p+1] (engine.fold_dft_weights); mel_* / bn_* / out as below. */
ACX_API int acx_frame_fold(const float* wave, void* f_hi, void* f_lo, int B, int L, int T, int n_fft, int hop, void* stream);
ACX_API int acx_frame_fold_pcm16(const int16_t* pcm, void* f_hi, void* f_lo, int B, int L, int T, int n_fft, int hop, void* stream);
ACX_API int acx_frontend_folded(const void* f_hi, const void* f_lo, const void* w_hi, const void* w_lo, const void* mel_hi,
                        const void* mel_lo, int n_chunks, const float* bn_scale, const float* bn_shift, float* out,
                        int B, int T, int n_fft, int n_mels, void* stream);

ACX_API int acx_frontend_fused(const void* hi, const void* lo, int ld_pad, const void* dft_hi, const void* dft_lo,
                       const void* mel_hi, const void* mel_lo, int n_chunks, const float* bn_scale,
                       const float* bn_shift, float* out, int B, int T, int n_fft, int hop, int n_mels,
                       void* stream);

/* ---- stem: Conv2d(1,96,4x4,s4,pad=(4,0)) + channels_first LayerNorm (CX:688-691, CX:227) ---- */
/* logmel (B, T, 224) fp32 -> out (B, H0, 56, 96) act_dtype, H0 = (T+4)/4 + 1.
 * w (16, 96) fp32 = stem weight transposed to (ky*4+kx, cout). */
ACX_API int acx_stem(const float* logmel, const float* w, const float* bias, const float* ln_w, const float* ln_b,
             void* out, int B, int T, int n_mels, int act_dtype, void* stream);
/* bf16 mode, same stem writing the group-planar layout [12][Mp][8] (Mp = B*H0*56 rounded up to 128) directly */
ACX_API int acx_stem_gp(const float* logmel, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                void* out, int B, int T, int n_mels, void* stream);

/* ---- Block part 1: depthwise 7x7 (pad 3, bias) + channels-last LayerNorm eps 1e-6 (CX:76-78) - */
/* x (B,H,W,C) -> y (B,H,W,C); w (49, C) act_dtype (dwconv weight transposed), W % 7 == 0. */
ACX_API int acx_dwconv_ln(const void* x, const void* w, const float* bias, const float* ln_w, const float* ln_b,
                  void* y, int B, int H, int W, int C, int act_dtype, void* stream);

/* ---- Block part 1 on the tensor cores (bf16 mode): depthwise 7x7 (pad 3) + bias (CX:76) as banded-Toeplitz
 *      tcgen05 GEMMs, then the channels-last LayerNorm (CX:78) as its own pass. -------------------------------- */
/* x (B,H,W,C) bf16 -> v (B,H,W,C) bf16 = conv + bias, NOT normalised; w (49, C) bf16; W in {56,28,14,7}, C % 8 == 0;
 * out of place. */
ACX_API int acx_dwconv_tc(const void* x, const void* w, const float* bias, void* v, int B, int H, int W, int C,
                  void* stream);
/* rows of C bf16 values: out = LayerNorm(in) * ln_w + ln_b, eps 1e-6, fp32 statistics; in == out allowed. */
ACX_API int acx_layernorm_rows(const void* in, const float* ln_w, const float* ln_b, void* out, long long M, int C,
                       void* stream);

/* Group-planar hand-off layout of stages 0 / 1 in bf16 mode: [C/8][Mp][8] (M = B*H*W pixels, each 16-byte group of 8
 * channels is a plane; the plane stride Mp is M rounded up to a multiple of 128 rows -- buffers hold C * Mp elements) -- the tensor-core depthwise conv then moves whole cache lines.  acx_gp_transpose converts
 * (M, C) row-major <-> planar (to_gp = 1 / 0); acx_dwconv_tc_gp is acx_dwconv_tc on planar x / v (W in {56, 28}). */
ACX_API int acx_dwconv_tc_gp(const void* x, const void* w, const float* bias, void* v, int B, int H, int W, int C,
                     void* stream);
ACX_API int acx_gp_transpose(const void* in, void* out, long long M, int C, int to_gp, void* stream);
/* stats[row] = (rstd, -mean * rstd) of the C channels of each row of a group-planar tensor (LayerNorm eps 1e-6) */
ACX_API int acx_gp_row_stats(const void* v, float* stats, long long M, int C, void* stream);

/* ---- downsample prologue: channels_first LayerNorm + 2x2/s2 patch gather (CX:231-234) ------- */
/* x (B,H,W,C) -> a (B*(H/2)*(W/2), 4C) with k = (dy*2+dx)*C + c  (the GEMM A operand). */
ACX_API int acx_ln_patchify(const void* x, const float* ln_w, const float* ln_b, void* a, int B, int H, int W, int C,
                    int act_dtype, void* stream);
/* same, x group-planar [C/8][B*H*W][8] bf16 (C = 96 / 192); `a` row-major as above */
ACX_API int acx_ln_patchify_gp(const void* x, const float* ln_w, const float* ln_b, void* a, int B, int H, int W, int C,
                       void* stream);

/* ---- the whole downsample layer as one implicit GEMM (CX:230-235: LayerNorm(channels_first) -> Conv2d(k2, s2)) ----
 * x group-planar bf16 [C/8][Mp][8] (Mp = B*H*W rounded up to 128); the LayerNorm'd 2x2 patches are formed in shared memory
 * as the tensor-core A operand (no patch matrix in HBM).  w: (2C, 4C) bf16, k = ((dy * C/8 + g) * 2 + dx) * 8 + c8 for
 * input channel 8 g + c8 at patch position (dy, dx).  out: (B*(H/2)*(W/2), 2C) bf16 row-major, or group-planar
 * [2C/8][Mp_out][8] when out_gp.  C = 96 / 192 / 384, W even (an odd last row of H is dropped like the conv does). */
ACX_API int acx_downsample_fused_gp(const void* x, const float* ln_w, const float* ln_b, const void* w, const float* bias,
                            void* out, int B, int H, int W, int C, int out_gp, void* stream);

/* ---- GEMMs: out[M,N] = epi(A[M,K] . W[N,K]^T)  (nn.Linear / conv-as-GEMM weight layout) ------ */
/* bf16 operands, fp32 accumulation in TMEM (tcgen05.mma, TMA-fed).  K % 8 == 0, N % 32 == 0. */
ACX_API int acx_gemm_bf16(const void* A, const void* W, void* out, int M, int N, int K, int epilogue,
                  const float* bias, const float* gamma, const void* resid, void* stream);
/* out = A . W^T + bias written group-planar, [N/8][Mp][8] with Mp = M rounded up to 128 (bias epilogue only) */
ACX_API int acx_gemm_bf16_gp_out(const void* A, const void* W, void* out, int M, int N, int K, const float* bias,
                         void* stream);
/* The two GEMMs of a Block MLP on group-planar activations (stage 2, whose MLP is not the fused kernel):
 *   pw1: hid (M, N = 4C) row-major = GELU(LN(v) . W1^T + b1), v planar [K/8][Mp][8], LayerNorm folded: w1f = bf16(W1 ln_w),
 *        b1f = b1 + W1 ln_b, ln_s[j] = sum_k w1f[j, k], stats = acx_gp_row_stats(v);
 *   pw2: x (planar [N/8][Mp][8], in place) += gamma * (hid . W2^T + b2). */
ACX_API int acx_gemm_bf16_pw1_gp(const void* v_gp, const void* w1f, void* hid, int M, int N, int K, const float* b1f,
                         const float* stats, const float* ln_s, void* stream);
ACX_API int acx_gemm_bf16_pw2_gp(const void* hid, const void* w2, void* x_gp, int M, int N, int K, const float* b2,
                         const float* gamma, void* stream);
/* fp32 SIMT GEMM of the fp32-accurate path.  Row m of A starts at
 * A + (m / rows_per_batch) * batch_stride + (m % rows_per_batch) * row_stride (elements), which
 * also expresses the overlapping STFT frames (row_stride = hop). */
ACX_API int acx_gemm_f32(const float* A, long long batch_stride, int rows_per_batch, int row_stride, const float* W,
                 float* out, int ldo, int M, int N, int K, int epilogue, const float* bias,
                 const float* gamma, const float* resid, void* stream);

/* ---- Block part 2, fused: pwconv1 -> GELU -> pwconv2 -> gamma -> +residual (CX:79-86) -------- */
/* y (M,C) bf16 LN output; x (M,C) bf16 residual stream, updated IN PLACE.
 * w1 (4C,C), w2 (C,4C) bf16.  The 4C hidden tile lives in TMEM/SMEM only.  C in {96,192}. */
ACX_API int acx_mlp_fused(const void* y, void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                  const float* gamma, int M, int C, void* stream);
/* Same with the Block's channels-last LayerNorm (CX:78, eps 1e-6) applied to the operand tile in shared memory:
 * v (M, C) bf16 is the RAW depthwise-conv output of acx_dwconv_tc; no normalised copy ever exists in HBM. */
ACX_API int acx_mlp_fused_ln(const void* v, void* x, const float* ln_w, const float* ln_b, const void* w1,
                     const float* b1, const void* w2, const float* b2, const float* gamma, int M, int C,
                     void* stream);
/* Group-planar variant: v and x are [C/8][Mp][8] bf16 (each 16-byte group of 8 channels of all rows is a plane, plane
 * stride Mp = M rounded up to 128), the hand-off layout of acx_dwconv_tc_gp.  Three LayerNorm modes:
 *   ln_w = ln_b = ln_s = NULL   v is already normalised;
 *   ln_w, ln_b given            the LayerNorm is applied to the operand tile in shared memory;
 *   ln_s given (ln_w/ln_b NULL) FOLDED: w1 = bf16(W1 diag(ln_w)), b1 = b1 + W1 ln_b, ln_s[j] = sum_c w1[j, c]; the kernel
 *                               computes per-row statistics only and applies the LayerNorm as a rank-1 correction
 *                               rstd (G - mean s) + b1 in the GELU epilogue. */
ACX_API int acx_mlp_fused_gp(const void* v, void* x, const float* ln_w, const float* ln_b, const float* ln_s,
                     const void* w1, const float* b1, const void* w2, const float* b2, const float* gamma, int M, int C,
                     void* stream);

/* ---- tail: mean over mel, max_t + mean_t, LayerNorm(768), fc 768->527, sigmoid (CX:279-285, 321-325) */
/* x (B,H,W,C) act_dtype -> scene (B,C) fp32 [post-LN], logits (B,n_cls), probs (B,n_cls).
 * pooled (B,C) fp32 is caller-provided scratch (the pre-LayerNorm pooled vector). */
ACX_API int acx_head(const void* x, const float* ln_w, const float* ln_b, const float* fc_w, const float* fc_b,
             float* pooled, float* scene, float* logits, float* probs, int B, int H, int W, int C, int n_cls,
             int act_dtype, void* stream);

/* x (B,H,W,C) act_dtype -> out (B,C,H,W) fp32: the layout forward_frame_embeddings returns (CX:399-402). */
ACX_API int acx_nhwc_to_nchw_f32(const void* x, float* out, int B, int H, int W, int C, int act_dtype, void* stream);

/* ---- next row (f4): demo preprocessing, reference demo_convnext.py:52-67 --------------------- */
/* torchaudio.functional.resample (sinc_interp_hann polyphase FIR, demo_convnext.py:53-59) fused with the demo's
 * constant pad / crop to a fixed clip length (demo_convnext.py:61-67):
 *   out[b, f*newf + p] = sum_{k<K} taps[k*newf + p] * x[b, f*orig + k - width]      (x = 0 outside [0, L_in))
 * for output index i < ceil(newf * L_in / orig), zero from there to n_out, cropped at n_out.
 * orig/newf are the rates divided by their gcd, K = 2*width + orig; taps (K, newf) fp32 is the TRANSPOSED
 * torchaudio kernel, built on the host (preprocess.sinc_resample_taps).  x (B, ld_in), out (B, ld_out). */
ACX_API int acx_resample_fit(const float* x, int ld_in, const float* taps, float* out, int ld_out, int B, int L_in,
                     int orig, int newf, int width, int n_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACX_H_ */

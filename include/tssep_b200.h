/*
 * tssep_b200 — C ABI of the B200 (sm_100a) TS-SEP inference hot path.
 *
 * The reference (merlresearch/tssep) is pure Python and has no FFI of its own;
 * its plug-in boundary is the `factory:` key of its config dicts
 * (tssep/exp/init_cfg_common.yaml:7-83, README.md:98-99).  The Python classes
 * in tssep_b200/ mirror those factories and call the functions below through
 * ctypes.  Every function states the reference operator (file:line) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless said otherwise;
 *     the library never allocates, frees or retains caller memory;
 *   - work is enqueued on `stream` (a cudaStream_t); no implicit device sync;
 *   - return 0 on success, negative on error; message via tssep_last_error()
 *     (thread-local, valid until the next failing call on the same thread);
 *   - no exceptions cross this boundary, there is no CPU fallback;
 *   - no hidden inputs: the library reads no environment variables (tuning knobs exist only in a debug
 *     build, -DTSSEP_DEBUG_KNOBS) and keeps no mutable global state besides the thread-local error message;
 *   - kernels launch on the CURRENT CUDA device: the caller makes the device that owns `stream` and the
 *     pointers current (tssep_b200/_lib.py does so for every call);
 *   - "bf16" pointers are raw uint16 bfloat16 bit patterns; cfloat = float[2].
 */
#ifndef TSSEP_B200_H_
#define TSSEP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tssep_stream_t; /* cudaStream_t */

/* Bumped whenever a signature or struct layout below changes; tssep_b200/_lib.py refuses a library that
 * reports another value. */
#define TSSEP_ABI_VERSION 4

const char* tssep_last_error(void);
int tssep_abi_version(void);
/* Fills SM count and compute capability of the current device. */
int tssep_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------
 * (1) Feature front end
 * ---------------------------------------------------------------------- */

/* fe.stft(signal): padertorch STFT, call site tssep/train/model.py:504,
 * parameters tssep/exp/init_cfg_common.yaml:45-50.
 * audio (n_signals, num_samples) f32 -> X (n_signals, T, size/2+1) cfloat.
 * window: (window_length) f32 analysis window; twiddle: (size/2) cfloat,
 * twiddle[k] = exp(-2*pi*i*k/size).  size must be a power of two in [8, 4096],
 * window_length <= size.  T must equal the reference frame count. */
int tssep_stft(const float* audio, int64_t n_signals, int64_t num_samples, const float* window,
               const float* twiddle, int size, int shift, int window_length, int fading, int64_t T,
               float* X, tssep_stream_t stream);

/* First pass of stft_to_feature: global statistics.
 * TorchMFCC.stft_to_feature   tssep/train/feature_extractor_torchaudio.py:93-106
 * Log1pMaxNormAbsSTFT         tssep/train/feature_extractor.py:233-248
 * X (n_items, T, F) cfloat (item stride x_item_stride cfloats, so the reference
 * channel of a multi-channel STFT can be addressed in place).
 * absmax_key / maxdb_key: (n_items) u32 order-preserving keys, ZEROED by this call.
 * mel_t (n_mels, F) f32 transposed filterbank with support [mel_lo, mel_hi) per
 * filter; meldb (n_items, T, n_mels) f32 out.  n_mels == 0 skips the mel part. */
int tssep_feature_stats(const float* X, int64_t n_items, int64_t x_item_stride, int64_t T, int F,
                        const float* mel_t, const int32_t* mel_lo, const int32_t* mel_hi, int n_mels,
                        uint32_t* absmax_key, uint32_t* maxdb_key, float* meldb, tssep_stream_t stream);

/* Second pass: writes feature rows [mfcc(n_mfcc) | log1p spectrum(F if with_log1p)].
 * ConcaternatedSTFTFeatures.stft_to_feature tssep/train/feature_extractor.py:352-360.
 * couple_batch != 0 reproduces torchaudio's batch-coupled top_db cut-off (one max
 * over the whole batched tensor).  feat_f32 (n_items, T, Din) and/or feat_bf16
 * (n_items*T, ld_bf16) may be NULL. */
int tssep_feature_write(const float* X, int64_t n_items, int64_t x_item_stride, int64_t T, int F,
                        const uint32_t* absmax_key, const uint32_t* maxdb_key, const float* meldb,
                        const float* dct, int n_mels, int n_mfcc, int with_log1p, float top_db,
                        int couple_batch, float* feat_f32, uint16_t* feat_bf16, int64_t ld_bf16,
                        tssep_stream_t stream);

/* Weighted prediction error dereverberation (tssep/train/enhancer.py:292-367: WPE / ChannelWiseWPE, wrappers around
 * nara_wpe.wpe.wpe_v8 -- nara_wpe>=0.0.11 is a third-party dependency absent from the reference tree; the kernels
 * restate its published algorithm).  Y, X: (D, T, F) complex64 (interleaved float pairs), D <= 8, D * taps small enough
 * for the shared-memory solve ((D*taps) * (D*taps + D) * 16 bytes <= 200 KiB).  Per frequency and iteration:
 * lambda_t = mean_d |X|^2 (mean over +-psd_context frames, floor 1e-10 * max_t), R = sum_t Yt Yt^H / lambda_t,
 * P = sum_t Yt Y^H / lambda_t with Yt the taps delayed frames (delay, ..., delay + taps - 1), G = R^-1 P in f64,
 * X = Y - G^H Yt.  statistics_mode: 0 = 'full' (every frame), 1 = 'valid' (frames >= delay + taps - 1).
 * workspace: device memory of tssep_wpe_workspace_bytes(...) bytes, 256-byte aligned (scratch, no state between calls).
 * ChannelWiseWPE = the same call on the (1, T, D*F) view of the channel-major signal. */
int64_t tssep_wpe_workspace_bytes(int D, int64_t T, int F, int taps);
int tssep_wpe(const float* Y, int D, int64_t T, int F, int taps, int delay, int iterations, int psd_context,
              int statistics_mode, float* X, void* workspace, int64_t workspace_bytes, tssep_stream_t stream);

/* Multi-channel / normalised feature variants of the reference (tssep/train/feature_extractor.py:13-168, :266-287):
 * tssep_log1p_abs: out[i] = log1p(|X[i]|) for n complex values (Log1pAbsSTFT; MVNLog1pAbsSTFT subtracts the mean over
 *   frames afterwards: tssep_instance_norm mode 1 along the frame axis);
 * tssep_ipd: inter-channel phase differences of X (lead, D, TF) cfloat against channel second_channel[d] (D int32):
 *   cos_out / sin_out (lead, D, TF) f32 = Re / Im of X[d] conj(X[second[d]]) / |.|. */
int tssep_log1p_abs(const float* X, int64_t n, float* out, tssep_stream_t stream);
int tssep_ipd(const float* X, int64_t lead, int D, int64_t TF, const int32_t* second_channel, float* cos_out,
              float* sin_out, tssep_stream_t stream);

/* f32 (rows, cols) -> bf16 (rows, ld) conversion (zero padded columns). */
int tssep_cast_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, uint16_t* dst,
                    int64_t ld_dst, tssep_stream_t stream);

/* InstanceNorm / InstanceNorm_v2 (tssep/train/net.py:250-330) along the middle axis of a contiguous
 * (outer, cols, inner) f32 tensor (any single `dim` of a contiguous tensor is such a view).
 * mode 0: (x - mean) / std  (population std, or Bessel-corrected when unbiased != 0)      InstanceNorm
 * mode 1: x - mean;  mode 2: x / sqrt(mean(x^2))       the two steps of InstanceNorm_v2 when its axes differ */
int tssep_instance_norm(const float* src, int64_t outer, int64_t cols, int64_t inner, int mode, int unbiased,
                        float* dst, tssep_stream_t stream);

/* ------------------------------------------------------------------------
 * (2) Conditioning + bulk contractions (torch.nn.Linear / LSTM input GEMMs,
 *     tssep/train/rnnp.py:87-96, tssep/train/net.py:663-666, :871-894)
 * ---------------------------------------------------------------------- */

/* Speaker-embedding conditioning folded into the first post_net projection.
 * mode 0 ('mul', net.py:871-874):  Wk[z,n,f] = bf16(W[n,f] * e[z,f]),  bias_k[z,n] = b[n]
 * mode 1 ('cat', net.py:879-894):  bias_k[z,n] = b[n] + sum_a W[n,F+a] * e[z,a]
 *                                  (Wk is not written; the GEMM uses the shared W[:, :F])
 * W (N, ldw) f32 packed input weights, e (Z, A) f32 embeddings already in slot order. */
int tssep_fold_embedding(int mode, const float* W, int64_t ldw, const float* b, const float* e,
                         int64_t Z, int N, int F, int A, uint16_t* Wk, int64_t ld_wk, float* bias_k,
                         tssep_stream_t stream);

/* Batched contraction on tcgen05/TMEM tensor cores:
 *   out[z] = act(alpha * A[z / a_div] . B[z % b_mod]^T + bias[z % b_mod])      z in [0, batch)
 * A (rows, K) bf16 row-major (lda), B (N, K) bf16 row-major (ldb); lda, ldb multiples
 * of 8 elements, base pointers 16-byte aligned, batch strides whole rows.
 * Offsets in elements:  A += (z / a_div) * a_stride,  B += (z % b_mod) * b_stride,
 * bias += (z % b_mod) * bias_stride,
 * out += (z / out_div) * out_stride_hi + (z % out_div) * out_stride  (plain batching:
 * out_div = batch, out_stride_hi = 0).
 * mode TSSEP_EPI_F32 / TSSEP_EPI_BF16: out[m * ldo + n] (act: 0 none, 1 tanh).
 * mode TSSEP_EPI_HEAD (TS-VAD/TS-SEP output head, net.py:629-668, :928-986): column
 *   n = q * row_len + f of item z goes to plane p = plane_map[z * n_blocks + q]:
 *   logit[(p * M + m) * row_len + f] = v and mask[...] = sigmoid(v); either may be NULL.  Choose row_len % 8 == 0
 *   (pad every block of B with zero rows: 513 -> 520): with 513-float rows the warp-wide 128-byte stores of the
 *   epilogue start at arbitrary 4-byte offsets and half of the 32-byte sectors they touch are written in part; with
 *   520 every store covers whole sectors (measured on B200: 1.9 -> 3.5 TB/s of output).  The consumers of the mask
 *   take the row pitch (tssep_mask_istft, tssep_activity).
 *   The caller folds the speaker rotation / trial mean into B and bias (K = trials*projs)
 *   and the un-permutation into plane_map.
 * impl: 0 = tcgen05 (product path), 1 = plain SIMT kernel (debug / bisecting only). */
enum { TSSEP_EPI_F32 = 0, TSSEP_EPI_BF16 = 1, TSSEP_EPI_HEAD = 2 };

typedef struct tssep_gemm_desc {
  const uint16_t* A; int64_t lda; int64_t a_stride; int32_t a_div;
  const uint16_t* B; int64_t ldb; int64_t b_stride; int32_t b_mod;
  const float* bias; int64_t bias_stride;
  int64_t M; int32_t N; int32_t K; int32_t batch;
  float alpha; int32_t act; int32_t mode;
  void* out; int64_t ldo; int64_t out_stride; int32_t out_div; int64_t out_stride_hi;
  float* mask; const int32_t* plane_map; int32_t n_blocks; int32_t row_len;
  int32_t impl; int32_t max_ctas; /* max_ctas: 0 = one persistent CTA per SM, else an upper bound */
} tssep_gemm_desc;

int tssep_gemm(const tssep_gemm_desc* desc, tssep_stream_t stream);

/* output_resolution 't' (net.py:653-659): logits (Z, T, n_blocks) f32 -> logit/mask
 * planes (P, T, F) broadcast over frequency; block q of item z goes to plane
 * plane_map[z * n_blocks + q] (speakers un-permuted by the map). */
int tssep_head_expand_t(const float* small, int64_t Z, int64_t T, int n_blocks, int F,
                        const int32_t* plane_map, float* logit, float* mask, tssep_stream_t stream);

/* ------------------------------------------------------------------------
 * (3) BLSTM recurrence (torch.nn.LSTM inside RNNP_packed, rnnp.py:87-95, :143-159)
 * ---------------------------------------------------------------------- */

/* Tensor-memory cluster kernel (the product path for every row count): one cluster per (rows_per_cluster batch
 * rows, direction), both directions concurrently; every CTA owns tiles_per_cta row tiles of 32 hidden units, i.e.
 * clusters of ceil(Up/64) CTAs at 2 tiles (throughput shape) or 2*ceil(Up/64) CTAs at 1 tile (latency shape).  The recurrent weights live in TENSOR
 * MEMORY for the whole sequence as the A operand of tcgen05.mma; per step the tensor core computes
 * P . G_t + W_hh . h_{t-1} (P: scaled permutation matrix, G_t: a TMA box of G in shared memory, h_{t-1}: exchanged
 * through distributed shared memory with st.async), the epilogue warps apply the gates out of TMEM.
 * G    (rows, T, 2, 4, Up) bf16: input projections + both biases, gate order i,f,g,o
 * Wimg from tssep_pack_whh_ts: 2 * C * 2 * (Up/16) * 128 * 8 words, C = ceil(Up/64)
 * H    (rows, T, 2*Up) bf16 out: [h_fwd(Up) | h_bwd(Up)]
 * Up = hidden units rounded up to a multiple of 16; padded units stay 0.  Limits: Up <= 384, and
 * tiles * roundup(Up/2, 32) + 64 + tiles * sub_batches * max(rows_per_cluster / sub_batches, 16) <= 512 tensor-memory
 * columns (Up <= 320 at 2 tiles and 32 rows per sub-batch); clusters of at most 16 CTAs.
 * rows_per_cluster: 8, 16, 32 or 64; tiles_per_cta: 1 or 2; sub_batches: 1, or 2 (2 tiles only; 16, 32 or 64 rows per
 * cluster): the cluster advances two independent halves of its rows in anti-phase -- while the h of one travels through
 * DSMEM and its gate math runs, the tensor pipe works on the other (64 rows per cluster exist only in this form).
 * 0 for any of the three = choose (the shape with the shortest step whose clusters still fit in one wave of
 * co-resident clusters, from a cost table measured on B200).
 * gate_math: 0 = exp-based sigmoid / tanh, 1 = tanh.approx.f32.
 * k_split: 1 = the W_hh . h MMAs start on the half of h that arrives first (2 tiles only); 0 (and -1 = default) = one phase
 * (measured: the second barrier wait + proxy fence cost more than the split hides). */
int tssep_blstm_recurrence_ts(const uint16_t* G, const uint32_t* Wimg, uint16_t* H, int64_t rows, int64_t T, int Up,
                              int rows_per_cluster, int tiles_per_cta, int sub_batches, int gate_math, int k_split,
                              tssep_stream_t stream);
/* Batch rows that fit in ONE wave of co-resident clusters (both directions running) on the current
 * device for the given cluster shape (rows_per_cluster 8 / 16 / 32 / 64, tiles_per_cta 1 / 2, sub_batches 1 / 2);
 * 0 if the shape does not exist or does not fit, < 0 on error. */
int tssep_blstm_recurrence_ts_capacity(int Up, int rows_per_cluster, int tiles_per_cta, int sub_batches);
int tssep_pack_whh_ts(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* Wimg,
                      tssep_stream_t stream);

/* Training step (BASELINE config 5: forward + backward through torch.nn.LSTM, tssep/train/rnnp.py:143-159 under
 * tssep/train/loss.py:219-247).  tssep_blstm_recurrence_train = tssep_blstm_recurrence_ts (two row tiles per CTA) that
 * also stores the gate activations and cell states:
 *   gates  (rows, T, 2, Up) x {i, f, g, o} bf16  (4 x uint16 per hidden unit),  cstate (rows, T, 2, Up) f32.
 * tssep_blstm_recurrence_bwd = the reverse-time recurrence  dh_t = dH_t + W_hh^T da_{t+1},  dc_t = dc_{t+1} f_{t+1} +
 * dh_t o_t (1 - tanh^2 c_t),  da_t = gate derivatives:  dH (rows, T, 2*Up) bf16 in (the gradient of H),
 *   dG (rows, T, 2, Up) x {i, f, g, o} bf16 out: the gradient of the gate pre-activations, [unit][gate] innermost
 *   (a (rows*T, 8*Up) K-major operand for the input-projection dgrad / wgrad with correspondingly ordered weights).
 * Cluster of ceil(Up/64) CTAs per (8 rows, direction); every CTA keeps the transposed rows of W_hh it owns in tensor
 * memory (WTimg from tssep_pack_whh_bwd: 2 * C * ceil(C/2) * 16 * 128 * 8 words), multiplies them with its own da and
 * ships partial sums to the owning CTAs through distributed shared memory.  The gradients of W_hh, W_ih, biases are
 * plain GEMMs / reductions over dG and are left to the caller. */
int tssep_blstm_recurrence_train(const uint16_t* G, const uint32_t* Wimg, uint16_t* H, uint16_t* gates, float* cstate,
                                 int64_t rows, int64_t T, int Up, int rows_per_cluster, int gate_math,
                                 tssep_stream_t stream);
int tssep_blstm_recurrence_bwd(const uint16_t* gates, const float* cstate, const uint16_t* dH, const uint32_t* WTimg,
                               uint16_t* dG, int64_t rows, int64_t T, int Up, tssep_stream_t stream);
int tssep_pack_whh_bwd(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* WTimg,
                       tssep_stream_t stream);

/* Register-resident variant of the same operator (mma.sync, recurrent weights in registers, cluster of up to 8,
 * 8 rows per cluster): accepts f32 G, which the parity tests use to separate the rounding of G from the rest.
 * G    (rows, T, 2, 4, Up) f32 (g_dtype 0) or bf16 (g_dtype 1)
 * Wfrag packed recurrent weights from tssep_pack_whh (2 * Up/4 * Up/16 * 128 u32); Up <= 320.
 * cluster: CTAs per cluster (1,2,4,8), 0 = choose.  fast_math: 0 accurate exp-based gates, 1 tanh.approx. */
int tssep_blstm_recurrence(const void* G, int g_dtype, const uint32_t* Wfrag, uint16_t* H, int64_t rows,
                           int64_t T, int Up, int cluster, int fast_math, tssep_stream_t stream);

/* weight_hh_l0 / weight_hh_l0_reverse (4U, U) f32 -> mma fragment order. */
int tssep_pack_whh(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* Wfrag,
                   tssep_stream_t stream);

/* ------------------------------------------------------------------------
 * (4) Enhancement: mask x mixture STFT, iSTFT overlap-add
 *     Masking.__call__ tssep/train/enhancer.py:73-100; fe.istft model.py:661-664
 * ---------------------------------------------------------------------- */

/* If mask != NULL:  Y[z,k] = X[z] * mask[z,k]  (X (Z,T,F) cfloat with item stride,
 * mask (Z,K,T,F) f32, rows mask_pitch floats apart -- 0 = F; the head GEMM writes 513-float rows at a pitch of 520 so
 * that every row starts on a 32-byte sector) else Y = X viewed as (Z*K, T, F) cfloat.
 * stft_estimate (Z,K,T,F) cfloat and time (Z,K,num_samples) f32 are optional outputs.
 * activity (Z,K,T) f32, optional (needs mask): activity[z,k,t] = mean_f mask[z,k,t,f], the frame activity of the
 * diarization stage (section 5), reduced from the mask rows this kernel reads anyway.
 * synwin (window_length) f32 synthesis window, twiddle as in tssep_stft.
 * size 1024 / shift 256 / window_length 1024 (every shipped config) takes a specialised kernel: 16 x 32
 * register FFT, two speakers per warp, overlap-add accumulator in registers; other geometries a generic
 * shared-memory kernel. */
int tssep_mask_istft(const float* X, int64_t x_item_stride, const float* mask, int64_t mask_pitch, int64_t Z, int n_spk,
                     int64_t T, int size, int shift, int window_length, int fading,
                     const float* synwin, const float* twiddle, float* stft_estimate, float* time,
                     int64_t num_samples, float* activity, tssep_stream_t stream);

/* out[i] = saturate_int16(round_to_nearest_even(x[i] * scale)): separated audio as 16-bit PCM, the format the evaluation
 * driver writes (tssep_b200/eval.py::write_wav); scale = 32767 / peak. */
int tssep_pcm16(const float* x, int64_t n, float scale, int16_t* out, tssep_stream_t stream);

/* Mask-based MVDR beamformer, Souden formulation: TorchBF.__call__ (tssep/train/enhancer.py:140-283).
 * Y (Z, D, T, F) cfloat multi-channel STFT, D <= 8 channels; mask (Z, K, nmask, T, F) f32, nmask 1 (interference
 * weight = 1 - mask) or 2 (target, interference); K * nmask (+1) <= 17.
 * tssep_bf_psd: psd (Z, planes, F, D(D+1)/2) complex float64 = upper triangles of sum_t w Y Y^H, planes = K * nmask
 *   weights in mask order, plus the all-ones weight as the last plane when nmask == 1 (the buffer is zeroed by the call).
 * tssep_bf_mvdr_souden: w (Z, K, F, D) cfloat = phi[:, ref] / max(Re trace(phi), eps), phi = interference^-1 target
 *   (Gaussian elimination with partial pivoting, complex float64).
 * tssep_bf_apply: out (Z, K, T, F) cfloat = sum_d conj(w) Y, times max(mask[:, :, 0], masking_eps) when mask != NULL
 *   (TorchBF(masking=True), enhancer.py:276-281). */
int tssep_bf_psd(const float* Y, const float* mask, int64_t Z, int K, int nmask, int D, int64_t T, int F, double* psd,
                 tssep_stream_t stream);
int tssep_bf_mvdr_souden(const double* psd, int64_t Z, int K, int nmask, int D, int F, int reference_channel, double eps,
                         float* w, tssep_stream_t stream);
int tssep_bf_apply(const float* Y, const float* w, const float* mask, int64_t Z, int K, int nmask, int D, int64_t T, int F,
                   float masking_eps, float* out, tssep_stream_t stream);

/* ------------------------------------------------------------------------
 * (5) Diarization post-processing (no reference implementation; anchors
 *     tssep/util/utils.py:11-129, tssep/train/loss.py:343)
 * ---------------------------------------------------------------------- */

/* activity[n,t] = mean_f mask[n,t,f]; mask rows mask_pitch floats apart (0 = F). */
int tssep_activity(const float* mask, int64_t n, int64_t T, int F, int64_t mask_pitch, float* activity,
                   tssep_stream_t stream);

/* smooth = running median (odd width <= 63, edges replicated); active = smooth > thr. */
int tssep_median_threshold(const float* activity, int64_t n, int64_t T, int width, float threshold,
                           float* smooth, uint8_t* active, tssep_stream_t stream);

/* Maximal runs of active frames -> sample intervals.  segments (n, max_segments, 2)
 * i32, counts (n) i32 (the true number of runs, may exceed max_segments).
 * index_mode 0: run [t0, t1) -> samples [c(t0), c(t1)), c(t) = the first sample nearest to the centre of frame t
 *               (this repo's diarization spec);
 * index_mode 1: istft_vad (tssep/util/utils.py:80-129): [first(t0), last(t1)) with the paderbox index mapping
 *               restated in csrc/postproc.cu (paderbox is absent: parity unpinned). */
int tssep_segments(const uint8_t* active, int64_t n, int64_t T, int window_length, int shift, int fading,
                   int64_t num_samples, int32_t* segments, int32_t* counts, int max_segments, int index_mode,
                   tssep_stream_t stream);

/* stft_vad (tssep/util/utils.py:11-77): sample activity vad (n, num_samples) u8 -> frame activity frames (n, T) u8,
 * T = the STFT frame count of num_samples; a frame is active when a run of active samples [s, e) has
 * frame(s) <= t < frame(e) (same index mapping as index_mode 1 above). */
int tssep_stft_vad(const uint8_t* vad, int64_t n, int64_t num_samples, int window_length, int shift, int fading,
                   int64_t T, uint8_t* frames, tssep_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TSSEP_B200_H_ */

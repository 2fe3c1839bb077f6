/* dcs.h -- C ABI of libdcs.so, the B200 (sm_100a) separation hot path of DeepConvSep.
 *
 * The reference (MTG/DeepConvSep) is pure Python and has no FFI layer; the boundary it
 * exposes for this path is the Python surface `transform.transformFFT` and
 * `examples/<family>/separate_*.py`.  Each entry point below names the reference function(s) it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer would add to the reference.
 *
 * Conventions
 *  - every function returns 0 on success or a negative DCS_E* code; dcs_last_error() returns
 *    a thread-local message for the last failure on the calling thread;
 *  - `d_` pointers are device memory owned by the caller, `h_` pointers are host memory;
 *  - every launch is asynchronous on the given `stream` (a cudaStream_t passed as void*;
 *    NULL = the legacy default stream) unless the name ends in `_host`, which synchronises
 *    the stream before returning;
 *  - a dcs_ctx is bound to one device and owns the (grow-only) workspace of ONE in-flight
 *    pipeline: use one ctx per stream / host thread;
 *  - spectrogram row stride `ldf` is in elements and must be >= F = N/2+1; use
 *    dcs_padded_bins(N) for buffers handed to dcs_separate_spec.
 *  - there is NO CPU fallback: with no usable CUDA device dcs_create fails.
 */
#ifndef DCS_H_
#define DCS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCS_VERSION 100

enum {
  DCS_OK = 0,
  DCS_EINVAL = -1,   /* bad argument / unsupported shape */
  DCS_ECUDA = -2,    /* CUDA runtime error (message has the cudaError string) */
  DCS_ENOMEM = -3,   /* device allocation failed */
  DCS_EMODEL = -4    /* parameter list does not match the architecture */
};

/* network families = the reference's build_ca() variants */
enum {
  DCS_ARCH_DSD = 0,          /* examples/dsd100/separate_dsd.py:172-236 (= hiphopss/separate_hhds.py) */
  DCS_ARCH_IKALA = 1,        /* examples/ikala/separate_ikala.py:172-192 (max-pool) */
  DCS_ARCH_IKALA_NOPOOL = 2, /* examples/ikala/trainCNN.py:66-110 */
  DCS_ARCH_BACH10 = 3,       /* examples/bach10/separate_bach10.py:172-229 */
  DCS_ARCH_BACH10_SCORE = 4, /* examples/bach10_scoreinformed/trainCNNrwc.py:134-193 */
  DCS_ARCH_DSD_ILD = 5       /* examples/dsd100_2ch_ILD/trainCNN_ILD_DSD100.py:66-113 (stereo in, nsrc x 2 out) */
};

/* patch generators */
enum {
  DCS_PATCHER_STANDALONE = 0, /* separate_dsd.py:114-135  (while start+time_context < T, tail dropped) */
  DCS_PATCHER_UTIL = 1        /* util.py:220-248          (while start+overlap < T, zero padded) */
};

typedef struct dcs_ctx dcs_ctx;
typedef struct dcs_stft dcs_stft;
typedef struct dcs_model dcs_model;
typedef struct { float x, y; } dcs_complex; /* layout-compatible with CUDA float2 / numpy complex64 */

int dcs_version(void);
const char* dcs_last_error(void);

/* ---- context ---------------------------------------------------------------------------- */
int dcs_create(int device, dcs_ctx** out);
int dcs_destroy(dcs_ctx* ctx);
/* bytes of device workspace currently held by the ctx */
int64_t dcs_workspace_bytes(const dcs_ctx* ctx);
/* number of kernels this library has launched through `ctx` since creation */
int64_t dcs_launch_count(const dcs_ctx* ctx);

/* Inspection tap (used by the parity tests): while set, every dcs_separate_audio* / *_host call on this ctx
 * also copies the blended masked spectra its inverse STFT consumed -- complex[nplanes][T][ldf],
 * ldf = dcs_padded_bins(N), nplanes = nsrc (x 2 channels for the stereo net) -- to d_S (capacity in
 * elements; the call fails if it is too small).  d_S = NULL switches it off.  These are the tensors
 * `overlapadd_multi(...)/scale * exp(j*phase)` of separate_dsd.py:301-304. */
int dcs_set_spectrum_tap(dcs_ctx* ctx, dcs_complex* d_S, int64_t capacity);
/* Same for the max-pool network (DCS_ARCH_IKALA): the tie bits of MaxPool2DLayer((1,4)) the un-pool
 * (InverseLayer(pool), separate_ikala.py:183,188) routes by -- uint8[T][WP][32], bit r set: position 4*jp+r
 * of the window equals its maximum, channel = last index (30 used), WP = ((F-30)/3+1)/4.  The routing is
 * a discrete decision of the reference's graph; the parity tests adopt the device's where float64 flags
 * the window as ill-conditioned and require agreement elsewhere.  capacity in bytes; NULL = off. */
int dcs_set_pool_tap(dcs_ctx* ctx, uint8_t* d_bits, int64_t capacity);

/* per-stage device timing (CUDA events on the launching stream): enable, run, synchronise the
 * stream, then read.  dcs_profile_read writes up to max_n durations (ms) and the stage names
 * joined by '\n' into names_buf, clears the records and returns the number of records. */
int dcs_profile(dcs_ctx* ctx, int enable);
int dcs_profile_read(dcs_ctx* ctx, char* names_buf, int names_len, float* ms, int max_n);

/* ---- STFT / iSTFT: transform.py:277-396 (stft_norm / istft_norm), :224-274 ----------------- */
/* frame_size N in {256,512,1024,2048,4096}; hop even, hop <= N.
 * `window` = analysis window, `syn_window` = synthesis window (NULL: same), both host
 * double[N] (transformFFT passes window(frameSize) for both, transform.py:273). */
int dcs_stft_plan(dcs_ctx* ctx, int frame_size, int hop, const double* window,
                  const double* syn_window, dcs_stft** out);
int dcs_stft_plan_destroy(dcs_stft* plan);
/* T = ceil(L/hop) + 2            (transform.py:309) */
int64_t dcs_num_frames(int64_t num_samples, int hop);
/* F rounded up to a multiple of 8 */
int64_t dcs_padded_bins(int frame_size);

/* stft_norm (+ the |X|*scale/sqrt(N) of compute_file, transform.py:243-245, fused):
 * d_audio float[L] -> d_X complex[T][ldf] (may be NULL) and d_mag float[T][ldf] (may be NULL),
 * d_mag = mag_scale * |X| / sqrt(N).  Pad columns F..ldf-1 are written as zeros. */
int dcs_stft_forward(dcs_stft* plan, const float* d_audio, int64_t num_samples, dcs_complex* d_X,
                     float* d_mag, float mag_scale, int64_t ldf, void* stream);
/* same analysis, polar output: d_mag = mag_scale*|X|/sqrt(N), d_phase = angle(X)  (compute_file
 * with phase=True, transform.py:243-247) */
int dcs_stft_forward_polar(dcs_stft* plan, const float* d_audio, int64_t num_samples, float* d_mag,
                           float* d_phase, float mag_scale, int64_t ldf, void* stream);
/* istft_norm for nsrc spectrograms d_S complex[nsrc][T][ldf] (source stride src_stride elements)
 * -> d_out float[nsrc][out_stride], the first num_out samples of each (= data[:L],
 * separate_dsd.py:305-306; num_out <= (T-1)*hop + N - N/2).  Imaginary parts of the DC and
 * Nyquist bins are ignored like np.fft.irfft does. */
int dcs_istft(dcs_stft* plan, const dcs_complex* d_S, int nsrc, int64_t num_frames, int64_t ldf,
              int64_t src_stride, float* d_out, int64_t num_out, int64_t out_stride, void* stream);
/* compute_inverse (transform.py:271-273): X = mag_scale*sqrt(N)*mag*exp(j*phase) -> istft_norm */
int dcs_istft_polar(dcs_stft* plan, dcs_ctx* ctx, const float* d_mag, const float* d_phase,
                    float mag_scale, int64_t num_frames, int64_t ldf, float* d_out, int64_t num_out,
                    void* stream);

/* ---- model: load_model + build_ca + set_all_param_values (separate_dsd.py:17-21,246-250) --- */
/* `h_params[i]` = the i-th array of the pickled `lasagne.layers.get_all_param_values(net)` list
 * (float32, C order), `shapes` = nparams x 4 int64 (unused dims = 1), `ndims[i]` = its rank.
 * Weights are re-laid-out for the kernels and uploaded once. */
int dcs_model_create(dcs_ctx* ctx, int arch, int feat_size, int time_context, int nparams,
                     const float* const* h_params, const int64_t* shapes, const int* ndims,
                     dcs_model** out);
int dcs_model_destroy(dcs_model* model);
int dcs_model_nsources(const dcs_model* model);
/* number of patches either patcher cuts from T frames */
int64_t dcs_num_patches(int64_t num_frames, int time_context, int overlap, int patcher);

/* ---- the separation graph on spectrograms ------------------------------------------------ */
/* generate_overlapadd -> predict_function2 (network + soft mask) -> overlapadd[_multi] ->
 * magnitude/scale * exp(j*phase), i.e. separate_dsd.py:292-304, with every per-patch tensor kept
 * on chip: d_mag float[T][ldf] (the scaled magnitude the network sees), d_X complex[T][ldf] (the
 * mixture STFT) -> d_S complex[nsrc][T][ldf] = blended mask_s * X. */
int dcs_separate_spec(dcs_ctx* ctx, dcs_model* model, const float* d_mag, const dcs_complex* d_X,
                      int64_t num_frames, int64_t ldf, int overlap, int patcher, dcs_complex* d_S,
                      int64_t src_stride, void* stream);

/* Score-informed Bach10 (examples/bach10_scoreinformed/trainCNNrwc.py:357-416): the network sees 4
 * input channels in_ch = filter_ch * scaled magnitude, where filter_ch[T][F] are the normalised
 * harmonic masks derived from the scores on the host (LargeDatasetMask2.filterSpec,
 * dataset.py:839-879; deepconvsep_b200/score.py).  d_in: 4 planes [T][ldf], plane stride in_plane. */
int dcs_separate_spec_channels(dcs_ctx* ctx, dcs_model* model, const float* d_in, int64_t in_plane,
                               const dcs_complex* d_X, int64_t num_frames, int64_t ldf, int overlap,
                               int patcher, dcs_complex* d_S, int64_t src_stride, void* stream);
/* whole path on device buffers: d_filters float[4][T][ldf] (ldf = dcs_padded_bins(N), pad columns
 * arbitrary), d_audio float[L] -> d_stems float[4][stem_stride] */
int dcs_separate_audio_score(dcs_ctx* ctx, dcs_model* model, dcs_stft* plan, const float* d_audio,
                             int64_t num_samples, const float* d_filters, float scale_factor, int overlap,
                             int patcher, float* d_stems, int64_t stem_stride, void* stream);

/* ---- building block: the dense layer / im2col-free convolution GEMM ------------------------ */
/* d_C[M][ldc] = act(d_A[M][lda] * h_B[K][ldb] (+ h_bias[N])), fp32 in / fp32 out.  The weight is a
 * HOST array (it is transposed, padded and split for the tensor cores on the fly -- the models
 * do that once at load time).  engine 1: tcgen05 kind::tf32 with the 3xTF32 split (the product
 * path); engine 0: exact-fp32 FFMA kernel (bring-up cross-check).  lasagne DenseLayer
 * (separate_dsd.py:206-221) is this with relu=1.  Synchronises the stream before returning. */
int dcs_gemm_f32(dcs_ctx* ctx, int engine, const float* d_A, int64_t lda, const float* h_B, int64_t ldb,
                 const float* h_bias, float* d_C, int64_t ldc, int M, int N, int K, int relu, void* stream);

/* ---- stereo / ILD variant (examples/dsd100_2ch_ILD/trainCNN_ILD_DSD100.py:299-327) ------------ */
/* d_audio float[2][audio_stride] (left, right; first num_samples valid) ->
 * d_stems float[nsrc*2][stem_stride], plane (s*2 + j) = source s, channel j (`sep_audio[:, s, j]`).
 * One STFT per channel, both scaled magnitudes into the network, per-channel masks normalised over
 * the sources (:183-186), per-channel cross-fade, iSTFT with that channel's mixture phase. */
int dcs_separate_audio_stereo(dcs_ctx* ctx, dcs_model* model, dcs_stft* plan, const float* d_audio,
                              int64_t audio_stride, int64_t num_samples, float scale_factor, int overlap,
                              int patcher, float* d_stems, int64_t stem_stride, void* stream);

/* ---- evaluation: BSS-Eval 3.0 correlation lags ------------------------ */
/* The O(num_samples) part of evaluation/bss_eval/bss_eval_sources.m: the inner products between
 * delayed copies of the true sources (:120-136) and between them and an estimate (:138-145), which
 * the reference takes from FFT cross-correlations and of which only lags |m| < flen (512) are used.
 * For each pair p of float device signals of num_samples samples:
 *     h_out[p][li] = sum_t a_p[t + li - (flen-1)] * b_p[t],   li = 0 .. 2*flen-2      (float64)
 * h_a / h_b: HOST arrays of npairs DEVICE pointers; h_out: HOST double[npairs][2*flen-1].
 * flen <= 512.  Deterministic (fixed summation order).  Synchronises the stream. */
int dcs_xcorr_lags(dcs_ctx* ctx, const float* const* h_a, const float* const* h_b, int npairs,
                   int64_t num_samples, int flen, double* h_out, void* stream);

/* ---- whole train_auto() on device buffers (separate_dsd.py:289-306) ----------------------- */
/* d_audio float[L] mono in [-1,1] -> d_stems float[nsrc][stem_stride] (first L samples valid) */
int dcs_separate_audio(dcs_ctx* ctx, dcs_model* model, dcs_stft* plan, const float* d_audio,
                       int64_t num_samples, float scale_factor, int overlap, int patcher,
                       float* d_stems, int64_t stem_stride, void* stream);
/* same with HOST buffers: H2D copy of the audio, pipeline, D2H copy of the stems, then
 * cudaStreamSynchronize.  Pinned host memory makes the copies asynchronous. */
int dcs_separate_host(dcs_ctx* ctx, dcs_model* model, dcs_stft* plan, const float* h_audio,
                      int64_t num_samples, float scale_factor, int overlap, int patcher,
                      float* h_stems, int64_t stem_stride, void* stream);
/* int16 PCM in / int16 PCM out, the wav-file contract of train_auto (separate_dsd.py:275-287,
 * 307-309): h_pcm int16[L][channels] interleaved; mono = (L+R)/2/32767 (downmix 1) or (L+R)/32767
 * (downmix 2, iKala separate_ikala.py:229) or channel 0 (channels == 1);
 * h_out int16[nsrc][out_stride] = (int16)(stem*32767) (C truncation, no clipping, as astype does) */
int dcs_separate_pcm16_host(dcs_ctx* ctx, dcs_model* model, dcs_stft* plan, const int16_t* h_pcm,
                            int64_t num_samples, int channels, int downmix, float scale_factor,
                            int overlap, int patcher, int16_t* h_out, int64_t out_stride,
                            void* stream);

/* Multi-clip scheduler: `nclips` clips through one context as a pipeline -- H2D of clip i+1 | kernels of clip i | D2H of
 * clip i-1 -- on two internal copy streams and `stream`, with double-buffered int16 staging on the device.  Replaces the
 * reference's process-per-file loop (examples/dsd100/separate_multiple.ipynb cell 3); per clip the contract is that of
 * dcs_separate_pcm16_host.  h_pcm[i]: int16[num_samples[i]][channels] (pinned for real overlap), h_out[i]:
 * int16[nsrc][out_strides[i]].  Order the clips longest first if their lengths differ much (grow-only workspace).
 * Synchronises before returning -- also when it returns an error: every copy in flight has drained, so the host
 * buffers are the caller's again (outputs of clips after the failure are undefined).  Arguments are validated
 * before anything is queued (DCS_EINVAL names the offending clip). */
int dcs_separate_batch_pcm16_host(dcs_ctx* ctx, dcs_model* model, dcs_stft* plan, int nclips,
                                  const int16_t* const* h_pcm, const int64_t* num_samples, int channels, int downmix,
                                  float scale_factor, int overlap, int patcher, int16_t* const* h_out,
                                  const int64_t* out_strides, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCS_H_ */

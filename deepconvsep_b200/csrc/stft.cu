// stft.cu -- framed STFT (K1) and inverse STFT + overlap-add (K4) for sm_100a.
//
// Reference semantics: transform.py:277-335 (stft_norm), :337-396 (istft_norm), :243-247 and
// :271-273 (magnitude / phase conventions of compute_file / compute_inverse).
//
// K1: one CTA of N/8 threads transforms `frames_per_cta` consecutive frames.  The hop-512
//     overlapped windows are read straight from the audio (each sample is re-read N/hop times,
//     served by L1/L2 -- HBM sees the audio once), multiplied by the window in registers and
//     fed to the first radix-4 pass; the spectrum leaves through coalesced float2 / float
//     stores as X[T][ldf] and mag[T][ldf] = scale*|X|/sqrt(N).
// K4: one CTA owns `hops_per_cta` output hops of one source: it inverse-transforms every frame
//     that overlaps them (N/hop - 1 halo frames are recomputed instead of using atomics, so
//     the sum order is the reference's frame order and the result is deterministic),
//     accumulates window * frame in shared memory, divides by sum(window*analysisWindow) and
//     stores the samples once.
#include "common.cuh"
#include "fft.cuh"

namespace dcs {

template <int N>
__global__ void __launch_bounds__(N / 8)
stft_kernel(const float* __restrict__ audio, int64_t L, int hop, const float* __restrict__ win,
            const float2* __restrict__ tw, float2* __restrict__ X, float* __restrict__ mag,
            float* __restrict__ phase, int64_t ldf, int64_t T, float mag_scale, int frames_per_cta) {
  constexpr int N2 = N / 2, T4 = N2 / 4, F = N2 + 1;
  __shared__ __align__(16) float2 bufA[N2];
  __shared__ __align__(16) float2 bufB[N2];
  const int tid = threadIdx.x;

  float2 w[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int idx = 2 * (tid + m * T4);
    w[m] = make_float2(__ldg(win + idx), __ldg(win + idx + 1));
  }
  const int64_t n0 = (int64_t)blockIdx.x * frames_per_cta;
  for (int f = 0; f < frames_per_cta; ++f) {
    const int64_t n = n0 + f;
    if (n >= T) break;  // uniform over the CTA
    const int64_t base = n * hop - N / 2;
    float2 u[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int64_t s = base + 2 * (tid + m * T4);
      const float a0 = (s >= 0 && s < L) ? __ldg(audio + s) : 0.f;
      const float a1 = (s + 1 >= 0 && s + 1 < L) ? __ldg(audio + s + 1) : 0.f;
      u[m] = make_float2(a0 * w[m].x, a1 * w[m].y);
    }
    const float2* Z = fft_forward<N2>(u, bufA, bufB, tw, tid);
    const int64_t row = n * ldf;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = tid + m * T4;
      const float2 xk = real_post<N2>(Z, tw, k);
      if (X) X[row + k] = xk;
      if (mag) mag[row + k] = mag_scale * sqrtf(xk.x * xk.x + xk.y * xk.y);
      if (phase) phase[row + k] = atan2f(xk.y, xk.x);
    }
    if (tid == 0) {  // Nyquist bin
      const float2 z0 = Z[0];
      const float v = z0.x - z0.y;
      if (X) X[row + N2] = make_float2(v, 0.f);
      if (mag) mag[row + N2] = mag_scale * fabsf(v);
      if (phase) phase[row + N2] = atan2f(0.f, v);
    }
    if (tid < ldf - F) {  // pad columns
      if (X) X[row + F + tid] = make_float2(0.f, 0.f);
      if (mag) mag[row + F + tid] = 0.f;
      if (phase) phase[row + F + tid] = 0.f;
    }
    __syncthreads();  // the next frame's first pass overwrites bufA
  }
}

template <int N>
__global__ void __launch_bounds__(N / 8)
istft_kernel(const float2* __restrict__ S, const float* __restrict__ pmag, const float* __restrict__ pphase,
             float polar_scale, int64_t T, int64_t ldf, int64_t src_stride, const float* __restrict__ wsyn,
             const float* __restrict__ w2, const float2* __restrict__ tw, float* __restrict__ out, int64_t Lout,
             int64_t out_stride, int hop, int hops_per_cta) {
  constexpr int N2 = N / 2, T4 = N2 / 4;
  __shared__ __align__(16) float2 bufA[N2];
  __shared__ __align__(16) float2 bufB[N2];
  extern __shared__ float acc[];  // hops_per_cta * hop
  const int tid = threadIdx.x;
  const int src = blockIdx.y;
  const int span = hops_per_cta * hop;
  // output sample i (after the first N/2 samples are dropped, transform.py:390) <-> padded
  // coordinate q = i + N/2.  This CTA owns q in [q_lo, q_lo + span).
  const int64_t q_lo = (int64_t)blockIdx.x * span + N / 2;
  for (int i = tid; i < span; i += T4) acc[i] = 0.f;
  // frames n with n*hop <= q < n*hop + N for some owned q
  int64_t n_lo = (q_lo - N) / hop + 1;  // q_lo >= N/2 > 0; for q_lo < N this is <= 0 -> clamp
  if (q_lo < N) n_lo = 0;
  int64_t n_hi = (q_lo + span - 1) / hop;
  if (n_hi > T - 1) n_hi = T - 1;
  const float inv_n2 = 1.0f / (float)N2;
  __syncthreads();
  for (int64_t n = n_lo; n <= n_hi; ++n) {
    const int64_t row = (int64_t)src * src_stride + n * ldf;
    float2 u[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = tid + m * T4;
      float2 xk, xn;
      if (S) {
        xk = S[row + k];
        xn = S[row + N2 - k];
      } else {
        float sn, cs;
        const float mk = polar_scale * pmag[row + k];
        sincosf(pphase[row + k], &sn, &cs);
        xk = make_float2(mk * cs, mk * sn);
        const float mn = polar_scale * pmag[row + N2 - k];
        sincosf(pphase[row + N2 - k], &sn, &cs);
        xn = make_float2(mn * cs, mn * sn);
      }
      if (k == 0) { xk.y = 0.f; xn.y = 0.f; }  // irfft ignores Im of DC and Nyquist
      u[m] = real_pre_conj(xk, xn, __ldg(tw + k));
    }
    const float2* R = fft_forward<N2>(u, bufA, bufB, tw, tid);
    const int64_t off = n * hop - q_lo;  // local index of sample t=0 of this frame
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int mm = tid + m * T4;
      const float2 r = R[mm];
      const int t0 = 2 * mm;
      const int64_t l0 = off + t0;
      // z = conj(R)/N2 ; x[2m] = Re z, x[2m+1] = Im z
      if (l0 >= 0 && l0 < span) acc[l0] += __ldg(wsyn + t0) * (r.x * inv_n2);
      if (l0 + 1 >= 0 && l0 + 1 < span) acc[l0 + 1] += __ldg(wsyn + t0 + 1) * (-r.y * inv_n2);
    }
    __syncthreads();
  }
  float* o = out + (int64_t)src * out_stride;
  for (int i = tid; i < span; i += T4) {
    const int64_t q = q_lo + i;
    const int64_t oi = q - N / 2;
    if (oi >= Lout) break;
    // normalisation: sum of window*analysisWindow over all frames covering q (transform.py:384-386)
    int64_t a = (q < N) ? 0 : (q - N) / hop + 1;
    int64_t b = q / hop;
    if (b > T - 1) b = T - 1;
    float c = 0.f;
    for (int64_t n = a; n <= b; ++n) c += __ldg(w2 + (q - n * hop));
    if (c == 0.f) c = 1.f;  // transform.py:392
    o[oi] = acc[i] / c;
  }
}

__global__ void pcm_decode_kernel(const int16_t* __restrict__ pcm, int64_t L, int channels, int downmix,
                                  float* __restrict__ audio) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const float maxv = 32767.0f;
  float v;
  if (channels == 1 || downmix == 0) {
    v = (float)pcm[i * channels] / maxv;
  } else {
    // astype(float)/maxv per channel, then (L+R)/2 (separate_dsd.py:282-286) or L+R (iKala)
    const float l = (float)pcm[i * channels] / maxv, r = (float)pcm[i * channels + 1] / maxv;
    v = (downmix == 1) ? (l + r) * 0.5f : (l + r);
  }
  audio[i] = v;
}

__global__ void pcm_encode_kernel(const float* __restrict__ stems, int64_t L, int64_t stem_stride,
                                  int16_t* __restrict__ out, int64_t out_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const int s = blockIdx.y;
  // (audio_out*maxn).astype('int16'): C truncation toward zero, wraps instead of clipping
  const float v = stems[(int64_t)s * stem_stride + i] * 32767.0f;
  out[(int64_t)s * out_stride + i] = (int16_t)(int)v;
}

template <int N>
static int launch_stft_n(dcs_stft* p, const float* d_audio, int64_t L, float2* d_X, float* d_mag, float* d_phase,
                         float mag_scale, int64_t ldf, int64_t T, cudaStream_t st) {
  const int fpc = 8;
  const unsigned grid = (unsigned)ceil_div64(T, fpc);
  stft_kernel<N><<<grid, N / 8, 0, st>>>(d_audio, L, p->hop, p->d_win, p->d_tw, d_X, d_mag, d_phase, ldf, T,
                                          mag_scale / sqrtf((float)N), fpc);
  DCS_CHECK_LAUNCH();
  p->ctx->launches++;
  return DCS_OK;
}

int launch_stft(dcs_stft* p, const float* d_audio, int64_t L, float2* d_X, float* d_mag, float* d_phase,
                float mag_scale, int64_t ldf, cudaStream_t st) {
  const int64_t T = dcs_num_frames(L, p->hop);
  DCS_REQUIRE(ldf >= p->N / 2 + 1, "ldf %lld < F %d", (long long)ldf, p->N / 2 + 1);
  DCS_REQUIRE(ldf - (p->N / 2 + 1) <= 16, "ldf %lld pads more than 16 columns", (long long)ldf);
  if (stft_reg_supported(p->N) && !p->ctx->debug_smem_fft)
    return launch_stft_reg(p, d_audio, L, d_X, d_mag, d_phase, mag_scale, ldf, T, st);
  switch (p->N) {
    case 256: return launch_stft_n<256>(p, d_audio, L, d_X, d_mag, d_phase, mag_scale, ldf, T, st);
    case 512: return launch_stft_n<512>(p, d_audio, L, d_X, d_mag, d_phase, mag_scale, ldf, T, st);
    case 1024: return launch_stft_n<1024>(p, d_audio, L, d_X, d_mag, d_phase, mag_scale, ldf, T, st);
    case 2048: return launch_stft_n<2048>(p, d_audio, L, d_X, d_mag, d_phase, mag_scale, ldf, T, st);
    case 4096: return launch_stft_n<4096>(p, d_audio, L, d_X, d_mag, d_phase, mag_scale, ldf, T, st);
  }
  DCS_REQUIRE(false, "unsupported frame size %d", p->N);
}

template <int N>
static int launch_istft_n(dcs_stft* p, const float2* d_S, const float* d_mag, const float* d_phase, float polar_scale,
                          int nsrc, int64_t T, int64_t ldf, int64_t src_stride, float* d_out, int64_t Lout,
                          int64_t out_stride, cudaStream_t st) {
  // enough hops per CTA to amortise the N/hop-1 halo frames, small enough for many CTAs
  int hpc = 4 * (p->N / p->hop);
  if (hpc < 8) hpc = 8;
  while ((size_t)hpc * p->hop * sizeof(float) > 64 * 1024 && hpc > 1) hpc /= 2;
  const int64_t span = (int64_t)hpc * p->hop;
  const size_t dyn = (size_t)span * sizeof(float);
  dim3 grid((unsigned)ceil_div64(Lout, span), (unsigned)nsrc);
  DCS_TRY(ensure_smem_attr(istft_kernel<N>, 96 * 1024));
  istft_kernel<N><<<grid, N / 8, dyn, st>>>(d_S, d_mag, d_phase, polar_scale, T, ldf, src_stride, p->d_wsyn, p->d_w2,
                                            p->d_tw, d_out, Lout, out_stride, p->hop, hpc);
  DCS_CHECK_LAUNCH();
  p->ctx->launches++;
  return DCS_OK;
}

int launch_istft(dcs_stft* p, const float2* d_S, const float* d_mag, const float* d_phase, float polar_scale,
                 int nsrc, int64_t T, int64_t ldf, int64_t src_stride, float* d_out, int64_t Lout,
                 int64_t out_stride, cudaStream_t st) {
  DCS_REQUIRE(Lout <= (T - 1) * p->hop + p->N - p->N / 2, "num_out %lld exceeds the istft length", (long long)Lout);
  if (Lout <= 0 || nsrc <= 0) return DCS_OK;
  if (d_S && istft_reg_supported(p, d_out, out_stride) && !p->ctx->debug_smem_fft && ldf % 2 == 0 &&
      src_stride % 2 == 0 && ((uintptr_t)d_S % 16 == 0) && ldf >= (p->N / 2 + 2) / 2 * 2)
    return launch_istft_reg(p, d_S, nsrc, T, ldf, src_stride, d_out, Lout, out_stride, st);
#define DCS_ISTFT_CASE(NN) \
  case NN: return launch_istft_n<NN>(p, d_S, d_mag, d_phase, polar_scale, nsrc, T, ldf, src_stride, d_out, Lout, out_stride, st);
  switch (p->N) {
    DCS_ISTFT_CASE(256)
    DCS_ISTFT_CASE(512)
    DCS_ISTFT_CASE(1024)
    DCS_ISTFT_CASE(2048)
    DCS_ISTFT_CASE(4096)
  }
#undef DCS_ISTFT_CASE
  DCS_REQUIRE(false, "unsupported frame size %d", p->N);
}

int launch_pcm_decode(dcs_ctx* ctx, const int16_t* d_pcm, int64_t L, int channels, int downmix, float* d_audio,
                      cudaStream_t st) {
  if (L <= 0) return DCS_OK;
  pcm_decode_kernel<<<(unsigned)ceil_div64(L, 256), 256, 0, st>>>(d_pcm, L, channels, downmix, d_audio);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

int launch_pcm_encode(dcs_ctx* ctx, const float* d_stems, int64_t L, int nsrc, int64_t stem_stride, int16_t* d_out,
                      int64_t out_stride, cudaStream_t st) {
  if (L <= 0) return DCS_OK;
  dim3 grid((unsigned)ceil_div64(L, 256), (unsigned)nsrc);
  pcm_encode_kernel<<<grid, 256, 0, st>>>(d_stems, L, stem_stride, d_out, out_stride);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

}  // namespace dcs

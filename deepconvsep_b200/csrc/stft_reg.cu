// stft_reg.cu -- STFT (K1) and iSTFT + overlap-add (K4) with the register-resident FFT of
// fft_reg.cuh, for frame sizes 1024 and 2048 (the DSD100 / iKala configurations).  Same reference
// semantics as stft.cu (transform.py:277-396); that file's shared-memory Stockham kernels remain
// the path for the other frame sizes / hops.
//
// K1: one group of T = N/64 threads (a warp for N = 2048) owns a frame: hop-overlapped, windowed
//     samples go from global memory straight into registers, the spectrum is staged once in the
//     group's scratch so that the real-FFT split and the X / mag rows leave fully coalesced.
// K4: a group owns `hops_per_group` consecutive output hops of one source and walks the frames
//     that overlap them IN FRAME ORDER, keeping the overlap-add window (N samples) in registers:
//     after frame n is added, the oldest hop is complete -> divided by sum(win*syn_win), stored
//     coalesced, and the window shifts.  No shared accumulator, no atomics: the summation order
//     is the reference's (ascending frame index) and the result is run-to-run deterministic.
//     N/hop - 1 halo frames are recomputed at the start of each span.
#include <stdlib.h>
#include "common.cuh"
#include "fft_reg.cuh"

namespace dcs {

constexpr int REG_THREADS = 128;
// K4 runs ONE CTA of 12 warps per SM (same 12 resident warps as three 4-warp CTAs) so that the
// twiddle table and the synthesis window -- per-thread constants re-read for every frame -- fit in
// shared memory next to the 12 scratch + row buffers: with ~200 KB of shared memory carved out, the
// L1 is down to ~28 KB and those 24 KB of tables kept missing to L2 (long-scoreboard stalls were 44 %
// of the kernel's stall samples, profiles/r1_notes.md).
constexpr int ISTFT_THREADS = 384;

template <int T>
__global__ void __launch_bounds__(REG_THREADS)
stft_reg_kernel(const float* __restrict__ audio, int64_t L, int hop, const float* __restrict__ win,
                const float2* __restrict__ tw, float2* __restrict__ X, float* __restrict__ mag,
                float* __restrict__ phase, int64_t ldf, int64_t nframes, float mag_scale, int frames_per_group) {
  using G = FftGroup<T>;
  constexpr int N2 = G::N2, N = 2 * N2, F = N2 + 1, GPC = REG_THREADS / T;
  extern __shared__ __align__(16) float2 scratch[];
  const int tid = threadIdx.x, b = tid % T, gl = tid / T;
  float2* scr = scratch + gl * G::SCRATCH;
  // shared copies of the twiddle table and of the analysis window (per-thread constants of every frame)
  float2* stw = scratch + GPC * G::SCRATCH;        // tw[0..N2): the real-FFT split twiddles
  float2* stw2 = stw + N2;                         // inter-stage twiddles, lane-contiguous (FftGroup::fill_tw2)
  float2* swin = stw2 + N2;
  for (int i = tid; i < N2; i += REG_THREADS) stw[i] = __ldg(tw + i);
  G::fill_tw2(stw2, tw, tid, REG_THREADS);
  for (int i = tid; i < N2; i += REG_THREADS) swin[i] = __ldg(reinterpret_cast<const float2*>(win) + i);
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * GPC + gl;
  const bool audio_aligned8 = (reinterpret_cast<uintptr_t>(audio) & 7) == 0;
  const int64_t warp_first = ((int64_t)blockIdx.x * GPC + (tid / 32) * (32 / T)) * frames_per_group;
  for (int i = 0; i < frames_per_group; ++i) {
    if (warp_first + i >= nframes) break;  // warp-uniform: even the warp's first group is past the end
    const int64_t n = g * frames_per_group + i;
    const bool valid = n < nframes;
    const int64_t base = n * hop - N / 2;
    float2 v[32];
    // frames that lie inside the clip (all but N/hop at either end) take 8-byte loads with no bounds
    // arithmetic: base is even, so the pairs are aligned whenever the buffer is
    if (valid && base >= 0 && base + N <= L && audio_aligned8) {
      const float2* __restrict__ ap = reinterpret_cast<const float2*>(audio + base);
#pragma unroll
      for (int a = 0; a < 32; ++a) {
        const int idx = a * T + b;
        const float2 w = swin[idx], x = __ldg(ap + idx);
        v[a] = make_float2(x.x * w.x, x.y * w.y);
      }
    } else {
#pragma unroll
      for (int a = 0; a < 32; ++a) {
        const int idx = a * T + b;
        const int64_t s = base + 2 * idx;
        const float2 w = swin[idx];
        const float x0 = (valid && s >= 0 && s < L) ? __ldg(audio + s) : 0.f;
        const float x1 = (valid && s + 1 >= 0 && s + 1 < L) ? __ldg(audio + s + 1) : 0.f;
        v[a] = make_float2(x0 * w.x, x1 * w.y);
      }
    }
    G::template forward<true>(v, scr, stw2, b);
#pragma unroll
    for (int q = 0; q < G::Q; ++q)
#pragma unroll
      for (int kb = 0; kb < T; ++kb) scr[(b + T * q) + 32 * kb] = v[q * T + kb];
    __syncwarp();
    if (valid) {
      const int64_t row = n * ldf;
#pragma unroll 4
      for (int m = 0; m < N2 / T; ++m) {
        const int k = b + T * m;
        const float2 xk = real_post_shared<N2>(scr, stw, k);
        if (X) X[row + k] = xk;
        if (mag) mag[row + k] = mag_scale * sqrtf(xk.x * xk.x + xk.y * xk.y);
        if (phase) phase[row + k] = atan2f(xk.y, xk.x);
      }
      if (b == 0) {  // Nyquist bin
        const float2 z0 = scr[0];
        const float vv = z0.x - z0.y;
        if (X) X[row + N2] = make_float2(vv, 0.f);
        if (mag) mag[row + N2] = mag_scale * fabsf(vv);
        if (phase) phase[row + N2] = atan2f(0.f, vv);
      }
      if (b < ldf - F) {  // pad columns
        if (X) X[row + F + b] = make_float2(0.f, 0.f);
        if (mag) mag[row + F + b] = 0.f;
        if (phase) phase[row + F + b] = 0.f;
      }
    }
    __syncwarp();
  }
}

template <int T, int HS>  // HS = hop / 64: window slots (32 float2 each) per hop
__global__ void __launch_bounds__(ISTFT_THREADS, 1)
istft_reg_kernel(const float2* __restrict__ S, int64_t nframes, int64_t ldf, int64_t src_stride,
                 const float* __restrict__ wsyn, const float* __restrict__ w2, const float2* __restrict__ tw,
                 float* __restrict__ out, int64_t Lout, int64_t out_stride, int hops_per_group, int64_t num_hops,
                 int64_t groups_per_src, int64_t total_groups) {
  using G = FftGroup<T>;
  constexpr int N2 = G::N2, N = 2 * N2, hop = 64 * HS, R = N / hop, C0 = (N / 2) / hop, GPC = ISTFT_THREADS / T;
  extern __shared__ __align__(16) float2 scratch[];
  const int tid = threadIdx.x, b = tid % T, gl = tid / T;
  float2* scr = scratch + gl * G::SCRATCH;
  int64_t gi = (int64_t)blockIdx.x * GPC + gl;
  const bool active = gi < total_groups;
  if (!active) gi = total_groups - 1;  // keeps the warp convergent; nothing is stored
  const int src = (int)(gi / groups_per_src);
  const int64_t h0 = (gi % groups_per_src) * hops_per_group;
  const float inv_n2 = 1.0f / (float)N2;
  float2 acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = make_float2(0.f, 0.f);
  float* o = out + (int64_t)src * out_stride;
  // 1 / sum_r win*syn_win for the interior of the clip (every hop sees the same R frames)
  float2 cinv[G::Q * HS];
#pragma unroll
  for (int q = 0; q < G::Q; ++q)
#pragma unroll
    for (int kb = 0; kb < HS; ++kb) {
      const int e = 2 * ((b + T * q) + 32 * kb);
      float c0 = 0.f, c1 = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float2 ww = __ldg(reinterpret_cast<const float2*>(w2 + e + r * hop));
        c0 += ww.x;
        c1 += ww.y;
      }
      cinv[q * HS + kb] = make_float2(c0 == 0.f ? 1.f : 1.f / c0, c1 == 0.f ? 1.f : 1.f / c1);
    }

  // The spectrum row of the NEXT frame is fetched into shared memory with cp.async while the
  // current frame is transformed and overlap-added: HBM latency is off the critical path.
  constexpr int ROWP = (N2 + 1 + 7) / 8 * 8;          // float2 per staged row
  constexpr int CHUNKS = (N2 + 2) / 2;                // 16-byte pieces covering bins 0..N2
  float2* srow = scratch + GPC * G::SCRATCH + gl * ROWP;
  // shared copies of the twiddle table (N entries, exp(-2 pi i j / N)) and of the synthesis window (as pairs)
  float2* stw = scratch + GPC * (G::SCRATCH + ROWP);   // tw[0..N2): the real-FFT merge twiddles
  float2* stw2 = stw + N2;                              // inter-stage twiddles, lane-contiguous (FftGroup::fill_tw2)
  float2* swsyn = stw2 + N2;
  for (int i = tid; i < N2; i += ISTFT_THREADS) stw[i] = __ldg(tw + i);
  G::fill_tw2(stw2, tw, tid, ISTFT_THREADS);
  for (int i = tid; i < N2; i += ISTFT_THREADS) swsyn[i] = __ldg(reinterpret_cast<const float2*>(wsyn) + i);
  __syncthreads();
  auto prefetch = [&](int64_t nn) {
    if (nn >= 0 && nn < nframes) {
      const float2* rowp = S + (int64_t)src * src_stride + nn * ldf;
      for (int c = b; c < CHUNKS; c += T) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(srow + 2 * c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rowp + 2 * c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t n_first = h0 + C0 - R + 1, n_last = h0 + hops_per_group - 1 + C0;
  prefetch(n_first);
  for (int64_t n = n_first; n <= n_last; ++n) {
    const bool fvalid = n >= 0 && n < nframes;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    float2 v[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) {
      const int idx = a * T + b;
      float2 xk = make_float2(0.f, 0.f), xn = xk;
      if (fvalid) {
        xk = srow[idx];
        xn = srow[N2 - idx];
      }
      if (idx == 0) { xk.y = 0.f; xn.y = 0.f; }  // irfft ignores Im of DC and Nyquist
      v[a] = real_pre_conj(xk, xn, stw[idx]);
    }
    __syncwarp();                    // every lane has consumed the row
    if (n < n_last) prefetch(n + 1);
    G::template forward<true>(v, scr, stw2, b);
    // z = conj(V)/N2: samples 2n', 2n'+1 of the frame, n' = (b + T q) + 32 kb  <->  v[q*T + kb]
#pragma unroll
    for (int q = 0; q < G::Q; ++q)
#pragma unroll
      for (int kb = 0; kb < T; ++kb) {
        const int np = (b + T * q) + 32 * kb;
        const float2 w = swsyn[np];
        const float2 r = v[q * T + kb];
        acc[q * T + kb].x = fmaf(w.x, r.x * inv_n2, acc[q * T + kb].x);
        acc[q * T + kb].y = fmaf(w.y, -r.y * inv_n2, acc[q * T + kb].y);
      }
    // the oldest hop of the window is complete: padded samples [n*hop, (n+1)*hop) -> output hop n - C0
    const int64_t h = n - C0;
    if (active && h >= h0 && h < num_hops) {
      const bool interior = (n - (R - 1) >= 0) && (n < nframes);   // all R overlapping frames exist
#pragma unroll
      for (int q = 0; q < G::Q; ++q)
#pragma unroll
        for (int kb = 0; kb < HS; ++kb) {
          const int e = 2 * ((b + T * q) + 32 * kb);  // sample offset inside the hop
          float2 ic = cinv[q * HS + kb];
          if (!interior) {   // clip edges: sum win*syn_win over the frames that exist (transform.py:384-392)
            float c0 = 0.f, c1 = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const int64_t nf = n - r;
              if (nf >= 0 && nf < nframes) {
                const float2 ww = __ldg(reinterpret_cast<const float2*>(w2 + e + r * hop));
                c0 += ww.x;
                c1 += ww.y;
              }
            }
            ic = make_float2(c0 == 0.f ? 1.f : 1.f / c0, c1 == 0.f ? 1.f : 1.f / c1);
          }
          const int64_t oi = h * hop + e;
          const float2 a2 = acc[q * T + kb];
          if (oi + 1 < Lout) {
            *reinterpret_cast<float2*>(o + oi) = make_float2(a2.x * ic.x, a2.y * ic.y);
          } else if (oi < Lout) {
            o[oi] = a2.x * ic.x;
          }
        }
    }
    // shift the window by one hop
#pragma unroll
    for (int q = 0; q < G::Q; ++q) {
#pragma unroll
      for (int kb = 0; kb < T - HS; ++kb) acc[q * T + kb] = acc[q * T + kb + HS];
#pragma unroll
      for (int kb = T - HS; kb < T; ++kb) acc[q * T + kb] = make_float2(0.f, 0.f);
    }
  }
}

bool stft_reg_supported(int N) { return N == 1024 || N == 2048; }

int launch_stft_reg(dcs_stft* p, const float* d_audio, int64_t L, float2* d_X, float* d_mag, float* d_phase,
                    float mag_scale, int64_t ldf, int64_t nframes, cudaStream_t st) {
  const int fpg = 4;
  const float ms = mag_scale / sqrtf((float)p->N);
  if (p->N == 2048) {
    using G = FftGroup<32>;
    const int gpc = REG_THREADS / 32;
    const size_t smem = (gpc * G::SCRATCH + 3 * G::N2) * sizeof(float2);   // scratch + twiddles + window
    DCS_TRY(ensure_smem_attr(stft_reg_kernel<32>, (int)smem));
    const unsigned grid = (unsigned)ceil_div64(nframes, (int64_t)gpc * fpg);
    stft_reg_kernel<32><<<grid, REG_THREADS, smem, st>>>(
        d_audio, L, p->hop, p->d_win, p->d_tw, d_X, d_mag, d_phase, ldf, nframes, ms, fpg);
  } else {
    using G = FftGroup<16>;
    const int gpc = REG_THREADS / 16;
    const size_t smem = (gpc * G::SCRATCH + 3 * G::N2) * sizeof(float2);
    DCS_TRY(ensure_smem_attr(stft_reg_kernel<16>, (int)smem));
    const unsigned grid = (unsigned)ceil_div64(nframes, (int64_t)gpc * fpg);
    stft_reg_kernel<16><<<grid, REG_THREADS, smem, st>>>(
        d_audio, L, p->hop, p->d_win, p->d_tw, d_X, d_mag, d_phase, ldf, nframes, ms, fpg);
  }
  DCS_CHECK_LAUNCH();
  p->ctx->launches++;
  return DCS_OK;
}

bool istft_reg_supported(const dcs_stft* p, const float* d_out, int64_t out_stride) {
  // (the caller also guarantees 16-byte aligned spectrum rows: ldf % 2 == 0, see launch_istft)
  return (p->N == 1024 || p->N == 2048) && (p->hop == 512 || p->hop == 256) && ((uintptr_t)d_out % 8 == 0) &&
         out_stride % 2 == 0;
}

template <int T, int HS>
static int launch_istft_reg_t(dcs_stft* p, const float2* d_S, int nsrc, int64_t nframes, int64_t ldf, int64_t src_stride,
                              float* d_out, int64_t Lout, int64_t out_stride, cudaStream_t st) {
  using G = FftGroup<T>;
  constexpr int GPC = ISTFT_THREADS / T;
  const int hop = 64 * HS;
  const int64_t num_hops = ceil_div64(Lout, hop);
  // hops per group: as long as possible (each group recomputes N/hop-1 halo frames) while every SM
  // still gets its 12 resident warps: ONE full wave of equal-sized groups (DCS_DEBUG_ISTFT_WAVES=2:
  // the earlier two-wave split)
  static const int waves = [] { const char* e = getenv("DCS_DEBUG_ISTFT_WAVES"); return e && e[0] == '2' ? 2 : 1; }();
  const int64_t target_groups = (int64_t)p->ctx->num_sms * 12 * (32 / T) * waves;
  int64_t hpg = ceil_div64((int64_t)nsrc * num_hops, target_groups);
  if (hpg < 12) hpg = 12;
  if (hpg > 64) hpg = 64;
  const int64_t groups_per_src = ceil_div64(num_hops, hpg);
  const int64_t total = groups_per_src * nsrc;
  const unsigned grid = (unsigned)ceil_div64(total, GPC);
  constexpr int ROWP = (G::N2 + 1 + 7) / 8 * 8;
  const size_t smem = ((size_t)GPC * (G::SCRATCH + ROWP) + 2 * G::N2 + G::N2) * sizeof(float2);   // + twiddles + window
  DCS_TRY(ensure_smem_attr(istft_reg_kernel<T, HS>, (int)smem));
  istft_reg_kernel<T, HS><<<grid, ISTFT_THREADS, smem, st>>>(
      d_S, nframes, ldf, src_stride, p->d_wsyn, p->d_w2, p->d_tw, d_out, Lout, out_stride, (int)hpg, num_hops,
      groups_per_src, total);
  DCS_CHECK_LAUNCH();
  p->ctx->launches++;
  return DCS_OK;
}

int launch_istft_reg(dcs_stft* p, const float2* d_S, int nsrc, int64_t nframes, int64_t ldf, int64_t src_stride,
                     float* d_out, int64_t Lout, int64_t out_stride, cudaStream_t st) {
  if (p->N == 2048 && p->hop == 512) return launch_istft_reg_t<32, 8>(p, d_S, nsrc, nframes, ldf, src_stride, d_out, Lout, out_stride, st);
  if (p->N == 2048 && p->hop == 256) return launch_istft_reg_t<32, 4>(p, d_S, nsrc, nframes, ldf, src_stride, d_out, Lout, out_stride, st);
  if (p->N == 1024 && p->hop == 512) return launch_istft_reg_t<16, 8>(p, d_S, nsrc, nframes, ldf, src_stride, d_out, Lout, out_stride, st);
  if (p->N == 1024 && p->hop == 256) return launch_istft_reg_t<16, 4>(p, d_S, nsrc, nframes, ldf, src_stride, d_out, Lout, out_stride, st);
  DCS_REQUIRE(false, "istft_reg: unsupported frame size / hop");
}

}  // namespace dcs

// fft.cuh -- shared-memory Stockham autosort FFT (radix-4 passes, one trailing radix-2 pass when
// log2 is odd) for the framed STFT / iSTFT kernels.  One FFT of N2 complex points is computed
// by N2/4 cooperating threads; the first pass takes its inputs from registers (straight from
// the global-memory loads), every later pass ping-pongs between two shared buffers.
//
// A real frame of N samples is transformed as an N/2-point complex FFT of z[m] = x[2m] + i x[2m+1]
// plus the split/merge step (`real_post` / `real_pre`), so the shared buffers hold N/2 float2.
#pragma once
#include <cuda_runtime.h>

namespace dcs {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// 4-point DFT, natural order in and out
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  float2 jd = make_float2(d13.y, -d13.x);  // (-i) * d13
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = cadd(d02, jd);
  a3 = csub(d02, jd);
}

// Forward DFT of N2 points.  u[m] = x[tid + m*N2/4] on entry.  `tw[q] = exp(-2*pi*i*q/(2*N2))`,
// 2*N2 entries.  Returns the shared buffer holding the spectrum in natural order; all threads
// have passed a __syncthreads() after the last write.
template <int N2>
__device__ __forceinline__ float2* fft_forward(float2 (&u)[4], float2* bufA, float2* bufB,
                                               const float2* __restrict__ tw, int tid) {
  constexpr int T4 = N2 / 4;
  constexpr int TWN = 2 * N2;
  bfly4(u[0], u[1], u[2], u[3]);
  {
    float4* d = reinterpret_cast<float4*>(bufA + 4 * tid);
    d[0] = make_float4(u[0].x, u[0].y, u[1].x, u[1].y);
    d[1] = make_float4(u[2].x, u[2].y, u[3].x, u[3].y);
  }
  __syncthreads();
  float2* src = bufA;
  float2* dst = bufB;
#pragma unroll
  for (int p = 4; p < N2; p *= 4) {
    if (N2 / p >= 4) {
      const int k = tid & (p - 1);
      const int j = ((tid - k) << 2) + k;
      const int s = TWN / (4 * p);
      float2 a0 = src[tid], a1 = src[tid + T4], a2 = src[tid + 2 * T4], a3 = src[tid + 3 * T4];
      a1 = cmul(a1, __ldg(tw + k * s));
      a2 = cmul(a2, __ldg(tw + 2 * k * s));
      a3 = cmul(a3, __ldg(tw + 3 * k * s));
      bfly4(a0, a1, a2, a3);
      dst[j] = a0;
      dst[j + p] = a1;
      dst[j + 2 * p] = a2;
      dst[j + 3 * p] = a3;
    } else {  // one radix-2 pass, two butterflies per thread
      constexpr int T2 = N2 / 2;
      const int s = TWN / (2 * p);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = tid + r * T4;
        const int k = i & (p - 1);
        const int j = ((i - k) << 1) + k;
        float2 a0 = src[i];
        float2 a1 = cmul(src[i + T2], __ldg(tw + k * s));
        dst[j] = cadd(a0, a1);
        dst[j + p] = csub(a0, a1);
      }
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

// split step of the real FFT: spectrum bin k (0 <= k < N2) of the N = 2*N2 real sequence from
// the N2-point spectrum Z of the packed sequence.  Bin N2 is  (Re Z0 - Im Z0, 0).
template <int N2>
__device__ __forceinline__ float2 real_post(const float2* Z, const float2* __restrict__ tw, int k) {
  const float2 zk = Z[k];
  const float2 zn = Z[(N2 - k) & (N2 - 1)];
  const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));   // (Zk + conj Zn)/2
  const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));  // (Zk - conj Zn)/(2i)
  return cadd(e, cmul(__ldg(tw + k), o));
}

// same with the twiddle table in shared memory
template <int N2>
__device__ __forceinline__ float2 real_post_shared(const float2* Z, const float2* tw, int k) {
  const float2 zk = Z[k];
  const float2 zn = Z[(N2 - k) & (N2 - 1)];
  const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
  const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
  return cadd(e, cmul(tw[k], o));
}

// merge step of the inverse real FFT: packed spectrum bin k (0 <= k < N2) from the half
// spectrum S (bins 0..N2); returns conj(Z[k]) so that a FORWARD transform followed by a
// conjugation gives the inverse.  xk = S[k], xn = S[N2-k], imaginary parts of bins 0 and N2
// must already be zeroed by the caller (np.fft.irfft ignores them).
__device__ __forceinline__ float2 real_pre_conj(float2 xk, float2 xn, float2 twk) {
  const float2 e = make_float2(0.5f * (xk.x + xn.x), 0.5f * (xk.y - xn.y));  // (Xk + conj Xn)/2
  const float2 d = make_float2(0.5f * (xk.x - xn.x), 0.5f * (xk.y + xn.y));  // (Xk - conj Xn)/2
  const float2 o = cmul(d, make_float2(twk.x, -twk.y));                      // * exp(+2 pi i k/N)
  // Z = E + i*O ; return conj(Z)
  return make_float2(e.x - o.y, -(e.y + o.x));
}

}  // namespace dcs

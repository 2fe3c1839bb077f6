// common.cuh -- shared helpers for libdcs (sm_100a only).
#pragma once
#include <cuda.h>           // CUtensorMap (types only: the encoder is fetched through the runtime)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <string.h>
#include "../../include/dcs.h"

namespace dcs {

void set_error(const char* fmt, ...);

#define DCS_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      dcs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return DCS_ECUDA;                                                                 \
    }                                                                                   \
  } while (0)

#define DCS_CHECK_LAUNCH()  DCS_CUDA(cudaGetLastError())

#define DCS_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      dcs::set_error(__VA_ARGS__);             \
      return DCS_EINVAL;                       \
    }                                          \
  } while (0)

#define DCS_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != DCS_OK) return _r; \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: remembered per
// (kernel, current device), thread-safe (api.cu).  Call with the ctx's device current.
int ensure_smem_attr_impl(const void* kernel, int bytes);
template <typename K> inline int ensure_smem_attr(K kernel, int bytes) { return ensure_smem_attr_impl((const void*)kernel, bytes); }

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grow-only device buffer; newly allocated memory is zero-filled
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, cudaStream_t stream, bool* grew = nullptr);
  void release();
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace dcs

struct dcs_prof_rec {
  const char* name;
  cudaEvent_t e0, e1;
};

struct dcs_ctx {
  int device = 0;
  int num_sms = 148;
  int64_t launches = 0;
  bool prof_on = false;
  bool debug_simt_gemm = false;
  int tc_acc_mode = 2;
  int tc_debug = 0;
  int tma_mode = 1;      // DCS_DEBUG_TMA: 0 = register-staged GEMM everywhere, 2 = rewrite the high plane
  int tma_stages = 2;    // DCS_DEBUG_TMA_STAGES (64-wide tiles: 2 or 4)
  bool tma_wide = false; // DCS_DEBUG_TMA_WIDE: 128-wide tiles (one CTA per SM) for N > 64
  int tma_mask = 63;     // DCS_DEBUG_TMA_MASK: which operand views may take the TMA kernel (32: the window view of conv1)
  bool tma_sync = false; // DCS_DEBUG_TMA_SYNC: synchronise after each TMA GEMM
  int tma_prefetch = 0;  // DCS_DEBUG_TMA_PREFETCH: activation boxes prefetched into L2 ahead of the stage ring (measured: no gain)
  int tma_probe = 0;     // DCS_DEBUG_TMA_PROBE: timing experiments on the TMA GEMM (results are wrong)
  int tma_persist_wide = 1;   // DCS_DEBUG_TMA_PERSIST_WIDE: persistent kernel also for N > 64 (Bach10 decoder dense 2.86 -> 2.57 ms, DSD100 neutral)
  int tma_atm = 1;       // DCS_DEBUG_TMA_ATM: persistent GEMM takes its A operand from tensor memory (0: from shared memory)
  int tma_persist = 8;   // DCS_DEBUG_TMA_PERSIST: persistent GEMM for short-K tiles when tiles >= this x SMs (0 = never)
  bool debug_smem_fft = false;
  std::vector<dcs_prof_rec> prof;
  // workspace of one in-flight pipeline
  dcs::DevBuf audio, X, mag, S, stems, pcm_in, pcm_out;
  // multi-clip scheduler (dcs_separate_batch_pcm16_host): copy streams, double-buffered staging, hand-over events
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_dec[2] = {nullptr, nullptr}, ev_enc[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  dcs::DevBuf pcm_in2[2], pcm_out2[2];
  dcs::DevBuf net[12];
  uint64_t net_sig[12] = {0};   // layout signature of what each net[] buffer currently holds
  float2* tap = nullptr;        // dcs_set_spectrum_tap: copy of the masked spectra the iSTFT consumed
  int64_t tap_cap = 0;
  uint8_t* pool_tap = nullptr;  // dcs_set_pool_tap: copy of the max-pool tie bits of the forward pass
  int64_t pool_tap_cap = 0;
  int64_t workspace_bytes() const;
};

struct dcs_stft {
  dcs_ctx* ctx;
  int N, hop;
  float* d_win;    // analysis window  float[N]
  float* d_wsyn;   // synthesis window float[N]
  float* d_w2;     // wsyn * win       float[N]
  float2* d_tw;    // exp(-2*pi*i*q/N), q < N
};

namespace dcs {
// records a pair of CUDA events around a pipeline stage when profiling is enabled on the ctx
struct ProfScope {
  dcs_ctx* c; cudaStream_t st; int idx;
  ProfScope(dcs_ctx* ctx, const char* name, cudaStream_t s) : c(ctx), st(s), idx(-1) {
    if (!c || !c->prof_on) return;
    dcs_prof_rec r; r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    c->prof.push_back(r);
    idx = (int)c->prof.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof[idx].e1, st); }
};
}  // namespace dcs

namespace dcs {
// weight matrix prepared for the tensor-core path: K-major, zero padded, split for 3xTF32
struct TcWeight {
  float* hi = nullptr;   // [Np][Kp], K-major, zero padded
  float* lo = nullptr;   // = hi + Np*Kp (one allocation: the TMA kernel sees the planes stacked)
  int K = 0, N = 0, Kp = 0, Np = 0;
  // tensor maps over the stacked planes, box = {32, 64, 128} rows x 32 floats; made on first use
  mutable CUtensorMap tmap[3];
  mutable bool tmap_ok[3] = {false, false, false};
};
}  // namespace dcs

// ---- model (device-resident network)
struct dcs_sconv {   // strided-conv1 families: iKala (pool / no pool), Bach10
  int nch, sw1, J, pool, WP, kh2, kw2, h2, w2, HP, WPP, ndec, nfc, rule;
  dcs::TcWeight tW[8];                 // 0 conv1, 1 conv2, 2 fc, 3 convT2, 4.. decoder dense layers
  dcs::TcWeight tW1p;                  // conv1 with a 32-row K pitch per input channel (copy-engine window view)
  float *b1, *b2, *bfc, *bdec[4], *bout, *Wsc;
};
struct dcs_model {
  dcs_ctx* ctx;
  int arch, F, tc, nsrc;
  // DSD dims
  int C1, C2, kh2, h2, nfc, ndec;
  int C1p, C2p;   // channel pitch of the activation buffers (multiple of 4 floats)
  int nch = 1;    // input channels (2: stereo / ILD net); W1t is [nch][C1][ldw], bout [nch][4]
  int64_t ldw;
  std::vector<float*> dev;  // owned device arrays
  float *W1f, *b1, *W2c, *b2, *Wfc, *bfc, *Wdec, *bdec, *Wt2, *W1t, *bout;
  // tensor-core copies of the GEMM weights (K-major, 3xTF32 split)
  dcs::TcWeight tW1f, tW2c, tWfc, tWdec, tWt2;
  dcs_sconv sc;
};
namespace dcs {
int upload(const std::vector<float>& h, float** d);
int model_create_sconv(dcs_model* m, int nparams, const float* const* hp, const int64_t* shp, const int* nd);
// d_in: nch planes [T][ldf] (plane stride in_plane elements; nch = 1: the scaled magnitude)
int sconv_forward(dcs_ctx* ctx, dcs_model* m, const float* d_in, int64_t in_plane, const float2* d_X, int64_t T, int64_t ldf,
                  int overlap, int patcher, float2* d_S, int64_t src_stride, cudaStream_t st);
// (re)zero a workspace buffer whenever what it holds changes layout: zero padding is relied upon
int ensure_layout(dcs_ctx* ctx, int idx, size_t bytes, uint64_t sig, cudaStream_t st);
}  // namespace dcs

// ---- kernel launchers (each returns a DCS_* code) ---------------------------------------------
namespace dcs {

int launch_stft(dcs_stft* plan, const float* d_audio, int64_t L, float2* d_X, float* d_mag,
                float* d_phase, float mag_scale, int64_t ldf, cudaStream_t st);
int launch_istft(dcs_stft* plan, const float2* d_S, const float* d_mag, const float* d_phase,
                 float polar_scale, int nsrc, int64_t T, int64_t ldf, int64_t src_stride, float* d_out,
                 int64_t Lout, int64_t out_stride, cudaStream_t st);

bool stft_reg_supported(int N);
int launch_stft_reg(dcs_stft* plan, const float* d_audio, int64_t L, float2* d_X, float* d_mag, float* d_phase,
                    float mag_scale, int64_t ldf, int64_t nframes, cudaStream_t st);
bool istft_reg_supported(const dcs_stft* plan, const float* d_out, int64_t out_stride);
int launch_istft_reg(dcs_stft* plan, const float2* d_S, int nsrc, int64_t nframes, int64_t ldf, int64_t src_stride,
                     float* d_out, int64_t Lout, int64_t out_stride, cudaStream_t st);

// generic strided-operand GEMM  C = act(A*B + bias)
struct GemmDesc {
  const float* A; const float* B; const float* bias; float* C;
  int M, N, K;
  int a_valid_rows;          // rows >= a_valid_rows of A read as zeros
  // A row offset = (m / m_inner) * a_so + ((m % m_inner) / m_inner2) * a_si + (m % m_inner2) * a_s2
  int m_inner; int64_t a_so, a_si;
  int m_inner2; int64_t a_s2;
  int k_seg; int64_t k_ss;           // A col offset  = (k / k_seg) * k_ss + (k % k_seg)
  // optional "window" view for the copy engine (win_stride > 0; conv1 of the 30-channel nets with a 16-byte position
  // stride): row m = (frame, position j) is the 32-float window starting at frame * a_so + j * win_stride of plane
  // k / 32 (k_seg = 32, planes k_ss apart); windows overlap and run 2 floats past the 30-tap filter (zero weight rows)
  int win_stride;
  int64_t ldb;                       // B[k][n] at B + k*ldb + n
  int cm_inner; int64_t c_so, c_si;  // C row offset (same three-level form)
  int cm_inner2; int64_t c_s2;
  int n_seg; int64_t n_ss, c_col0;   // C col offset  = c_col0 + (n / n_seg) * n_ss + (n % n_seg)
  int relu;
  // optional K clipping for transposed convolutions on a zero-padded operand: rows are grouped by
  // output position u = m / kc_rows; only taps q with kc_pad <= u + q < kc_pad + kc_n touch
  // non-zero input, so a tile skips the k-blocks outside [kc_unit*q_lo, kc_unit*(q_hi+1)).
  // kc_rows = 0 disables it.  (Pure optimisation: the skipped products are exact zeros.)
  int kc_rows, kc_unit, kc_pad, kc_n, kc_taps;
  // optional frame-major C rows for the transposed conv2 of the DSD100 net (fm_step > 0): GEMM row
  // m = (p, k, d) = (frame in patch, patch, decoder), m = p * cm_inner + k * fm_ndec + d, is stored at
  // row ((t * fm_slots + j) * fm_ndec + d) of C with t = k * fm_step + p the mixture frame and
  // j = k - k_lo(t) the patch slot (k_lo(t) = first patch covering t) -- the order in which the fused
  // mask kernel consumes them, so that a group of frames is ONE contiguous box.  Rows with t >= fm_T
  // are dropped.  c_so is then the row pitch.
  int fm_step, fm_tc, fm_T, fm_slots, fm_ndec;
};
// element offset of row m of C (without the column part); -1: the row is not stored
__host__ __device__ __forceinline__ int64_t gemm_c_row_offset(const GemmDesc& d, int m) {
  if (d.fm_step > 0) {
    const int p = m / d.cm_inner, kd = m - p * d.cm_inner;
    const int k = kd / d.fm_ndec, dd = kd - k * d.fm_ndec;
    const int t = k * d.fm_step + p;
    if (t >= d.fm_T) return -1;
    const int lo = t - d.fm_tc + 1;
    const int k_lo = lo > 0 ? (lo + d.fm_step - 1) / d.fm_step : 0;
    return ((int64_t)((int64_t)t * d.fm_slots + (k - k_lo)) * d.fm_ndec + dd) * d.c_so + d.c_col0;
  }
  return (int64_t)(m / d.cm_inner) * d.c_so + (int64_t)((m % d.cm_inner) / d.cm_inner2) * d.c_si +
         (int64_t)(m % d.cm_inner2) * d.c_s2 + d.c_col0;
}
GemmDesc gemm_plain(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C,
                    int64_t ldc, int M, int N, int K, int relu);
int launch_gemm(dcs_ctx* ctx, const GemmDesc& d, cudaStream_t st);

int tc_weight_create(const float* B_rowmajor, int64_t ldb, int K, int N, TcWeight* out);
void tc_weight_destroy(TcWeight* w);
void tc_weight_encode_maps(TcWeight* w);
int launch_gemm_tc(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st);
constexpr int DCS_TMA_FALLBACK = 1;   // launch_gemm_tma: "use the register-staged kernel" (not an error)
bool gemm_tma_eligible(const GemmDesc& d, int mask);
// fp32 [rows][cols] tensor (row pitch in bytes), box = box_rows x 32 floats, 128-byte swizzle, zero fill
int tma_encode_2d_f32(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_rows);
// 3-D fp32 tensor (d0 innermost = 32-float boxes, SWIZZLE_128B), box = {32, box_rows, 1}
int tma_encode_3d_f32(CUtensorMap* map, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box_rows);
int launch_splitk_reduce(dcs_ctx* ctx, const GemmDesc& d, const float* partial, int ldp, int k_splits, cudaStream_t st);
int launch_gemm_tma(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st);

struct DsdMaskArgs {
  const float* G;      // [P][3][tc][ldg]  decoder activations after the transposed conv2
  int ldg;
  const float* W1t;    // [50][ldw]  W1t[c][b] = conv1.W[c,0,0,F-1-b]
  int ldw;
  const float* bout;   // [4]
  const float2* X;     // [T][ldf]
  float2* S;           // [4][T][ldf]
  int64_t ldf, src_stride;
  int T, P, tc, overlap, F;
  int ndec;            // 3: DSD100 (4th output = decoder 2, all-zero bins get 1/4); 4: one decoder per source,
                       //    all-zero bins get 0 (stereo / ILD net, one launch per channel)
};
int launch_dsd_mask(dcs_ctx* ctx, const DsdMaskArgs& a, cudaStream_t st);
// strided-conv1 families (iKala / Bach10): K3s arguments
struct SconvMaskArgs {
  int arch;
  const float* G;       // [P*ndec][tc][WP or J][32] decoder activations after the transposed conv2
  const uint8_t* tie;   // [T'][WP][32] max-pool tie bits of the forward pass (pooled nets only)
  const float* W;       // float4 [ND][32]: W[dd][f][r] = conv1.W[f][0][0][KW-1-r-STRIDE*dd]
  const float* bout;    // [nsrc]
  const float2* X;      // [T][ldf]
  float2* S;            // [nsrc][T][ldf]
  int64_t ldf, src_stride;
  int T, P, tc, overlap, F, J, WP;
};
int launch_pool4(dcs_ctx* ctx, const float* H1, float* Hp, uint8_t* tie, int64_t rows, int J, int WP, cudaStream_t st);
int launch_sconv_mask(dcs_ctx* ctx, const SconvMaskArgs& a, cudaStream_t st);
bool sconv_mask_tc_supported(const SconvMaskArgs& a);
int launch_sconv_mask_tc(dcs_ctx* ctx, const SconvMaskArgs& a, cudaStream_t st);   // tcgen05 (sconv_tc.cu)
int launch_channel_mul(dcs_ctx* ctx, const float* mag, const float* filt, float* out, int64_t plane, int nch, cudaStream_t st);
bool dsd_mask_tc_supported(const DsdMaskArgs& a);
int launch_dsd_mask_tc(dcs_ctx* ctx, const DsdMaskArgs& a, cudaStream_t st);

int launch_xcorr_lags(dcs_ctx* ctx, const float* const* h_a, const float* const* h_b, int npairs, int64_t L, int flen,
                      double* h_out, cudaStream_t st);
int launch_pcm_decode(dcs_ctx* ctx, const int16_t* d_pcm, int64_t L, int channels, int downmix, float* d_audio,
                      cudaStream_t st);
int launch_pcm_encode(dcs_ctx* ctx, const float* d_stems, int64_t L, int nsrc, int64_t stem_stride, int16_t* d_out,
                      int64_t out_stride, cudaStream_t st);

}  // namespace dcs

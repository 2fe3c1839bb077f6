// api.cu -- C ABI of libdcs.so (include/dcs.h): context / workspace, STFT plans, model
// re-layout + upload, and the host-side orchestration of the separation pipeline.
#include <stdarg.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <utility>
#include <vector>
#include "common.cuh"

namespace dcs {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int DevBuf::ensure(size_t bytes, cudaStream_t stream, bool* grew) {
  if (grew) *grew = false;
  if (bytes <= cap) return DCS_OK;
  if (p) {
    DCS_CUDA(cudaStreamSynchronize(stream));
    DCS_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
  }
  size_t want = (bytes + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    p = nullptr;
    set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    return DCS_ENOMEM;
  }
  cap = want;
  DCS_CUDA(cudaMemsetAsync(p, 0, want, stream));
  if (grew) *grew = true;
  return DCS_OK;
}

void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

int ensure_smem_attr_impl(const void* kernel, int bytes) {
  int dev = 0;
  DCS_CUDA(cudaGetDevice(&dev));
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  std::lock_guard<std::mutex> g(mu);
  const auto key = std::make_pair(kernel, dev);
  const auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return DCS_OK;
  DCS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[key] = bytes;
  return DCS_OK;
}

int upload(const std::vector<float>& h, float** d) {
  DCS_CUDA(cudaMalloc((void**)d, h.size() * sizeof(float)));
  DCS_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return DCS_OK;
}

}  // namespace dcs

using namespace dcs;

int64_t dcs_ctx::workspace_bytes() const {
  size_t s = audio.cap + X.cap + mag.cap + S.cap + stems.cap + pcm_in.cap + pcm_out.cap + pcm_in2[0].cap + pcm_in2[1].cap +
             pcm_out2[0].cap + pcm_out2[1].cap;
  for (const auto& b : net) s += b.cap;
  return (int64_t)s;
}

extern "C" {

int dcs_version(void) { return DCS_VERSION; }
const char* dcs_last_error(void) { return g_err; }

int dcs_create(int device, dcs_ctx** out) {
  DCS_REQUIRE(out != nullptr, "dcs_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("dcs_create: no usable CUDA device (%s); libdcs has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return DCS_ECUDA;
  }
  DCS_REQUIRE(device >= 0 && device < n, "dcs_create: device %d out of range (%d devices)", device, n);
  DCS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DCS_CUDA(cudaGetDeviceProperties(&prop, device));
  DCS_REQUIRE(prop.major >= 10, "dcs_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
              prop.major, prop.minor);
  dcs_ctx* c = new dcs_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  const char* dbg = getenv("DCS_DEBUG_SIMT_GEMM");
  c->debug_simt_gemm = dbg && dbg[0] == '1';
  const char* sk = getenv("DCS_DEBUG_TC_SKIP");
  if (sk && sk[0] >= '0' && sk[0] <= '7') c->tc_debug = sk[0] - '0';
  const char* sf = getenv("DCS_DEBUG_SMEM_FFT");
  c->debug_smem_fft = sf && sf[0] == '1';
  const char* am = getenv("DCS_DEBUG_TC_ACC");
  if (am && am[0] >= '0' && am[0] <= '2') c->tc_acc_mode = am[0] - '0';
  const char* tm = getenv("DCS_DEBUG_TMA");
  if (tm && tm[0] >= '0' && tm[0] <= '2') c->tma_mode = tm[0] - '0';
  const char* ts = getenv("DCS_DEBUG_TMA_STAGES");
  if (ts && (ts[0] == '2' || ts[0] == '4')) c->tma_stages = ts[0] - '0';
  const char* tn = getenv("DCS_DEBUG_TMA_WIDE");
  c->tma_wide = tn && tn[0] == '1';
  const char* tk = getenv("DCS_DEBUG_TMA_MASK");
  if (tk && tk[0] >= '0' && tk[0] <= '9') c->tma_mask = atoi(tk);
  const char* tp = getenv("DCS_DEBUG_TMA_PREFETCH");
  if (tp && tp[0] >= '0' && tp[0] <= '9') c->tma_prefetch = atoi(tp);
  const char* tr = getenv("DCS_DEBUG_TMA_PROBE");
  if (tr && tr[0] >= '0' && tr[0] <= '9') c->tma_probe = atoi(tr);
  const char* tw = getenv("DCS_DEBUG_TMA_PERSIST_WIDE");
  if (tw && tw[0] >= '0' && tw[0] <= '9') c->tma_persist_wide = atoi(tw);
  const char* tu = getenv("DCS_DEBUG_TMA_ATM");
  if (tu && tu[0] >= '0' && tu[0] <= '9') c->tma_atm = atoi(tu);
  const char* tq = getenv("DCS_DEBUG_TMA_PERSIST");
  if (tq && tq[0] >= '0' && tq[0] <= '9') c->tma_persist = atoi(tq);
  const char* ty = getenv("DCS_DEBUG_TMA_SYNC");
  c->tma_sync = ty && ty[0] == '1';
  *out = c;
  return DCS_OK;
}

int dcs_destroy(dcs_ctx* c) {
  if (!c) return DCS_OK;
  cudaSetDevice(c->device);
  c->audio.release(); c->X.release(); c->mag.release(); c->S.release(); c->stems.release();
  c->pcm_in.release(); c->pcm_out.release();
  for (auto& b : c->net) b.release();
  for (int i = 0; i < 2; ++i) {
    c->pcm_in2[i].release(); c->pcm_out2[i].release();
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_dec[i]) cudaEventDestroy(c->ev_dec[i]);
    if (c->ev_enc[i]) cudaEventDestroy(c->ev_enc[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
  }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  delete c;
  return DCS_OK;
}

int64_t dcs_workspace_bytes(const dcs_ctx* c) { return c ? c->workspace_bytes() : 0; }
int64_t dcs_launch_count(const dcs_ctx* c) { return c ? c->launches : 0; }

int dcs_set_spectrum_tap(dcs_ctx* c, dcs_complex* d_S, int64_t capacity) {
  DCS_REQUIRE(c != nullptr && capacity >= 0, "dcs_set_spectrum_tap: bad argument");
  c->tap = (float2*)d_S;
  c->tap_cap = d_S ? capacity : 0;
  return DCS_OK;
}

int dcs_set_pool_tap(dcs_ctx* c, uint8_t* d_bits, int64_t capacity) {
  DCS_REQUIRE(c != nullptr && capacity >= 0, "dcs_set_pool_tap: bad argument");
  c->pool_tap = d_bits;
  c->pool_tap_cap = d_bits ? capacity : 0;
  return DCS_OK;
}

// the blended masked spectra the inverse STFT of this call consumed -> the caller's tap buffer
static int copy_tap(dcs_ctx* c, const float2* S, int64_t elems, cudaStream_t st) {
  if (!c->tap) return DCS_OK;
  DCS_REQUIRE(c->tap_cap >= elems, "spectrum tap holds %lld elements, this call produced %lld", (long long)c->tap_cap, (long long)elems);
  DCS_CUDA(cudaMemcpyAsync(c->tap, S, (size_t)elems * sizeof(float2), cudaMemcpyDeviceToDevice, st));
  return DCS_OK;
}

int dcs_profile(dcs_ctx* c, int enable) {
  DCS_REQUIRE(c != nullptr, "dcs_profile: NULL ctx");
  c->prof_on = enable != 0;
  return DCS_OK;
}

int dcs_profile_read(dcs_ctx* c, char* names_buf, int names_len, float* ms, int max_n) {
  DCS_REQUIRE(c && names_buf && ms && names_len > 0, "dcs_profile_read: bad argument");
  int n = 0;
  size_t pos = 0;
  names_buf[0] = 0;
  for (auto& r : c->prof) {
    float t = 0.f;
    if (n < max_n && cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      const size_t len = strlen(r.name);
      if (pos + len + 2 < (size_t)names_len) {
        memcpy(names_buf + pos, r.name, len);
        pos += len;
        names_buf[pos++] = '\n';
        names_buf[pos] = 0;
        ms[n++] = t;
      }
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  c->prof.clear();
  return n;
}

// ------------------------------------------------------------------------------------ STFT plan
int64_t dcs_num_frames(int64_t L, int hop) { return (L + hop - 1) / hop + 2; }
int64_t dcs_padded_bins(int N) { return ((int64_t)N / 2 + 1 + 7) / 8 * 8; }

int dcs_stft_plan(dcs_ctx* ctx, int N, int hop, const double* window, const double* syn_window, dcs_stft** out) {
  DCS_REQUIRE(ctx && window && out, "dcs_stft_plan: NULL argument");
  DCS_REQUIRE(N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096, "frame size %d not in {256..4096}", N);
  DCS_REQUIRE(hop > 0 && hop <= N && hop % 2 == 0, "hop %d must be even and in (0, %d]", hop, N);
  DCS_CUDA(cudaSetDevice(ctx->device));
  if (!syn_window) syn_window = window;
  std::vector<float> w(N), ws(N), w2(N), tw(2 * (size_t)N);
  for (int i = 0; i < N; ++i) {
    w[i] = (float)window[i];
    ws[i] = (float)syn_window[i];
    w2[i] = (float)(window[i] * syn_window[i]);
    const double a = -2.0 * M_PI * (double)i / (double)N;
    tw[2 * i] = (float)cos(a);
    tw[2 * i + 1] = (float)sin(a);
  }
  dcs_stft* p = new dcs_stft();
  p->ctx = ctx; p->N = N; p->hop = hop;
  p->d_win = p->d_wsyn = p->d_w2 = nullptr; p->d_tw = nullptr;
  int r = upload(w, &p->d_win);
  if (r == DCS_OK) r = upload(ws, &p->d_wsyn);
  if (r == DCS_OK) r = upload(w2, &p->d_w2);
  if (r == DCS_OK) r = upload(tw, (float**)&p->d_tw);
  if (r != DCS_OK) { dcs_stft_plan_destroy(p); return r; }
  *out = p;
  return DCS_OK;
}

int dcs_stft_plan_destroy(dcs_stft* p) {
  if (!p) return DCS_OK;
  cudaFree(p->d_win); cudaFree(p->d_wsyn); cudaFree(p->d_w2); cudaFree(p->d_tw);
  delete p;
  return DCS_OK;
}

int dcs_stft_forward(dcs_stft* p, const float* d_audio, int64_t L, dcs_complex* d_X, float* d_mag, float mag_scale,
                     int64_t ldf, void* stream) {
  DCS_REQUIRE(p && d_audio && L > 0, "dcs_stft_forward: bad argument");
  DCS_CUDA(cudaSetDevice(p->ctx->device));
  return launch_stft(p, d_audio, L, (float2*)d_X, d_mag, nullptr, mag_scale, ldf, (cudaStream_t)stream);
}

int dcs_stft_forward_polar(dcs_stft* p, const float* d_audio, int64_t L, float* d_mag, float* d_phase, float mag_scale,
                           int64_t ldf, void* stream) {
  DCS_REQUIRE(p && d_audio && L > 0, "dcs_stft_forward_polar: bad argument");
  DCS_CUDA(cudaSetDevice(p->ctx->device));
  return launch_stft(p, d_audio, L, nullptr, d_mag, d_phase, mag_scale, ldf, (cudaStream_t)stream);
}

int dcs_istft(dcs_stft* p, const dcs_complex* d_S, int nsrc, int64_t T, int64_t ldf, int64_t src_stride, float* d_out,
              int64_t Lout, int64_t out_stride, void* stream) {
  DCS_REQUIRE(p && d_S && d_out && T > 0, "dcs_istft: bad argument");
  DCS_CUDA(cudaSetDevice(p->ctx->device));
  return launch_istft(p, (const float2*)d_S, nullptr, nullptr, 1.f, nsrc, T, ldf, src_stride, d_out, Lout, out_stride,
                      (cudaStream_t)stream);
}

int dcs_istft_polar(dcs_stft* p, dcs_ctx* ctx, const float* d_mag, const float* d_phase, float mag_scale, int64_t T,
                    int64_t ldf, float* d_out, int64_t Lout, void* stream) {
  DCS_REQUIRE(p && d_mag && d_phase && d_out && T > 0, "dcs_istft_polar: bad argument");
  DCS_CUDA(cudaSetDevice(p->ctx->device));
  (void)ctx;
  return launch_istft(p, nullptr, d_mag, d_phase, mag_scale * sqrtf((float)p->N), 1, T, ldf, 0, d_out, Lout, Lout,
                      (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ model
int64_t dcs_num_patches(int64_t T, int tc, int overlap, int patcher) {
  const int64_t step = tc - overlap;
  if (step <= 0) return 0;
  const int64_t lim = patcher == DCS_PATCHER_UTIL ? overlap : tc;
  if (T <= lim) return 0;
  return (T - lim - 1) / step + 1;
}

int dcs_model_nsources(const dcs_model* m) { return m ? m->nsrc : 0; }

int dcs_model_destroy(dcs_model* m) {
  if (!m) return DCS_OK;
  for (float* d : m->dev) cudaFree(d);
  for (auto& w : m->sc.tW) tc_weight_destroy(&w);
  tc_weight_destroy(&m->sc.tW1p);
  tc_weight_destroy(&m->tW1f); tc_weight_destroy(&m->tW2c); tc_weight_destroy(&m->tWfc);
  tc_weight_destroy(&m->tWdec); tc_weight_destroy(&m->tWt2);
  delete m;
  return DCS_OK;
}

static bool shape_is(const int64_t* s, int nd, int want_nd, int64_t a, int64_t b = 1, int64_t c = 1, int64_t d = 1) {
  return nd == want_nd && s[0] == a && s[1] == b && s[2] == c && s[3] == d;
}

static int model_create_dsd(dcs_model* m, int nparams, const float* const* hp, const int64_t* shp, const int* nd) {
  const int F = m->F, tc = m->tc;
  // DSD100 / hiphopss: 1 input channel, 128-wide bottleneck, 3 decoders feeding 4 outputs
  // (separate_dsd.py:196-231); stereo / ILD: 2 input channels, 256-wide bottleneck, one decoder per
  // source, 4 x 2 outputs ordered (source, channel) (trainCNN_ILD_DSD100.py:88-108)
  const bool ild = m->arch == DCS_ARCH_DSD_ILD;
  const int nch = ild ? 2 : 1, ndec = ild ? 4 : 3, nfc = ild ? 256 : 128, nout = 4 * nch;
  const int C1 = 50, C2 = 50, kh2 = tc / 2, h2 = tc - kh2 + 1, flat = C2 * h2;
  m->C1 = C1; m->C2 = C2; m->kh2 = kh2; m->h2 = h2; m->nfc = nfc; m->ndec = ndec; m->nsrc = 4; m->nch = nch;
  const int want = 8 + 2 * ndec + 1;
  if (nparams != want) { set_error("this DSD model needs %d parameter arrays, got %d (SURVEY.md App. A.4)", want, nparams); return DCS_EMODEL; }
  bool ok = shape_is(shp + 0, nd[0], 4, C1, nch, 1, F) && shape_is(shp + 4, nd[1], 1, C1) &&
            shape_is(shp + 8, nd[2], 1, C1) && shape_is(shp + 12, nd[3], 4, C2, C1, kh2, 1) &&
            shape_is(shp + 16, nd[4], 1, C2) && shape_is(shp + 20, nd[5], 1, C2) &&
            shape_is(shp + 24, nd[6], 2, flat, nfc) && shape_is(shp + 28, nd[7], 1, nfc) &&
            shape_is(shp + 4 * (want - 1), nd[want - 1], 1, nout);
  for (int d = 0; d < ndec; ++d)
    ok = ok && shape_is(shp + 4 * (8 + 2 * d), nd[8 + 2 * d], 2, nfc, flat) && shape_is(shp + 4 * (9 + 2 * d), nd[9 + 2 * d], 1, flat);
  if (!ok) { set_error("DSD parameter shapes do not match feat_size=%d time_context=%d", F, tc); return DCS_EMODEL; }
  const int64_t ldf = dcs_padded_bins(2 * (F - 1));
  m->ldw = ldf;
  const float *W1 = hp[0], *W2 = hp[3], *Wfc = hp[6];
  // channel pitch of the activation buffers: 52 floats, so that every row and every time step starts
  // on a 16-byte boundary (what the TMA-fed GEMM needs); the K index of each weight follows the
  // same pitch with zero rows at the two pad channels
  const int C1p = (C1 + 3) / 4 * 4, C2p = (C2 + 3) / 4 * 4, flatp = C2p * h2;
  m->C1p = C1p; m->C2p = C2p;
  // W1f: conv1 as a GEMM weight, K index = ch * F + bin; W1t: its transpose per input channel for K3
  std::vector<float> W1f((size_t)nch * ldf * C1, 0.f), W1t((size_t)nch * C1 * ldf, 0.f), b1(C1), W2c((size_t)kh2 * C1p * C2, 0.f),
      Wt2((size_t)kh2 * C2p * C1, 0.f), b2(C2), Wfcp((size_t)flatp * nfc, 0.f), Wdec((size_t)nfc * ndec * flatp, 0.f),
      bdec((size_t)ndec * flatp, 0.f);
  for (int f = 0; f < C1; ++f)
    for (int ch = 0; ch < nch; ++ch)
      for (int b = 0; b < F; ++b) {
        const float v = W1[((size_t)f * nch + ch) * F + (F - 1 - b)];  // flip_filters
        W1f[((size_t)ch * F + b) * C1 + f] = v;
        W1t[((size_t)ch * C1 + f) * ldf + b] = v;
      }
  for (int f = 0; f < C1; ++f) b1[f] = hp[1][f] + hp[2][f];
  for (int f = 0; f < C2; ++f) b2[f] = hp[4][f] + hp[5][f];
  for (int f = 0; f < C2; ++f)
    for (int c = 0; c < C1; ++c)
      for (int q = 0; q < kh2; ++q) {
        const float v = W2[((size_t)f * C1 + c) * kh2 + q];
        W2c[((size_t)(kh2 - 1 - q) * C1p + c) * C2 + f] = v;  // conv2 forward, tap p' = kh2-1-q
        Wt2[((size_t)q * C2p + f) * C1 + c] = v;              // InverseLayer(conv2)
      }
  for (int f = 0; f < C2; ++f)
    for (int i = 0; i < h2; ++i)
      memcpy(&Wfcp[((size_t)i * C2p + f) * nfc], &Wfc[((size_t)f * h2 + i) * nfc], nfc * sizeof(float));
  for (int d = 0; d < ndec; ++d) {
    const float* Wd = hp[8 + 2 * d];
    const float* bd = hp[9 + 2 * d];
    for (int f = 0; f < C2; ++f)
      for (int i = 0; i < h2; ++i) {
        const size_t col = (size_t)d * flatp + (size_t)i * C2p + f;
        bdec[col] = bd[f * h2 + i];
        for (int o = 0; o < nfc; ++o) Wdec[(size_t)o * ndec * flatp + col] = Wd[(size_t)o * flat + f * h2 + i];
      }
  }
  // output bias per channel: bout[ch][s] = b[(s, ch)] (one K3 launch per channel)
  std::vector<float> bout((size_t)nch * 4), bfc(hp[7], hp[7] + nfc);
  for (int ch = 0; ch < nch; ++ch)
    for (int sidx = 0; sidx < 4; ++sidx) bout[(size_t)ch * 4 + sidx] = hp[want - 1][sidx * nch + ch];
  struct { const std::vector<float>* h; float** d; } ups[] = {
      {&W1f, &m->W1f}, {&b1, &m->b1}, {&W2c, &m->W2c}, {&b2, &m->b2}, {&Wfcp, &m->Wfc}, {&bfc, &m->bfc},
      {&Wdec, &m->Wdec}, {&bdec, &m->bdec}, {&Wt2, &m->Wt2}, {&W1t, &m->W1t}, {&bout, &m->bout}};
  for (auto& u : ups) {
    DCS_TRY(upload(*u.h, u.d));
    m->dev.push_back(*u.d);
  }
  DCS_TRY(tc_weight_create(W1f.data(), C1, nch * F, C1, &m->tW1f));
  DCS_TRY(tc_weight_create(W2c.data(), C2, kh2 * C1p, C2, &m->tW2c));
  DCS_TRY(tc_weight_create(Wfcp.data(), nfc, flatp, nfc, &m->tWfc));
  DCS_TRY(tc_weight_create(Wdec.data(), ndec * flatp, nfc, ndec * flatp, &m->tWdec));
  DCS_TRY(tc_weight_create(Wt2.data(), C1, kh2 * C2p, C1, &m->tWt2));
  return DCS_OK;
}

int dcs_model_create(dcs_ctx* ctx, int arch, int feat_size, int time_context, int nparams, const float* const* h_params,
                     const int64_t* shapes, const int* ndims, dcs_model** out) {
  DCS_REQUIRE(ctx && h_params && shapes && ndims && out, "dcs_model_create: NULL argument");
  DCS_REQUIRE(feat_size >= 3 && ((feat_size - 1) & (feat_size - 2)) == 0, "feat_size %d is not 2^k+1", feat_size);
  DCS_REQUIRE(time_context >= 4 && time_context <= 64, "time_context %d out of range", time_context);
  DCS_CUDA(cudaSetDevice(ctx->device));
  dcs_model* m = new dcs_model();
  m->ctx = ctx; m->arch = arch; m->F = feat_size; m->tc = time_context;
  int r;
  switch (arch) {
    case DCS_ARCH_DSD:
    case DCS_ARCH_DSD_ILD: r = model_create_dsd(m, nparams, h_params, shapes, ndims); break;
    case DCS_ARCH_IKALA:
    case DCS_ARCH_IKALA_NOPOOL:
    case DCS_ARCH_BACH10:
    case DCS_ARCH_BACH10_SCORE: r = model_create_sconv(m, nparams, h_params, shapes, ndims); break;
    default:
      set_error("dcs_model_create: architecture %d has no CUDA path yet", arch);
      r = DCS_EINVAL;
  }
  if (r != DCS_OK) { dcs_model_destroy(m); return r; }
  *out = m;
  return DCS_OK;
}

// ------------------------------------------------------------------------------------ pipeline
// every dense contraction goes to the tcgen05 kernel; DCS_DEBUG_SIMT_GEMM=1 (read once in
// dcs_create) routes them to the FFMA kernel instead -- a bring-up aid, not a fallback.
static int run_gemm(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st) {
  return ctx->debug_simt_gemm ? launch_gemm(ctx, d, st) : launch_gemm_tc(ctx, d, w, st);
}

// d_mag / d_X: nch planes [T][ldf] (plane strides mag_plane / x_plane; nch = 1: the DSD100 net);
// d_S: masked spectra, plane (s * nch + ch) at (s * nch + ch) * src_stride
static int dsd_forward(dcs_ctx* ctx, dcs_model* m, const float* d_mag, int64_t mag_plane, const float2* d_X, int64_t x_plane,
                       int64_t T, int64_t ldf, int overlap, int patcher, float2* d_S, int64_t src_stride, cudaStream_t st) {
  const int tc = m->tc, step = tc - overlap, C1 = m->C1, C2 = m->C2, kh2 = m->kh2, h2 = m->h2, nfc = m->nfc;
  const int C1p = m->C1p, C2p = m->C2p;   // channel pitch of H1 / H2 / the padded decoder activations
  const int nch = m->nch, ndec = m->ndec;
  const int64_t P = dcs_num_patches(T, tc, overlap, patcher);
  if (P == 0) {  // clip shorter than one patch: nothing is predicted, every stem is silence
    for (int s = 0; s < m->nsrc * nch; ++s) DCS_CUDA(cudaMemsetAsync(d_S + s * src_stride, 0, (size_t)T * ldf * sizeof(float2), st));
    return DCS_OK;
  }
  DCS_REQUIRE(P * ndec * tc < (int64_t)1 << 31, "clip too long (%lld patches)", (long long)P);
  const int64_t Tp = std::max<int64_t>(T, (P - 1) * step + tc);
  const int HP = h2 + 2 * (kh2 - 1), ldg = (C1 + 3) / 4 * 4;
  DevBuf &bH1 = ctx->net[0], &bH2 = ctx->net[1], &bz = ctx->net[2], &bap = ctx->net[3], &bG = ctx->net[4];
  const uint64_t sig = ((uint64_t)(m->arch + 1) << 48) ^ ((uint64_t)m->F << 24) ^ (uint64_t)(tc * 64);
  // zero on (re)allocation or layout change; afterwards only the interior (rows and the C of the
  // Cp channels) is ever written, so the zero padding persists
  DCS_TRY(ensure_layout(ctx, 0, (size_t)Tp * C1p * 4, sig, st));
  DCS_TRY(ensure_layout(ctx, 1, (size_t)(Tp - kh2 + 1) * C2p * 4, sig, st));
  DCS_TRY(ensure_layout(ctx, 2, (size_t)P * nfc * 4, sig, st));
  DCS_TRY(ensure_layout(ctx, 3, (size_t)P * ndec * HP * C2p * 4, sig, st));
  // G: the transposed conv2 output.  For the tensor-core mask kernel it is
  // stored frame-major ([T][6 slots][ndec decoders][ldg], GemmDesc fm_*) so that a group of frames is one
  // TMA box; the FFMA mask kernel reads the patch-major [P][ndec][tc][ldg] order.  Unwritten slots must
  // stay zero / finite: the layout kind is part of the signature, so switching re-zeroes the buffer.
  const bool mask_tc = !ctx->debug_simt_gemm && (tc + step - 1) / step <= 6;
  const size_t g_rows = mask_tc ? (size_t)T * 6 * ndec : (size_t)P * ndec * tc;
  DCS_TRY(ensure_layout(ctx, 4, g_rows * ldg * 4, sig ^ (mask_tc ? 0x5a5a : 0), st));
  float *H1 = bH1.as<float>(), *H2 = bH2.as<float>(), *z = bz.as<float>(), *ap = bap.as<float>(), *G = bG.as<float>();

  // conv1 + both biases, once per frame (kernel height 1): H1[Tp][C1] = mag[T][nch x F] * W1f
  GemmDesc g1 = gemm_plain(d_mag, ldf, m->W1f, C1, m->b1, H1, C1p, (int)Tp, C1, nch * m->F, 0);
  if (nch > 1) { g1.k_seg = m->F; g1.k_ss = mag_plane; }   // one K segment per input channel plane
  g1.a_valid_rows = (int)T;  // util patcher: frames beyond T are zero input
  { ProfScope ps(ctx, "enc_conv1_gemm", st); DCS_TRY(run_gemm(ctx, g1, m->tW1f, st)); }
  // conv2 + both biases, once per frame offset: rows overlap in H1 (stride C1p, length kh2*C1p)
  GemmDesc g2 = gemm_plain(H1, C1p, m->W2c, C2, m->b2, H2, C2p, (int)(Tp - kh2 + 1), C2, kh2 * C1p, 0);
  { ProfScope ps(ctx, "enc_conv2_gemm", st); DCS_TRY(run_gemm(ctx, g2, m->tW2c, st)); }
  // bottleneck: patch k reads H2 rows k*step .. k*step+h2-1 (contiguous h2*C2p floats)
  GemmDesc g3 = gemm_plain(H2, (int64_t)step * C2p, m->Wfc, nfc, m->bfc, z, nfc, (int)P, nfc, h2 * C2p, 1);
  { ProfScope ps(ctx, "bottleneck_gemm", st); DCS_TRY(run_gemm(ctx, g3, m->tWfc, st)); }
  // the decoder dense layers side by side, scattered into the zero-padded buffer
  GemmDesc g4 = gemm_plain(z, nfc, m->Wdec, ndec * h2 * C2p, m->bdec, ap, (int64_t)ndec * HP * C2p, (int)P, ndec * h2 * C2p, nfc, 1);
  g4.n_seg = h2 * C2p; g4.n_ss = (int64_t)HP * C2p; g4.c_col0 = (int64_t)(kh2 - 1) * C2p;
  { ProfScope ps(ctx, "dec_dense_gemm", st); DCS_TRY(run_gemm(ctx, g4, m->tWdec, st)); }
  // InverseLayer(conv2): full correlation on the padded activations, rows (k, d, u)
  // Rows are ordered (u, k, d) -- u-major -- so that a 128-row tile holds one output position
  // u and can skip the taps that only see the zero padding (on average 8 of the 15).
  GemmDesc g5 = gemm_plain(ap, 0, m->Wt2, C1, nullptr, G, ldg, (int)(P * ndec * tc), C1, kh2 * C2p, 0);
  g5.m_inner = (int)(P * ndec); g5.a_so = C2p; g5.a_si = (int64_t)HP * C2p;
  g5.cm_inner = (int)(P * ndec); g5.c_so = ldg; g5.c_si = (int64_t)tc * ldg;
  g5.kc_rows = (int)(P * ndec); g5.kc_unit = C2p; g5.kc_pad = kh2 - 1; g5.kc_n = h2; g5.kc_taps = kh2;
  if (mask_tc) { g5.fm_step = step; g5.fm_tc = tc; g5.fm_T = (int)T; g5.fm_slots = 6; g5.fm_ndec = ndec; }
  { ProfScope ps(ctx, "dec_convT2_gemm", st); DCS_TRY(run_gemm(ctx, g5, m->tWt2, st)); }
  // InverseLayer(conv1) + bias + ReLU + mask + cross-fade + phase; the stereo net: once per channel
  // with that channel's conv1 weights, output biases and mixture STFT (trainCNN_ILD_DSD100.py:183-186)
  ProfScope ps(ctx, "dec_convT1_mask_xfade", st);
  for (int ch = 0; ch < nch; ++ch) {
    DsdMaskArgs a;
    a.G = G; a.ldg = ldg; a.W1t = m->W1t + (int64_t)ch * C1 * m->ldw; a.ldw = (int)m->ldw; a.bout = m->bout + 4 * ch;
    a.X = d_X + ch * x_plane; a.S = d_S + ch * src_stride;
    a.ldf = ldf; a.src_stride = nch * src_stride; a.T = (int)T; a.P = (int)P; a.tc = tc; a.overlap = overlap; a.F = m->F;
    a.ndec = ndec;
    if (mask_tc) {
      DCS_REQUIRE(dsd_mask_tc_supported(a), "dsd_forward: tensor-core mask kernel does not take this shape");
      DCS_TRY(launch_dsd_mask_tc(ctx, a, st));
    } else {
      DCS_TRY(launch_dsd_mask(ctx, a, st));   // FFMA kernel: > 6 patches per frame, bring-up cross-check
    }
  }
  return DCS_OK;
}

int dcs_separate_spec(dcs_ctx* ctx, dcs_model* m, const float* d_mag, const dcs_complex* d_X, int64_t T, int64_t ldf,
                      int overlap, int patcher, dcs_complex* d_S, int64_t src_stride, void* stream) {
  DCS_REQUIRE(ctx && m && d_mag && d_X && d_S, "dcs_separate_spec: NULL argument");
  DCS_REQUIRE(T > 0 && ldf >= m->F && src_stride >= T * ldf, "dcs_separate_spec: bad shape");
  DCS_REQUIRE(overlap >= 0 && overlap < m->tc, "overlap %d must be in [0, time_context=%d)", overlap, m->tc);
  DCS_REQUIRE(patcher == DCS_PATCHER_STANDALONE || patcher == DCS_PATCHER_UTIL, "unknown patcher %d", patcher);
  DCS_CUDA(cudaSetDevice(ctx->device));
  switch (m->arch) {
    case DCS_ARCH_DSD:
      return dsd_forward(ctx, m, d_mag, 0, (const float2*)d_X, 0, T, ldf, overlap, patcher, (float2*)d_S, src_stride,
                         (cudaStream_t)stream);
    case DCS_ARCH_DSD_ILD:
      DCS_REQUIRE(false, "the stereo network takes two input channels: use dcs_separate_audio_stereo");
    case DCS_ARCH_IKALA:
    case DCS_ARCH_IKALA_NOPOOL:
    case DCS_ARCH_BACH10:
      return sconv_forward(ctx, m, d_mag, 0, (const float2*)d_X, T, ldf, overlap, patcher, (float2*)d_S, src_stride,
                           (cudaStream_t)stream);
    case DCS_ARCH_BACH10_SCORE:
      DCS_REQUIRE(false, "the score-informed network takes 4 input channels: use dcs_separate_spec_channels / dcs_separate_audio_score");
  }
  DCS_REQUIRE(false, "architecture %d has no CUDA path yet", m->arch);
}

int dcs_separate_spec_channels(dcs_ctx* ctx, dcs_model* m, const float* d_in, int64_t in_plane, const dcs_complex* d_X,
                               int64_t T, int64_t ldf, int overlap, int patcher, dcs_complex* d_S, int64_t src_stride,
                               void* stream) {
  DCS_REQUIRE(ctx && m && d_in && d_X && d_S, "dcs_separate_spec_channels: NULL argument");
  DCS_REQUIRE(m->arch == DCS_ARCH_BACH10_SCORE, "dcs_separate_spec_channels: architecture %d has a single input channel", m->arch);
  DCS_REQUIRE(T > 0 && ldf >= m->F && src_stride >= T * ldf && in_plane >= T * ldf, "dcs_separate_spec_channels: bad shape");
  DCS_REQUIRE(overlap >= 0 && overlap < m->tc, "overlap %d must be in [0, time_context=%d)", overlap, m->tc);
  DCS_REQUIRE(patcher == DCS_PATCHER_STANDALONE || patcher == DCS_PATCHER_UTIL, "unknown patcher %d", patcher);
  DCS_CUDA(cudaSetDevice(ctx->device));
  return sconv_forward(ctx, m, d_in, in_plane, (const float2*)d_X, T, ldf, overlap, patcher, (float2*)d_S, src_stride,
                       (cudaStream_t)stream);
}

int dcs_separate_audio_score(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, const float* d_audio, int64_t L, const float* d_filters,
                             float scale_factor, int overlap, int patcher, float* d_stems, int64_t stem_stride, void* stream) {
  DCS_REQUIRE(ctx && m && p && d_audio && d_filters && d_stems, "dcs_separate_audio_score: NULL argument");
  DCS_REQUIRE(m->arch == DCS_ARCH_BACH10_SCORE, "dcs_separate_audio_score: needs the score-informed architecture");
  DCS_REQUIRE(L > 0 && stem_stride >= L && p->N / 2 + 1 == m->F, "dcs_separate_audio_score: bad length / frame size");
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  const int64_t T = dcs_num_frames(L, p->hop), ldf = dcs_padded_bins(p->N), plane = T * ldf;
  DCS_TRY(ctx->X.ensure((size_t)plane * sizeof(float2), st));
  DCS_TRY(ctx->mag.ensure((size_t)plane * sizeof(float), st));
  DCS_TRY(ctx->S.ensure((size_t)m->nsrc * plane * sizeof(float2), st));
  DCS_TRY(ctx->net[7].ensure((size_t)4 * plane * sizeof(float), st));
  float2* X = ctx->X.as<float2>();
  float* mag = ctx->mag.as<float>();
  float2* S = ctx->S.as<float2>();
  float* chans = ctx->net[7].as<float>();
  { ProfScope ps(ctx, "stft_fwd", st); DCS_TRY(launch_stft(p, d_audio, L, X, mag, nullptr, scale_factor, ldf, st)); }
  { ProfScope ps(ctx, "score_channels", st); DCS_TRY(launch_channel_mul(ctx, mag, d_filters, chans, plane, 4, st)); }
  DCS_TRY(sconv_forward(ctx, m, chans, plane, X, T, ldf, overlap, patcher, S, plane, st));
  DCS_TRY(copy_tap(ctx, S, (int64_t)m->nsrc * plane, st));
  ProfScope ps(ctx, "istft_ola", st);
  return launch_istft(p, S, nullptr, nullptr, 1.f, m->nsrc, T, ldf, plane, d_stems, L, stem_stride, st);
}

int dcs_gemm_f32(dcs_ctx* ctx, int engine, const float* d_A, int64_t lda, const float* h_B, int64_t ldb,
                 const float* h_bias, float* d_C, int64_t ldc, int M, int N, int K, int relu, void* stream) {
  DCS_REQUIRE(ctx && d_A && h_B && d_C && M > 0 && N > 0 && K > 0, "dcs_gemm_f32: bad argument");
  DCS_REQUIRE(lda >= 1 && ldb >= N && ldc >= N, "dcs_gemm_f32: leading dimension too small");  // lda < K: overlapping rows
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  float* d_bias = nullptr;
  float* d_B = nullptr;
  TcWeight w;
  int r = DCS_OK;
  if (h_bias) {
    std::vector<float> hb(h_bias, h_bias + N);
    r = upload(hb, &d_bias);
  }
  GemmDesc g = gemm_plain(d_A, lda, nullptr, N, d_bias, d_C, ldc, M, N, K, relu);
  if (r == DCS_OK) {
    if (engine == 1) {
      r = tc_weight_create(h_B, ldb, K, N, &w);
      if (r == DCS_OK) r = launch_gemm_tc(ctx, g, w, st);
    } else {
      std::vector<float> hB((size_t)K * N);
      for (int k = 0; k < K; ++k) memcpy(&hB[(size_t)k * N], h_B + (size_t)k * ldb, (size_t)N * sizeof(float));
      r = upload(hB, &d_B);
      g.B = d_B;
      if (r == DCS_OK) r = launch_gemm(ctx, g, st);
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  tc_weight_destroy(&w);
  if (d_B) cudaFree(d_B);
  if (d_bias) cudaFree(d_bias);
  if (r == DCS_OK && e != cudaSuccess) {
    set_error("dcs_gemm_f32: %s", cudaGetErrorString(e));
    return DCS_ECUDA;
  }
  return r;
}

int dcs_separate_audio_stereo(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, const float* d_audio, int64_t audio_stride, int64_t L,
                              float scale_factor, int overlap, int patcher, float* d_stems, int64_t stem_stride, void* stream) {
  DCS_REQUIRE(ctx && m && p && d_audio && d_stems, "dcs_separate_audio_stereo: NULL argument");
  DCS_REQUIRE(m->arch == DCS_ARCH_DSD_ILD, "dcs_separate_audio_stereo: needs the stereo / ILD architecture");
  DCS_REQUIRE(L > 0 && stem_stride >= L && audio_stride >= L && p->N / 2 + 1 == m->F, "dcs_separate_audio_stereo: bad length / frame size");
  DCS_REQUIRE(overlap >= 0 && overlap < m->tc, "overlap %d must be in [0, time_context=%d)", overlap, m->tc);
  DCS_REQUIRE(patcher == DCS_PATCHER_STANDALONE || patcher == DCS_PATCHER_UTIL, "unknown patcher %d", patcher);
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  const int nch = m->nch;
  const int64_t T = dcs_num_frames(L, p->hop), ldf = dcs_padded_bins(p->N), plane = T * ldf;
  DCS_TRY(ctx->X.ensure((size_t)nch * plane * sizeof(float2), st));
  DCS_TRY(ctx->mag.ensure((size_t)nch * plane * sizeof(float), st));
  DCS_TRY(ctx->S.ensure((size_t)m->nsrc * nch * plane * sizeof(float2), st));
  float2* X = ctx->X.as<float2>();
  float* mag = ctx->mag.as<float>();
  float2* S = ctx->S.as<float2>();
  {
    ProfScope ps(ctx, "stft_fwd", st);   // compute_transform: one STFT per channel (transform.py:105-119)
    for (int ch = 0; ch < nch; ++ch)
      DCS_TRY(launch_stft(p, d_audio + ch * audio_stride, L, X + ch * plane, mag + ch * plane, nullptr, scale_factor, ldf, st));
  }
  DCS_TRY(dsd_forward(ctx, m, mag, plane, X, plane, T, ldf, overlap, patcher, S, plane, st));
  DCS_TRY(copy_tap(ctx, S, (int64_t)m->nsrc * nch * plane, st));
  ProfScope ps(ctx, "istft_ola", st);
  return launch_istft(p, S, nullptr, nullptr, 1.f, m->nsrc * nch, T, ldf, plane, d_stems, L, stem_stride, st);
}

int dcs_xcorr_lags(dcs_ctx* ctx, const float* const* h_a, const float* const* h_b, int npairs, int64_t num_samples, int flen,
                   double* h_out, void* stream) {
  DCS_REQUIRE(ctx && h_a && h_b && h_out, "dcs_xcorr_lags: NULL argument");
  DCS_CUDA(cudaSetDevice(ctx->device));
  return launch_xcorr_lags(ctx, h_a, h_b, npairs, num_samples, flen, h_out, (cudaStream_t)stream);
}

int dcs_separate_audio(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, const float* d_audio, int64_t L, float scale_factor,
                       int overlap, int patcher, float* d_stems, int64_t stem_stride, void* stream) {
  DCS_REQUIRE(ctx && m && p && d_audio && d_stems, "dcs_separate_audio: NULL argument");
  DCS_REQUIRE(L > 0 && stem_stride >= L, "dcs_separate_audio: bad length");
  DCS_REQUIRE(p->N / 2 + 1 == m->F, "frame size %d does not give the model's %d bins", p->N, m->F);
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  const int64_t T = dcs_num_frames(L, p->hop), ldf = dcs_padded_bins(p->N);
  DCS_TRY(ctx->X.ensure((size_t)T * ldf * sizeof(float2), st));
  DCS_TRY(ctx->mag.ensure((size_t)T * ldf * sizeof(float), st));
  DCS_TRY(ctx->S.ensure((size_t)m->nsrc * T * ldf * sizeof(float2), st));
  float2* X = ctx->X.as<float2>();
  float* mag = ctx->mag.as<float>();
  float2* S = ctx->S.as<float2>();
  { ProfScope ps(ctx, "stft_fwd", st); DCS_TRY(launch_stft(p, d_audio, L, X, mag, nullptr, scale_factor, ldf, st)); }
  DCS_TRY(dcs_separate_spec(ctx, m, mag, (const dcs_complex*)X, T, ldf, overlap, patcher, (dcs_complex*)S, T * ldf, stream));
  DCS_TRY(copy_tap(ctx, S, (int64_t)m->nsrc * T * ldf, st));
  ProfScope ps(ctx, "istft_ola", st);
  return launch_istft(p, S, nullptr, nullptr, 1.f, m->nsrc, T, ldf, T * ldf, d_stems, L, stem_stride, st);
}

int dcs_separate_host(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, const float* h_audio, int64_t L, float scale_factor,
                      int overlap, int patcher, float* h_stems, int64_t stem_stride, void* stream) {
  DCS_REQUIRE(ctx && m && h_audio && h_stems && L > 0 && stem_stride >= L, "dcs_separate_host: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  DCS_TRY(ctx->audio.ensure((size_t)L * sizeof(float), st));
  DCS_TRY(ctx->stems.ensure((size_t)m->nsrc * L * sizeof(float), st));
  DCS_CUDA(cudaMemcpyAsync(ctx->audio.p, h_audio, (size_t)L * sizeof(float), cudaMemcpyHostToDevice, st));
  DCS_TRY(dcs_separate_audio(ctx, m, p, ctx->audio.as<float>(), L, scale_factor, overlap, patcher, ctx->stems.as<float>(), L, stream));
  DCS_CUDA(cudaMemcpy2DAsync(h_stems, (size_t)stem_stride * sizeof(float), ctx->stems.p, (size_t)L * sizeof(float),
                             (size_t)L * sizeof(float), m->nsrc, cudaMemcpyDeviceToHost, st));
  DCS_CUDA(cudaStreamSynchronize(st));
  return DCS_OK;
}

int dcs_separate_pcm16_host(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, const int16_t* h_pcm, int64_t L, int channels,
                            int downmix, float scale_factor, int overlap, int patcher, int16_t* h_out,
                            int64_t out_stride, void* stream) {
  DCS_REQUIRE(ctx && m && h_pcm && h_out && L > 0 && out_stride >= L, "dcs_separate_pcm16_host: bad argument");
  DCS_REQUIRE(channels >= 1 && channels <= 8 && downmix >= 0 && downmix <= 2, "bad channels/downmix");
  DCS_REQUIRE(downmix == 0 || channels >= 2 || channels == 1, "downmix needs two channels");
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  DCS_TRY(ctx->pcm_in.ensure((size_t)L * channels * sizeof(int16_t), st));
  DCS_TRY(ctx->pcm_out.ensure((size_t)m->nsrc * L * sizeof(int16_t), st));
  DCS_TRY(ctx->audio.ensure((size_t)L * sizeof(float), st));
  DCS_TRY(ctx->stems.ensure((size_t)m->nsrc * L * sizeof(float), st));
  DCS_CUDA(cudaMemcpyAsync(ctx->pcm_in.p, h_pcm, (size_t)L * channels * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  DCS_TRY(launch_pcm_decode(ctx, ctx->pcm_in.as<int16_t>(), L, channels, downmix, ctx->audio.as<float>(), st));
  DCS_TRY(dcs_separate_audio(ctx, m, p, ctx->audio.as<float>(), L, scale_factor, overlap, patcher, ctx->stems.as<float>(), L, stream));
  DCS_TRY(launch_pcm_encode(ctx, ctx->stems.as<float>(), L, m->nsrc, L, ctx->pcm_out.as<int16_t>(), L, st));
  DCS_CUDA(cudaMemcpy2DAsync(h_out, (size_t)out_stride * sizeof(int16_t), ctx->pcm_out.p, (size_t)L * sizeof(int16_t),
                             (size_t)L * sizeof(int16_t), m->nsrc, cudaMemcpyDeviceToHost, st));
  DCS_CUDA(cudaStreamSynchronize(st));
  return DCS_OK;
}

// Multi-clip scheduler: the clips of a batch run through ONE context as a three-stage pipeline -- H2D of clip i+1
// (copy stream) | kernels of clip i (the caller's stream) | D2H of clip i-1 (second copy stream) -- with double-buffered
// int16 staging on the device and events for the hand-overs.  The reference's only multi-clip driver starts a Python
// process per file (examples/dsd100/separate_multiple.ipynb cell 3); per clip this is the wav contract of train_auto
// (separate_dsd.py:275-287,307-309), exactly dcs_separate_pcm16_host.  Host buffers should be pinned.
// the pipelined loop of dcs_separate_batch_pcm16_host; on any failure the caller drains the copy streams before it
// returns, because the copies in flight read and write the user's host buffers
static int batch_pipeline(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, int nclips, const int16_t* const* h_pcm,
                          const int64_t* num_samples, int channels, int downmix, float scale_factor, int overlap,
                          int patcher, int16_t* const* h_out, const int64_t* out_strides, cudaStream_t st) {
  // the copy streams start after whatever the caller queued on `st` (and after the memsets of fresh buffers)
  DCS_CUDA(cudaEventRecord(ctx->ev_dec[0], st));
  DCS_CUDA(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_dec[0], 0));
  DCS_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_dec[0], 0));
  for (int i = 0; i < nclips; ++i) {
    const int b = i & 1;
    const int64_t L = num_samples[i];
    // H2D of clip i: its staging buffer is free once the decode of clip i-2 has read it
    if (i >= 2) DCS_CUDA(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_dec[b], 0));
    DCS_CUDA(cudaMemcpyAsync(ctx->pcm_in2[b].p, h_pcm[i], (size_t)L * channels * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->s_h2d));
    DCS_CUDA(cudaEventRecord(ctx->ev_in[b], ctx->s_h2d));
    // kernels of clip i
    DCS_CUDA(cudaStreamWaitEvent(st, ctx->ev_in[b], 0));
    DCS_TRY(launch_pcm_decode(ctx, ctx->pcm_in2[b].as<int16_t>(), L, channels, downmix, ctx->audio.as<float>(), st));
    DCS_CUDA(cudaEventRecord(ctx->ev_dec[b], st));
    DCS_TRY(dcs_separate_audio(ctx, m, p, ctx->audio.as<float>(), L, scale_factor, overlap, patcher, ctx->stems.as<float>(), L, (void*)st));
    if (i >= 2) DCS_CUDA(cudaStreamWaitEvent(st, ctx->ev_out[b], 0));     // D2H of clip i-2 has drained the output staging
    DCS_TRY(launch_pcm_encode(ctx, ctx->stems.as<float>(), L, m->nsrc, L, ctx->pcm_out2[b].as<int16_t>(), L, st));
    DCS_CUDA(cudaEventRecord(ctx->ev_enc[b], st));
    // D2H of clip i
    DCS_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_enc[b], 0));
    DCS_CUDA(cudaMemcpy2DAsync(h_out[i], (size_t)out_strides[i] * sizeof(int16_t), ctx->pcm_out2[b].p, (size_t)L * sizeof(int16_t),
                               (size_t)L * sizeof(int16_t), m->nsrc, cudaMemcpyDeviceToHost, ctx->s_d2h));
    DCS_CUDA(cudaEventRecord(ctx->ev_out[b], ctx->s_d2h));
  }
  return DCS_OK;
}

int dcs_separate_batch_pcm16_host(dcs_ctx* ctx, dcs_model* m, dcs_stft* p, int nclips, const int16_t* const* h_pcm,
                                  const int64_t* num_samples, int channels, int downmix, float scale_factor, int overlap,
                                  int patcher, int16_t* const* h_out, const int64_t* out_strides, void* stream) {
  DCS_REQUIRE(ctx && m && p && h_pcm && num_samples && h_out && out_strides && nclips >= 0, "dcs_separate_batch_pcm16_host: bad argument");
  DCS_REQUIRE(channels >= 1 && channels <= 8 && downmix >= 0 && downmix <= 2, "bad channels/downmix");
  if (nclips == 0) return DCS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DCS_CUDA(cudaSetDevice(ctx->device));
  int64_t Lmax = 0;
  for (int i = 0; i < nclips; ++i) {
    DCS_REQUIRE(h_pcm[i] && h_out[i] && num_samples[i] > 0 && out_strides[i] >= num_samples[i], "clip %d: bad buffer / length", i);
    Lmax = std::max(Lmax, num_samples[i]);
  }
  // each resource on its own: a call that failed half-way through this block must not leave later calls with null handles
  if (!ctx->s_h2d) DCS_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
  if (!ctx->s_d2h) DCS_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    if (!ctx->ev_in[i]) DCS_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
    if (!ctx->ev_dec[i]) DCS_CUDA(cudaEventCreateWithFlags(&ctx->ev_dec[i], cudaEventDisableTiming));
    if (!ctx->ev_enc[i]) DCS_CUDA(cudaEventCreateWithFlags(&ctx->ev_enc[i], cudaEventDisableTiming));
    if (!ctx->ev_out[i]) DCS_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
  }
  // every buffer at the size of the longest clip before the pipeline starts: a grow-only buffer that had to be
  // re-allocated mid-batch would synchronise the stream
  for (int b = 0; b < 2; ++b) {
    DCS_TRY(ctx->pcm_in2[b].ensure((size_t)Lmax * channels * sizeof(int16_t), st));
    DCS_TRY(ctx->pcm_out2[b].ensure((size_t)m->nsrc * Lmax * sizeof(int16_t), st));
  }
  DCS_TRY(ctx->audio.ensure((size_t)Lmax * sizeof(float), st));
  DCS_TRY(ctx->stems.ensure((size_t)m->nsrc * Lmax * sizeof(float), st));
  {
    const int64_t T = dcs_num_frames(Lmax, p->hop), ldf = dcs_padded_bins(p->N);
    DCS_TRY(ctx->X.ensure((size_t)T * ldf * sizeof(float2), st));
    DCS_TRY(ctx->mag.ensure((size_t)T * ldf * sizeof(float), st));
    DCS_TRY(ctx->S.ensure((size_t)m->nsrc * T * ldf * sizeof(float2), st));
  }
  const int rc = batch_pipeline(ctx, m, p, nclips, h_pcm, num_samples, channels, downmix, scale_factor, overlap, patcher,
                                h_out, out_strides, st);
  // drain everything, success or not, before the host buffers go back to the caller
  const cudaError_t e0 = cudaStreamSynchronize(ctx->s_h2d), e1 = cudaStreamSynchronize(ctx->s_d2h), e2 = cudaStreamSynchronize(st);
  if (rc != DCS_OK) return rc;
  DCS_CUDA(e0);
  DCS_CUDA(e1);
  DCS_CUDA(e2);
  return DCS_OK;
}

}  // extern "C"

// gemm_tc.cu -- fp32-accurate GEMM on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with the
// 3xTF32 operand split, accumulators in TMEM, C = act(A*B + bias).
//
// Same operand model as gemm.cu (GemmDesc): A is a strided / overlapping / K-segmented *view*
// of an fp32 activation buffer in HBM (convolution rows are never materialised); B is a weight
// matrix that was transposed to K-major, zero padded and split into (hi, lo) once at model load
// (TcWeight).  Per CTA: one 128 x BN output tile.
//   warps 0-3  producers: global fp32 -> registers -> hi/lo split -> canonical K-major shared
//              tiles (tc.cuh), fence.proxy.async, mbarrier arrive; afterwards the epilogue:
//              tcgen05.ld (thread = output row = TMEM lane) -> +bias, ReLU -> global
//   warp 4     TMEM allocation; lane 0 issues 3 tcgen05.mma per k-step and tcgen05.commit's
//              the stage back to the producers
// STAGES-deep mbarrier ring; with BN=64, 2 stages = 96 KB so two CTAs share an SM and one's
// epilogue overlaps the other's main loop.
#include "common.cuh"
#include "tc.cuh"

namespace dcs {

using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_PGROUPS = 2;                 // producer groups of 4 warps, alternating stages
constexpr int TC_MMA_WARP = 4 * TC_PGROUPS;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;  // 288

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * ROW_BYTES;  // 16 KB
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // + alignment slack
};

// Accumulator plan.  The tensor core adds into the fp32 TMEM accumulator with truncation, so a
// chain of n sequential accumulations loses ~n * 2^-24 relative (measured: 5.7e-6 at K = 750 with
// a single accumulator, 3 MMAs per k-step).  acc_mode 2 (default) spreads the main hi*hi term
// round-robin over three accumulators and sends both correction terms (2^-11 smaller, their
// truncation is harmless) to a fourth; the epilogue adds the four in fp32 registers.
//   acc_mode 0: one accumulator for everything; 1: main + corrections; 2: 3 x main + corrections
template <int BN, int STAGES, int AVEC, bool SPLITK>
__global__ void __launch_bounds__(TC_THREADS)
gemm_tc_kernel(const GemmDesc d, const float* __restrict__ Bhi, const float* __restrict__ Blo, int Kp, int acc_mode, int dbg,
               int k_splits, float* __restrict__ partial, int ldp) {
  using SM = TcSmem<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  int kb_lo = 0, kb_hi = (d.K + KSTAGE - 1) / KSTAGE;
  if (d.kc_rows > 0) {
    const int u_min = m0 / d.kc_rows, u_max = min(d.M - 1, m0 + TC_BM - 1) / d.kc_rows;
    const int q_lo = max(0, d.kc_pad - u_max), q_hi = min(d.kc_taps - 1, d.kc_pad + d.kc_n - 1 - u_min);
    kb_lo = (d.kc_unit * q_lo) / KSTAGE;
    kb_hi = min(kb_hi, (d.kc_unit * (q_hi + 1) + KSTAGE - 1) / KSTAGE);
    if (kb_hi <= kb_lo) kb_hi = kb_lo + 1;   // keep one (all-zero) block so the accumulators are defined
  }
  if (SPLITK) {   // split-K: this CTA (blockIdx.z) owns a contiguous slice of the k-blocks
    const int per = (kb_hi - kb_lo + k_splits - 1) / k_splits;
    kb_lo += blockIdx.z * per;
    kb_hi = min(kb_hi, kb_lo + per);
    if (kb_hi <= kb_lo) kb_hi = kb_lo + 1;   // cannot happen (the launcher sizes the slices); loads past K are guarded
  }
  const int num_kb = kb_hi - kb_lo;   // k-block i of this tile is global block kb_lo + i

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  constexpr uint32_t TMEM_COLS = 4 * BN;
  const int n_main = acc_mode == 2 ? 3 : 1;
  const int corr_acc = acc_mode == 0 ? 0 : n_main;       // accumulator index of the corrections
  const int n_main_used = min(n_main, num_kb * (KSTAGE / 8));
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TC_MMA_WARP) {
    // ------------------------------------------------------------------ producers
    // group pg stages k-blocks pg, pg+2, ...: two stage loads are in flight per CTA
    const int pg = warp >> 2;
    const int tid = threadIdx.x & 127, warp = tid >> 5;   // index inside the producer group
    // Each warp-wide load covers whole 128-byte rows (AVEC=4: 4 rows x 8 chunks of 16 B per
    // instruction, AVEC=2: 2 rows x 16 pieces of 8 B, AVEC=1: 1 row x 32 floats) so global reads
    // are coalesced, and the swizzled shared stores of one row hit 8 distinct 16-byte slots.
    constexpr int RPI = AVEC == 4 ? 4 : (AVEC == 2 ? 2 : 1);   // rows per warp instruction
    constexpr int NI = 32 / RPI;                                // instructions per thread per stage (A)
    const int sub = lane / (32 / RPI);                          // row within the instruction
    const int piece = lane % (32 / RPI);                        // piece within the row
    // this thread's A rows: warp w owns rows [32w, 32w+32)
    const float* arow[NI];
    bool row_ok[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int m = m0 + warp * 32 + i * RPI + sub;
      row_ok[i] = (m < d.M) && (m < d.a_valid_rows);
      const int mc = m < d.M ? m : 0;
      arow[i] = d.A + (int64_t)(mc / d.m_inner) * d.a_so + (int64_t)((mc % d.m_inner) / d.m_inner2) * d.a_si + (int64_t)(mc % d.m_inner2) * d.a_s2;
    }
    const bool seg_vec = d.k_seg < d.K;   // only reached with AVEC > 1 when k_seg % KSTAGE == 0 (launcher)
    constexpr int NBI = (2 * BN) / 16;   // B: 2*BN rows (hi+lo planes), 16 rows per pass of 128 threads
    float ra[NI][AVEC];
    float4 rb[NBI];
    auto gload = [&](int kb) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int k = (kb_lo + kb) * KSTAGE + piece * AVEC;
        // K split in segments (one per convolution tap): with k_seg a multiple of the 32-wide
        // stage a vector never straddles two segments
        const int64_t koff = seg_vec ? (int64_t)(k / d.k_seg) * d.k_ss + (k % d.k_seg) : (int64_t)k;
#pragma unroll
        for (int e = 0; e < AVEC; ++e) ra[i][e] = 0.f;
        if (row_ok[i] && k < d.K) {
          if (AVEC == 4 && k + 4 <= d.K) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(arow[i] + koff));
            ra[i][0] = v.x; ra[i][1] = v.y; ra[i][AVEC > 2 ? 2 : 0] = v.z; ra[i][AVEC > 3 ? 3 : 0] = v.w;
          } else if (AVEC == 2 && k + 2 <= d.K) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(arow[i] + koff));
            ra[i][0] = v.x; ra[i][AVEC > 1 ? 1 : 0] = v.y;
          } else {
#pragma unroll
            for (int e = 0; e < AVEC; ++e) {
              const int kk = k + e;
              if (kk < d.K) ra[i][e] = __ldg(arow[i] + (int64_t)(kk / d.k_seg) * d.k_ss + (kk % d.k_seg));
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NBI; ++i) {
        const int rr = i * 16 + (tid >> 3);          // row in the stacked (hi, lo) B planes
        const int which = rr / BN, rn = rr - which * BN;
        const float* src = (which ? Blo : Bhi) + (int64_t)(n0 + rn) * Kp + (kb_lo + kb) * KSTAGE + 4 * (tid & 7);
        rb[i] = (kb_lo + kb) * KSTAGE < Kp ? ld_stream(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // bring-up probes (DCS_DEBUG_TC_SKIP): bit0 = no global loads, bit1 = no MMAs, bit2 = no shared stores
    if (pg < num_kb && !(dbg & 1)) gload(pg);
    for (int kb = pg; kb < num_kb; kb += TC_PGROUPS) {
      const int s = kb % STAGES;
      const uint32_t par = (kb / STAGES) & 1;
      mbar_wait_relaxed(&empty[s], par ^ 1);
      uint8_t* st = smem + s * SM::STAGE_BYTES;
      if (!(dbg & 4)) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int r = warp * 32 + i * RPI + sub;
        const int e0 = piece * AVEC;                 // first fp32 of this piece within the row
        const uint32_t off = tile_off(r, e0 >> 2) + (e0 & 3) * 4;
        float hi[AVEC], lo[AVEC];
#pragma unroll
        for (int e = 0; e < AVEC; ++e) split_tf32(ra[i][e], hi[e], lo[e]);
        if (AVEC == 4) {
          *reinterpret_cast<float4*>(st + off) = make_float4(hi[0], hi[1], hi[AVEC > 2 ? 2 : 0], hi[AVEC > 3 ? 3 : 0]);
          *reinterpret_cast<float4*>(st + SM::A_BYTES + off) = make_float4(lo[0], lo[1], lo[AVEC > 2 ? 2 : 0], lo[AVEC > 3 ? 3 : 0]);
        } else if (AVEC == 2) {
          *reinterpret_cast<float2*>(st + off) = make_float2(hi[0], hi[AVEC > 1 ? 1 : 0]);
          *reinterpret_cast<float2*>(st + SM::A_BYTES + off) = make_float2(lo[0], lo[AVEC > 1 ? 1 : 0]);
        } else {
          *reinterpret_cast<float*>(st + off) = hi[0];
          *reinterpret_cast<float*>(st + SM::A_BYTES + off) = lo[0];
        }
      }
#pragma unroll
      for (int i = 0; i < NBI; ++i) {
        const int rr = i * 16 + (tid >> 3);
        const int which = rr / BN, rn = rr - which * BN;
        *reinterpret_cast<float4*>(st + 2 * SM::A_BYTES + which * SM::B_BYTES + tile_off(rn, tid & 7)) = rb[i];
      }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
      if (kb + TC_PGROUPS < num_kb && !(dbg & 1)) gload(kb + TC_PGROUPS);
    }
    if (pg == 0) {               // epilogue: warps 0-3, thread = output row = TMEM lane
    const int r = tid;
    const int m = m0 + r;
    const int mc = m < d.M ? m : 0;
    // ------------------------------------------------------------------ epilogue
    mbar_wait_relaxed(tmem_full, 0);
    fence_after_sync();
    const int64_t roff = gemm_c_row_offset(d, mc);
    const bool m_ok = m < d.M && roff >= 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int j = 0; j < BN / 16; ++j) {
      if (n0 + 16 * j >= d.N) break;  // warp-uniform
      float v[16];
      tmem_ld16(taddr + 16 * j, v);
      for (int a = 1; a < n_main_used; ++a) {
        float u[16];
        tmem_ld16(taddr + a * BN + 16 * j, u);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += u[i];
      }
      if (corr_acc) {
        float u[16];
        tmem_ld16(taddr + corr_acc * BN + 16 * j, u);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += u[i];
      }
      if (SPLITK && m_ok) {   // split-K: raw partial sums, reduced (+bias, activation) by splitk_reduce_kernel
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + 16 * j + i;
          if (n < d.N) partial[((int64_t)blockIdx.z * d.M + m) * ldp + n] = v[i];
        }
      } else if (!SPLITK && m_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + 16 * j + i;
          if (n < d.N) {
            float x = v[i];
            if (d.bias) x += __ldg(d.bias + n);
            if (d.relu) x = fmaxf(x, 0.f);
            d.C[roff + (int64_t)(n / d.n_seg) * d.n_ss + (n % d.n_seg)] = x;
          }
        }
      }
    }
    }
    fence_before_sync();
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t par = (kb / STAGES) & 1;
      mbar_wait(&full[s], par);
      fence_after_sync();
      const uint32_t a_hi = smem_u32(smem + s * SM::STAGE_BYTES);
      const uint32_t a_lo = a_hi + SM::A_BYTES;
      const uint32_t b_hi = a_hi + 2 * SM::A_BYTES;
      const uint32_t b_lo = b_hi + SM::B_BYTES;
      if (!(dbg & 2))
#pragma unroll
      for (int j = 0; j < KSTAGE / 8; ++j) {
        const uint64_t dah = make_desc(a_hi + KSTEP_BYTES * j), dal = make_desc(a_lo + KSTEP_BYTES * j);
        const uint64_t dbh = make_desc(b_hi + KSTEP_BYTES * j), dbl = make_desc(b_lo + KSTEP_BYTES * j);
        const int ks = kb * (KSTAGE / 8) + j;
        const int ma = ks % n_main;
        umma_tf32(tmem_base + corr_acc * BN, dal, dbh, idesc, ks != 0);
        umma_tf32(tmem_base + corr_acc * BN, dah, dbl, idesc, 1);
        umma_tf32(tmem_base + ma * BN, dah, dbh, idesc, corr_acc == 0 ? 1 : (ks >= n_main));
      }
      umma_commit(&empty[s]);
    }
    umma_commit(tmem_full);
  }
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void splitk_reduce_kernel(const GemmDesc d, const float* __restrict__ partial, int ldp, int k_splits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)d.M * d.N) return;
  const int m = (int)(i / d.N), n = (int)(i - (int64_t)m * d.N);
  float x = 0.f;
  for (int z = 0; z < k_splits; ++z) x += partial[((int64_t)z * d.M + m) * ldp + n];   // fixed order: deterministic
  if (d.bias) x += __ldg(d.bias + n);
  if (d.relu) x = fmaxf(x, 0.f);
  const int64_t roff = gemm_c_row_offset(d, m);
  if (roff >= 0) d.C[roff + (int64_t)(n / d.n_seg) * d.n_ss + (n % d.n_seg)] = x;
}

int launch_splitk_reduce(dcs_ctx* ctx, const GemmDesc& d, const float* partial, int ldp, int k_splits, cudaStream_t st) {
  const int64_t tot = (int64_t)d.M * d.N;
  splitk_reduce_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(d, partial, ldp, k_splits);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

// ---- host side ------------------------------------------------------------------------------
int tc_weight_create(const float* B, int64_t ldb, int K, int N, TcWeight* out) {
  // B[k][n] row-major (ld = ldb) -> K-major Bt[n][k], zero padded to Np x Kp, split hi/lo
  const int Kp = (K + KSTAGE - 1) / KSTAGE * KSTAGE, Np = (N + 127) / 128 * 128;  // the widest tile reads 128 rows
  std::vector<float> hi((size_t)Np * Kp, 0.f), lo((size_t)Np * Kp, 0.f);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      const float x = B[(size_t)k * ldb + n];
      uint32_t u;
      memcpy(&u, &x, 4);
      u &= 0xFFFFE000u;
      float h;
      memcpy(&h, &u, 4);
      hi[(size_t)n * Kp + k] = h;
      lo[(size_t)n * Kp + k] = x - h;
    }
  out->K = K; out->N = N; out->Kp = Kp; out->Np = Np;
  DCS_CUDA(cudaMalloc((void**)&out->hi, 2 * hi.size() * sizeof(float)));   // planes stacked: [hi; lo]
  out->lo = out->hi + hi.size();
  DCS_CUDA(cudaMemcpy(out->hi, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice));
  DCS_CUDA(cudaMemcpy(out->lo, lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice));
  tc_weight_encode_maps(out);
  return DCS_OK;
}

void tc_weight_destroy(TcWeight* w) {
  if (w->hi) cudaFree(w->hi);
  w->hi = w->lo = nullptr;
  w->tmap_ok[0] = w->tmap_ok[1] = w->tmap_ok[2] = false;
}

template <int BN, int STAGES, int AVEC>
static int launch_tc(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st) {
  using SM = TcSmem<BN, STAGES>;
  DCS_TRY(ensure_smem_attr(gemm_tc_kernel<BN, STAGES, AVEC, false>, SM::TOTAL));
  DCS_TRY(ensure_smem_attr(gemm_tc_kernel<BN, STAGES, AVEC, true>, SM::TOTAL));
  const int m_tiles = (int)ceil_div64(d.M, TC_BM), n_tiles = (int)ceil_div64(d.N, BN);
  const int num_kb = (d.K + KSTAGE - 1) / KSTAGE;
  // skinny GEMMs (few output tiles, long K -- the bottleneck dense layers): split K over otherwise idle SMs
  int splits = 1;
  if ((int64_t)m_tiles * n_tiles * 2 <= ctx->num_sms && num_kb >= 16 && d.kc_rows == 0) {
    splits = (int)std::min<int64_t>(ctx->num_sms / ((int64_t)m_tiles * n_tiles), num_kb / 8);
    if (splits < 2) splits = 1;
    if (splits > 1) {   // no empty slice: every blockIdx.z must own at least one real k-block
      const int per = (num_kb + splits - 1) / splits;
      splits = (num_kb + per - 1) / per;
    }
  }
  float* partial = nullptr;
  const int ldp = (d.N + 3) / 4 * 4;
  if (splits > 1) {
    DCS_TRY(ctx->net[8].ensure((size_t)splits * d.M * ldp * sizeof(float), st));
    partial = ctx->net[8].as<float>();
  }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
  if (splits > 1)
    gemm_tc_kernel<BN, STAGES, AVEC, true><<<grid, TC_THREADS, SM::TOTAL, st>>>(d, w.hi, w.lo, w.Kp, ctx->tc_acc_mode,
                                                                                 ctx->tc_debug, splits, partial, ldp);
  else
    gemm_tc_kernel<BN, STAGES, AVEC, false><<<grid, TC_THREADS, SM::TOTAL, st>>>(d, w.hi, w.lo, w.Kp, ctx->tc_acc_mode,
                                                                                  ctx->tc_debug, 1, nullptr, 0);
  if (splits > 1) {
    DCS_CHECK_LAUNCH();
    ctx->launches++;
    const int64_t tot = (int64_t)d.M * d.N;
    splitk_reduce_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(d, partial, ldp, splits);
  }
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

// `d.B` is ignored: the weight comes pre-transposed in `w`.
int launch_gemm_tc(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st) {
  if (d.M <= 0 || d.N <= 0) return DCS_OK;
  DCS_REQUIRE(d.K == w.K && d.N == w.N, "tc gemm: weight is %dx%d, GEMM wants K=%d N=%d", w.K, w.N, d.K, d.N);
  DCS_REQUIRE(ceil_div64(d.N, 64) <= 65535, "tc gemm: N=%d too large", d.N);
  // views the copy engine can describe go to the TMA-fed kernel (gemm_tma.cu)
  if (ctx->tma_mode && gemm_tma_eligible(d, ctx->tma_mask)) {
    const int r = launch_gemm_tma(ctx, d, w, st);
    if (r != DCS_TMA_FALLBACK) return r;
  }
  // vector width the A view allows: every row start and every segment must keep the alignment
  int avec = 1;
  const bool one_seg = d.k_seg >= d.K;
  auto aligned = [&](int v) {
    return ((uintptr_t)d.A % (4 * v) == 0) && d.a_so % v == 0 && d.a_si % v == 0 && d.a_s2 % v == 0 &&
           (one_seg || (d.k_seg % KSTAGE == 0 && d.k_ss % v == 0));
  };
  if (aligned(4)) avec = 4; else if (aligned(2)) avec = 2;
  if (d.N > 64) {
    if (avec == 4) return launch_tc<128, 3, 4>(ctx, d, w, st);
    if (avec == 2) return launch_tc<128, 3, 2>(ctx, d, w, st);
    return launch_tc<128, 3, 1>(ctx, d, w, st);
  }
  if (d.N <= 32) {   // the 30-channel convolutions of the iKala / Bach10 nets: half the weight traffic and MMA time
    if (avec == 4) return launch_tc<32, 4, 4>(ctx, d, w, st);
    if (avec == 2) return launch_tc<32, 4, 2>(ctx, d, w, st);
    return launch_tc<32, 4, 1>(ctx, d, w, st);
  }
  if (avec == 4) return launch_tc<64, 4, 4>(ctx, d, w, st);
  if (avec == 2) return launch_tc<64, 4, 2>(ctx, d, w, st);
  return launch_tc<64, 4, 1>(ctx, d, w, st);
}

}  // namespace dcs

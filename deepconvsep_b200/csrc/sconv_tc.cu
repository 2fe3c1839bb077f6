// sconv_tc.cu -- K3s on the tensor cores: InverseLayer(pool) + InverseLayer(conv1) (transposed strided
// convolution over frequency) + ConcatLayer + bias + ReLU + soft ratio mask + patch cross-fade + phase for
// the strided-conv1 networks -- iKala (examples/ikala/separate_ikala.py:183-217), Bach10
// (examples/bach10/separate_bach10.py:207-266), score-informed Bach10
// (examples/bach10_scoreinformed/trainCNNrwc.py:189,248-263).  Same math as sconv.cu's FFMA kernel
// (which stays as the bring-up cross-check, DCS_DEBUG_SIMT_GEMM=1).
//
// The transposed strided convolution  Y[STRIDE*m + r] = sum_{dd < ND, f < 30} Gu[m - dd][f] * w[dd][f][r]
// is computed in two steps:
//   1. GEMM on tcgen05 (fp32-accurate 3xTF32):  Z[j][(dd, r)] = sum_f Gu[j][f] * w[dd][f][r]
//        M = 128 staged positions j (TMEM lanes), K = 32 channels (30 + 2 zero), N = ND*STRIDE (30/32 -> 32)
//        columns per filter bank; A = the activation tile [128 j][32 f] -- for the un-pooled nets ONE box of
//        the copy engine (rows past either end of the axis are its zero fill), raw fp32 = the HIGH operand, four
//        warps derive the LOW plane; for the max-pool net the same four warps gather through the tie bits of
//        the forward pass (Theano MaxPoolGrad routing) and write both planes; B = the filter bank(s), split once.
//   2. the epilogue thread of output m gathers the ND taps  Y[m][r] = sum_dd Z[m - dd][(dd, r)]  from its
//        neighbours' rows through a padded shared-memory tile (conflict-free), then bias + ReLU + ratio mask over
//        the sources + the sequential cross-fade recurrence, all in registers; the masked spectra leave as
//        STRIDE consecutive bins per thread.
// A patch slot's decoders (Bach10: 4, iKala: 2) -- or, score-informed, the 4 filter banks of decoder 1 -- sit side
// by side in one 128-column TMEM accumulator, double buffered across slots; CTAs are persistent over a range of
// frames of one tile of 128 - (ND-1) output positions.
//   warps 0-3 epilogue | warp 4 MMA issue + TMEM alloc | warp 5 copy engine | warps 6-9 low plane / un-pool gather
#include "common.cuh"
#include "tc.cuh"

namespace dcs {

using namespace tc;

constexpr int ST_ROWS = 128;
constexpr int ST_A_TILE = ST_ROWS * ROW_BYTES;     // 16 KB: [128][32] fp32, one k-block
constexpr int ST_STAGES = 2;
constexpr int ST_EPI = 128, ST_PROD = 128;
constexpr int ST_THREADS = 10 * 32;
constexpr int ST_ZP = 33;                          // padded row of the Z exchange tile
constexpr uint32_t ST_TMEM_COLS = 256;

template <int STRIDE, int ND, int NSRC, int NDEC, int NW>
struct SconvTile {
  static constexpr int OUT = ST_ROWS - (ND - 1);     // output positions per tile
  static constexpr int NB = NW * 32;                 // GEMM columns per activation tile
  static constexpr int ITEMS = NDEC;                 // activation tiles per patch slot
  static constexpr int B_PLANE = NB * ROW_BYTES;
  static constexpr int OFF_B = ST_STAGES * 2 * ST_A_TILE;
  static constexpr int OFF_Z = OFF_B + 2 * B_PLANE;
  static constexpr int OFF_BAR = OFF_Z + 2 * ST_ROWS * ST_ZP * 4;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static_assert(ITEMS * NB <= 128, "one slot's accumulators must fit 128 TMEM columns");
  static_assert(NW == 1 ? NSRC == NDEC : (NDEC == 1 && NSRC == NW), "sources = decoders, or = filter banks of one decoder");
  static_assert(ND * STRIDE <= 32, "taps x stride must fit one 32-column bank");
};

__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ void slot_range(int t, int tc, int step, int P, int& k_lo, int& k_hi) {
  k_hi = t / step;
  if (k_hi > P - 1) k_hi = P - 1;
  k_lo = t - tc + 1;
  k_lo = k_lo > 0 ? (k_lo + step - 1) / step : 0;
}

template <int STRIDE, int ND, int NSRC, int NDEC, int RULE, int POOL, int NW>
__global__ void __launch_bounds__(ST_THREADS, 1)
sconv_mask_tc_kernel(const SconvMaskArgs a, const __grid_constant__ CUtensorMap tmG, int frames_per_cta) {
  using TL = SconvTile<STRIDE, ND, NSRC, NDEC, NW>;
  constexpr int OUT = TL::OUT, NB = TL::NB, ITEMS = TL::ITEMS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;                                   // stage s: hi at s*2*TILE, lo at s*2*TILE + TILE
  uint8_t* sB = smem + TL::OFF_B;                       // hi plane, lo plane
  float* Zs = reinterpret_cast<float*>(smem + TL::OFF_Z);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TL::OFF_BAR);   // activation tile landed / gathered
  uint64_t* split = full + ST_STAGES;                                  // low plane written (un-pooled nets)
  uint64_t* empty = split + ST_STAGES;                                 // MMAs of the stage retired
  uint64_t* tmem_full = empty + ST_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * OUT;
  const int j_start = m0 - (ND - 1);
  const int t_begin = blockIdx.y * frames_per_cta;
  const int t_end = min(a.T, t_begin + frames_per_cta);
  const int step = a.tc - a.overlap;

  if (tid == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(&full[s], POOL ? ST_PROD : 1);
      mbar_init(&split[s], ST_PROD);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], ST_EPI);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, ST_TMEM_COLS);
  if (warp == 5 && lane == 0 && !POOL) prefetch_tensormap(&tmG);
  // filter banks: B[c = o*32 + dd*STRIDE + r][f] = w[o][dd][f][r] (a.W: float4 [NW][ND][32], .xyzw = r), hi / lo planes
  for (int i = tid; i < NB * 8; i += ST_THREADS) {
    const int c = i >> 3, c4 = i & 7;
    const int o = c >> 5, cc = c & 31, dd = cc / STRIDE, r = cc - dd * STRIDE;
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    if (dd < ND) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int f = 4 * c4 + q;
        e[q] = __ldg(a.W + ((int64_t)((o * ND + dd) * 32 + f)) * 4 + r);
      }
    }
    float4 hi, lo;
    split4(make_float4(e[0], e[1], e[2], e[3]), hi, lo);
    const uint32_t off = tile_off(c, c4);
    *reinterpret_cast<float4*>(sB + off) = hi;
    *reinterpret_cast<float4*>(sB + TL::B_PLANE + off) = lo;
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ------------------------------------------------------------------ copy engine (un-pooled nets)
    if (!POOL && elect_one()) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int k_lo, k_hi;
        slot_range(t, a.tc, step, a.P, k_lo, k_hi);
        for (int k = k_lo; k <= k_hi; ++k) {
          const int p = t - k * step;
          for (int d = 0; d < ITEMS; ++d, ++it) {
            const int s = it % ST_STAGES;
            mbar_wait(&empty[s], ((it / ST_STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], ST_A_TILE);
            tma_load_3d(sA + s * 2 * ST_A_TILE, &tmG, &full[s], 0, j_start, (k * NDEC + d) * a.tc + p);
          }
        }
      }
    }
  } else if (warp > 5) {
    // ------------------------------------------------------------------ low plane / un-pool gather
    const int pt = tid - 6 * 32;  // 0..127
    int it = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int k_lo, k_hi;
      slot_range(t, a.tc, step, a.P, k_lo, k_hi);
      for (int k = k_lo; k <= k_hi; ++k) {
        const int p = t - k * step;
        for (int d = 0; d < ITEMS; ++d, ++it) {
          const int s = it % ST_STAGES;
          uint8_t* hi_t = sA + s * 2 * ST_A_TILE;
          uint8_t* lo_t = hi_t + ST_A_TILE;
          if (POOL) {
            // InverseLayer(pool): position j of the un-pooled axis receives G[j / POOL] where the forward pass had
            // its window maximum (every tied position does -- Theano MaxPoolGrad), else 0
            mbar_wait_relaxed(&empty[s], ((it / ST_STAGES) & 1) ^ 1);
            const int c4 = pt & 7;
            const float* grow = a.G + ((int64_t)(k * NDEC + d) * a.tc + p) * a.WP * 32 + 4 * c4;
            const uint8_t* trow = a.tie + (int64_t)t * a.WP * 32 + 4 * c4;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int jl = (pt >> 3) + 16 * u, j = j_start + jl;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (j >= 0) {
                constexpr int PW = POOL ? POOL : 1;
                const int jp = j / PW, rr = j - jp * PW;
                if (jp < a.WP) {
                  const uchar4 bits = *reinterpret_cast<const uchar4*>(trow + (int64_t)jp * 32);
                  const float4 g = __ldg(reinterpret_cast<const float4*>(grow + (int64_t)jp * 32));
                  v.x = ((bits.x >> rr) & 1) ? g.x : 0.f;
                  v.y = ((bits.y >> rr) & 1) ? g.y : 0.f;
                  v.z = ((bits.z >> rr) & 1) ? g.z : 0.f;
                  v.w = ((bits.w >> rr) & 1) ? g.w : 0.f;
                }
              }
              float4 h, l;
              split4(v, h, l);
              const uint32_t off = tile_off(jl, c4);
              *reinterpret_cast<float4*>(hi_t + off) = h;
              *reinterpret_cast<float4*>(lo_t + off) = l;
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
          } else {
            mbar_wait(&full[s], (it / ST_STAGES) & 1);
            const float4* raw = reinterpret_cast<const float4*>(hi_t);
            float4* lo = reinterpret_cast<float4*>(lo_t);
#pragma unroll
            for (int u = 0; u < ST_A_TILE / 16 / ST_PROD; ++u) {   // 8 chunks per thread, same swizzled offsets
              float4 h, l;
              split4(raw[u * ST_PROD + pt], h, l);
              lo[u * ST_PROD + pt] = l;
            }
            fence_proxy_async();
            mbar_arrive(&split[s]);
          }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(ST_ROWS, NB);
      const uint32_t b_hi = smem_u32(sB), b_lo = b_hi + TL::B_PLANE;
      int it = 0, sl = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int k_lo, k_hi;
        slot_range(t, a.tc, step, a.P, k_lo, k_hi);
        for (int k = k_lo; k <= k_hi; ++k, ++sl) {
          const int buf = sl & 1;
          for (int d = 0; d < ITEMS; ++d, ++it) {
            const int s = it % ST_STAGES;
            const uint32_t par = (it / ST_STAGES) & 1;
            mbar_wait(&full[s], par);
            if (d == 0) mbar_wait(&tmem_empty[buf], ((sl >> 1) & 1) ^ 1);
            fence_after_sync();
            const uint32_t a_hi = smem_u32(sA + s * 2 * ST_A_TILE), a_lo = a_hi + ST_A_TILE;
            const uint32_t dcol = tmem_base + buf * 128 + d * NB;
            // corrections first (tiny partial sums), the main product last; the correction that needs only what
            // has landed (A_hi = the raw tile, B_lo precomputed) goes before the wait for the derived low plane
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_tf32(dcol, make_desc(a_hi + KSTEP_BYTES * j), make_desc(b_lo + KSTEP_BYTES * j), idesc, j != 0);
            if (!POOL) {
              mbar_wait(&split[s], par);
              fence_after_sync();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_tf32(dcol, make_desc(a_lo + KSTEP_BYTES * j), make_desc(b_hi + KSTEP_BYTES * j), idesc, 1);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_tf32(dcol, make_desc(a_hi + KSTEP_BYTES * j), make_desc(b_hi + KSTEP_BYTES * j), idesc, 1);
            umma_commit(&empty[s]);
          }
          umma_commit(&tmem_full[buf]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (thread = staged position / output m)
    const int i = tid;                       // TMEM lane = staged row jl = i; also output index: m = m0 + i (i < OUT)
    const int m = m0 + i;
    const int mtot = (a.F + STRIDE - 1) / STRIDE;
    const bool mok = i < OUT && m < mtot;
    float bo[NSRC];
#pragma unroll
    for (int o = 0; o < NSRC; ++o) bo[o] = __ldg(a.bout + o);
    const float inv_ov1 = a.overlap > 1 ? 1.0f / (float)(a.overlap - 1) : 0.f;
    int sl = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int k_lo, k_hi;
      slot_range(t, a.tc, step, a.P, k_lo, k_hi);
      float macc[NSRC][STRIDE];
#pragma unroll
      for (int o = 0; o < NSRC; ++o)
#pragma unroll
        for (int r = 0; r < STRIDE; ++r) macc[o][r] = 0.f;
      for (int k = k_lo; k <= k_hi; ++k, ++sl) {
        const int buf = sl & 1, p = t - k * step;
        mbar_wait_relaxed(&tmem_full[buf], (sl >> 1) & 1);
        fence_after_sync();
        float Y[NSRC][STRIDE];
#pragma unroll
        for (int z = 0; z < NSRC; ++z) {
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 128 + z * 32;
          float zr[32];
          tmem_ld16_nowait(taddr, zr);
          tmem_ld16_nowait(taddr + 16, zr + 16);
          tmem_wait_ld();
          if (z == NSRC - 1) {   // every accumulator column of this slot is in registers: hand the buffer back
            fence_before_sync();
            mbar_arrive(&tmem_empty[buf]);
          }
          float* zs = Zs + (z & 1) * ST_ROWS * ST_ZP;
#pragma unroll
          for (int c = 0; c < 32; ++c) zs[i * ST_ZP + c] = zr[c];
          asm volatile("bar.sync 1, 128;" ::: "memory");
          // Y[m][r] = sum_dd Z[m - dd][(dd, r)]; staged row of position m - dd is i + (ND-1) - dd
#pragma unroll
          for (int r = 0; r < STRIDE; ++r) {
            float y = 0.f;
            if (i < OUT) {
#pragma unroll
              for (int dd = 0; dd < ND; ++dd) y += zs[(i + ND - 1 - dd) * ST_ZP + dd * STRIDE + r];
            }
            Y[z][r] = y;
          }
        }
        // bias + ReLU + ratio mask across the sources + sequential cross-fade (separate_bach10.py:245-266,
        // separate_ikala.py:207-217; overlapadd_multi separate_dsd.py:139-169)
        const float up = k == k_lo ? 1.f : (float)p * inv_ov1;
        const float down = k == k_lo ? 0.f : (float)(a.overlap - 1 - p) * inv_ov1;
#pragma unroll
        for (int r = 0; r < STRIDE; ++r) {
          float pv[NSRC], tot = 0.f;
#pragma unroll
          for (int o = 0; o < NSRC; ++o) {
            pv[o] = fmaxf(Y[o][r] + bo[o], 0.f);
            tot += pv[o];
          }
          const bool pos = tot > 0.f;
          const float rr = pos ? __fdividef(up, tot) : 0.f;
          const float q = (pos || RULE == 1) ? 0.f : up / (float)NSRC;
#pragma unroll
          for (int o = 0; o < NSRC; ++o) macc[o][r] = fmaf(down, macc[o][r], fmaf(pv[o], rr, q));
        }
      }
      if (mok) {
#pragma unroll
        for (int r = 0; r < STRIDE; ++r) {
          const int b = STRIDE * m + r;
          if (b < a.F) {
            const int64_t o = (int64_t)t * a.ldf + b;
            const float2 x = a.X[o];
#pragma unroll
            for (int s = 0; s < NSRC; ++s) a.S[o + s * a.src_stride] = make_float2(macc[s][r] * x.x, macc[s][r] * x.y);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    fence_after_sync();
    tmem_dealloc(tmem_base, ST_TMEM_COLS);
  }
}

template <int STRIDE, int ND, int NSRC, int NDEC, int RULE, int POOL, int NW>
static int launch_sconv_tc_t(dcs_ctx* ctx, const SconvMaskArgs& a, cudaStream_t st) {
  using TL = SconvTile<STRIDE, ND, NSRC, NDEC, NW>;
  auto kern = sconv_mask_tc_kernel<STRIDE, ND, NSRC, NDEC, RULE, POOL, NW>;
  DCS_TRY(ensure_smem_attr(kern, TL::SMEM));
  const int mtot = (a.F + STRIDE - 1) / STRIDE;
  const int mtiles = (mtot + TL::OUT - 1) / TL::OUT;
  const int per_sm = TL::SMEM <= 110 * 1024 ? 2 : 1;
  int chunks = (ctx->num_sms * per_sm) / mtiles;
  if (chunks < 1) chunks = 1;
  if (chunks > a.T) chunks = a.T;
  const int fpc = (a.T + chunks - 1) / chunks;
  dim3 grid((unsigned)mtiles, (unsigned)((a.T + fpc - 1) / fpc));
  alignas(64) CUtensorMap tmG;
  memset(&tmG, 0, sizeof(tmG));
  if (!POOL) {
    // G: [P * NDEC * tc rows][J positions][32 channels]; one box = 128 positions of one row
    DCS_TRY(tma_encode_3d_f32(&tmG, a.G, 32, (uint64_t)a.J, (uint64_t)a.P * NDEC * a.tc, 128, (uint64_t)a.J * 128, ST_ROWS));
  }
  kern<<<grid, ST_THREADS, TL::SMEM, st>>>(a, tmG, fpc);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

bool sconv_mask_tc_supported(const SconvMaskArgs& a) {
  const int step = a.tc - a.overlap;
  return step > 0 && ((uintptr_t)a.G % 16 == 0) &&
         (a.arch == DCS_ARCH_BACH10 || a.arch == DCS_ARCH_BACH10_SCORE || a.arch == DCS_ARCH_IKALA || a.arch == DCS_ARCH_IKALA_NOPOOL);
}

int launch_sconv_mask_tc(dcs_ctx* ctx, const SconvMaskArgs& a, cudaStream_t st) {
  if (a.T <= 0) return DCS_OK;
  DCS_REQUIRE(sconv_mask_tc_supported(a), "sconv_mask_tc: unsupported shape");
  if (a.arch == DCS_ARCH_BACH10) return launch_sconv_tc_t<4, 8, 4, 4, 1, 0, 1>(ctx, a, st);
  if (a.arch == DCS_ARCH_BACH10_SCORE) return launch_sconv_tc_t<4, 8, 4, 1, 1, 0, 4>(ctx, a, st);
  if (a.arch == DCS_ARCH_IKALA) return launch_sconv_tc_t<3, 10, 2, 2, 0, 4, 1>(ctx, a, st);
  return launch_sconv_tc_t<3, 10, 2, 2, 0, 0, 1>(ctx, a, st);
}

}  // namespace dcs

// dsd_tc.cu -- K3 on the tensor cores: InverseLayer(conv1) for every (patch, decoder) covering a
// frame as ONE tcgen05 GEMM per tile, with bias + ReLU + soft ratio mask + patch cross-fade +
// mixture-phase re-apply fused into the TMEM epilogue.  Same math as dsd.cu (reference:
// separate_dsd.py:212-234, :258-271, :139-169, :304), which remains the path for
// (time_context, overlap) settings with more than 6 patches per frame.
//
// GEMM view (D = A * B^T, fp32-accurate 3xTF32):
//   M = frequency bins  -> TMEM lanes (one epilogue thread per bin; its 18 values per frame are
//                           thread-local, so mask + cross-fade need no shuffles)
//   N = (frame, patch slot, decoder) = 8 x 6 x 3 (or 6 x 6 x 4, the stereo net) = 144 columns per tile
//   K = conv1 filters (50, padded to 56 = 7 k-steps)
//   A = W1t tile [128 bins][K]  (weights; split hi/lo into shared memory ONCE per CTA)
//   B = G rows   [144][K]       (decoder activations).  The transposed conv2 writes G frame-major
//       (GemmDesc fm_*: row ((t*6 + slot)*3 + decoder)), so the 144 rows of a group are ONE
//       contiguous box: two cp.async.bulk.tensor loads (k 0-31, k 32-63; columns >= 52 are the copy
//       engine's zero fill) deliver the raw fp32 tile = the HIGH operand (the tf32 datapath
//       truncates), four warps derive the LOW plane in shared memory.
// Persistent CTAs: a CTA owns one 128-bin tile and a contiguous range of 8-frame groups.
//   warps 0-15 epilogue | warp 16 MMA issue + TMEM alloc | warp 17 TMA | warps 18-21 low plane
// (16 epilogue warps: a single warp per scheduler runs the dependent mask arithmetic at
//  IPC ~0.2 -- measured, profiles/r1_notes.md -- so each scheduler gets four.)
// Double-buffered B stages and TMEM accumulators: the MMAs of group g+1 overlap the epilogue
// of group g.  Within a tile the 14 small correction MMAs (Alo*Bhi, Ahi*Blo) are issued before
// the 7 main ones so the truncating TMEM accumulation only sees 7 large addends.
#include "common.cuh"
#include "tc.cuh"

namespace dcs {

using namespace tc;

constexpr int MT_BINS = 128;             // bins per CTA tile
constexpr int MT_SLOTS = 6;              // patch slots per frame
constexpr int MT_COLS = 144;             // GEMM columns per group = frames x 6 slots x decoders
constexpr int MT_C1 = 50;
constexpr int MT_KSTEPS = 7;             // ceil(50 / 8)
constexpr int MT_SPLIT_WARPS = 4;
constexpr int MT_SPLIT = MT_SPLIT_WARPS * 32;            // 128 low-plane threads
constexpr int MT_A_SUB = MT_BINS * ROW_BYTES;      // 16 KB: [128][32] fp32
constexpr int MT_B_SUB = MT_COLS * ROW_BYTES;      // 18 KB: [144][32] fp32
constexpr int MT_A_BYTES = 4 * MT_A_SUB;           // hi k0-31, hi k32-63, lo k0-31, lo k32-63
constexpr int MT_B_STAGE = 4 * MT_B_SUB;           // same four planes
constexpr int MT_BAR_OFF = MT_A_BYTES + 2 * MT_B_STAGE;
constexpr int MT_XF_OFF = MT_BAR_OFF + 128;        // cross-fade coefficient tables, 12 float4 per epilogue warp
constexpr int MT_SMEM = MT_XF_OFF + 16 * 12 * 16 + 1024;  // + alignment slack
constexpr uint32_t MT_TMEM_COLS = 512;

// NDEC = 3: the DSD100 / hiphopss net (4th output = decoder 2 with its own bias, all-zero bins get 1/4 each,
//           separate_dsd.py:228,258-266), 8 frames per group;
// NDEC = 4: the stereo / ILD net, one launch per input channel (one decoder per source, all-zero bins get 0,
//           trainCNN_ILD_DSD100.py:99-106,183-186), 6 frames per group -- the same 144-column tile either way.
template <int NDEC>
struct MaskTile {
  static constexpr int FRAMES = MT_COLS / (MT_SLOTS * NDEC);    // 8 or 6
  static constexpr int EPI_WARPS = 2 * FRAMES;                  // 4 TMEM lane quadrants x FRAMES/2 frame pairs
  static constexpr int TMA_WARP = EPI_WARPS + 1;
  static constexpr int THREADS = (EPI_WARPS + 2 + MT_SPLIT_WARPS) * 32;   // 704 or 576
  static constexpr int VALS = MT_SLOTS * NDEC;                  // accumulator columns per frame: 18 or 24
};

template <int NDEC>
__global__ void __launch_bounds__(MaskTile<NDEC>::THREADS, 1)
dsd_mask_tc_kernel(const DsdMaskArgs a, const __grid_constant__ CUtensorMap tmG, const float4* __restrict__ xtab,
                   int groups_per_cta, int num_groups) {
  using MT = MaskTile<NDEC>;
  constexpr int MT_FRAMES = MT::FRAMES, MT_EPI_WARPS = MT::EPI_WARPS, MT_TMA_WARP = MT::TMA_WARP, VALS = MT::VALS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + MT_A_BYTES;
  uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + MT_BAR_OFF);   // raw B tile landed
  uint64_t* split_b = full_b + 2;                                       // low plane written
  uint64_t* empty_b = split_b + 2;                                      // MMAs of the stage retired
  uint64_t* tmem_full = empty_b + 2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bin0 = blockIdx.x * MT_BINS;
  const int g_begin = blockIdx.y * groups_per_cta;
  const int g_end = min(num_groups, g_begin + groups_per_cta);
  const int step = a.tc - a.overlap;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&split_b[s], MT_SPLIT);
      mbar_init(&empty_b[s], 1);
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], MT_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == MT_EPI_WARPS) tmem_alloc(tmem_slot, MT_TMEM_COLS);
  if (warp == MT_TMA_WARP && lane == 0) prefetch_tensormap(&tmG);
  // A tile: thread = bin row (threads 0..127), W1t is [c][bin] so the reads are coalesced over bins
  if (tid < MT_BINS) {
    const int b = bin0 + tid;
    const bool ok = b < a.F;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      float e[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = 4 * c4 + i;
        e[i] = (ok && c < MT_C1) ? __ldg(a.W1t + (int64_t)c * a.ldw + b) : 0.f;
      }
      float4 hi, lo;
      split4(make_float4(e[0], e[1], e[2], e[3]), hi, lo);
      const uint32_t off = (c4 >> 3) * MT_A_SUB + tile_off(tid, c4 & 7);
      *reinterpret_cast<float4*>(sA + off) = hi;
      *reinterpret_cast<float4*>(sA + 2 * MT_A_SUB + off) = lo;
    }
    fence_proxy_async();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == MT_TMA_WARP) {
    // ------------------------------------------------------------------ copy engine
    if (elect_one()) {
      for (int g = g_begin; g < g_end; ++g) {
        const int it = g - g_begin, s = it & 1;
        mbar_wait(&empty_b[s], ((it >> 1) & 1) ^ 1);
        uint8_t* st = sB + s * MT_B_STAGE;
        mbar_arrive_expect_tx(&full_b[s], 2 * MT_B_SUB);
        tma_load_2d(st, &tmG, &full_b[s], 0, g * MT_COLS);                 // k 0..31 of the group's 144 rows
        tma_load_2d(st + MT_B_SUB, &tmG, &full_b[s], KSTAGE, g * MT_COLS);  // k 32..63 (>= 52: zero fill)
      }
    }
  } else if (warp > MT_TMA_WARP) {
    // ------------------------------------------------------------------ low-plane writers
    const int pt = tid - (MT_TMA_WARP + 1) * 32;  // 0..127
    constexpr int CHUNKS = 2 * MT_B_SUB / 16 / MT_SPLIT;   // 18 float4 per thread
    for (int g = g_begin; g < g_end; ++g) {
      const int it = g - g_begin, s = it & 1;
      mbar_wait(&full_b[s], (it >> 1) & 1);
      const float4* raw = reinterpret_cast<const float4*>(sB + s * MT_B_STAGE);
      float4* lo = reinterpret_cast<float4*>(sB + s * MT_B_STAGE + 2 * MT_B_SUB);
#pragma unroll 6
      for (int u = 0; u < CHUNKS; ++u) {
        float4 h, l;
        split4(raw[u * MT_SPLIT + pt], h, l);
        lo[u * MT_SPLIT + pt] = l;
      }
      fence_proxy_async();
      mbar_arrive(&split_b[s]);
    }
  } else if (warp == MT_EPI_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(MT_BINS, MT_COLS);
      const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + 2 * MT_A_SUB;
      for (int g = g_begin; g < g_end; ++g) {
        const int it = g - g_begin, s = it & 1;
        const uint32_t par = (it >> 1) & 1;
        mbar_wait(&full_b[s], par);
        mbar_wait(&tmem_empty[s], par ^ 1);
        fence_after_sync();
        const uint32_t b_hi = smem_u32(sB + s * MT_B_STAGE), b_lo = b_hi + 2 * MT_B_SUB;
        const uint32_t dcol = tmem_base + s * 256;
        // corrections first (tiny partial sums), then the 7 main products; the correction that needs
        // only what the copy engine delivered (Alo * Bhi, Bhi = the raw tile) goes before the wait for
        // the derived low plane
#pragma unroll
        for (int j = 0; j < MT_KSTEPS; ++j) {
          const uint32_t ao = (j >> 2) * MT_A_SUB + KSTEP_BYTES * (j & 3), bo = (j >> 2) * MT_B_SUB + KSTEP_BYTES * (j & 3);
          umma_tf32(dcol, make_desc(a_lo + ao), make_desc(b_hi + bo), idesc, j != 0);
        }
        mbar_wait(&split_b[s], par);
        fence_after_sync();
#pragma unroll
        for (int j = 0; j < MT_KSTEPS; ++j) {
          const uint32_t ao = (j >> 2) * MT_A_SUB + KSTEP_BYTES * (j & 3), bo = (j >> 2) * MT_B_SUB + KSTEP_BYTES * (j & 3);
          umma_tf32(dcol, make_desc(a_hi + ao), make_desc(b_lo + bo), idesc, 1);
        }
#pragma unroll
        for (int j = 0; j < MT_KSTEPS; ++j) {
          const uint32_t ao = (j >> 2) * MT_A_SUB + KSTEP_BYTES * (j & 3), bo = (j >> 2) * MT_B_SUB + KSTEP_BYTES * (j & 3);
          umma_tf32(dcol, make_desc(a_hi + ao), make_desc(b_hi + bo), idesc, 1);
        }
        umma_commit(&empty_b[s]);
        umma_commit(&tmem_full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (thread = bin)
    // 16 warps: warp e reads TMEM lane quadrant e%4 (bins) and handles frames 2*(e/4), 2*(e/4)+1
    // of every group; the two frames are evaluated interleaved for instruction-level parallelism.
    const int quad = warp & 3, fsub = warp >> 2;
    const int b = bin0 + quad * 32 + lane;
    const bool bok = b < a.F;
    const float bo0 = __ldg(a.bout + 0), bo1 = __ldg(a.bout + 1), bo2 = __ldg(a.bout + 2), bo3 = __ldg(a.bout + 3);
    float4* xf = reinterpret_cast<float4*>(smem + MT_XF_OFF) + warp * 12;
    for (int g = g_begin; g < g_end; ++g) {
      const int it = g - g_begin, s = it & 1;
      const int tA = g * MT_FRAMES + 2 * fsub;
      // cross-fade coefficients of this warp's 2 frames x 6 patch slots (dsd_xfade_table_kernel):
      // acc <- down*acc + up*mask, (up, down) = (1, 0) for the first covering patch, the linspace
      // ramp for later ones, (0, 1) for empty slots -- the evaluation below is branch free
      if (lane < 12) xf[lane] = __ldg(xtab + (int64_t)tA * MT_SLOTS + lane);
      float2 xs[2];
#pragma unroll
      for (int ff = 0; ff < 2; ++ff) {
        const int t = tA + ff;
        xs[ff] = (bok && t < a.T) ? a.X[(int64_t)t * a.ldf + b] : make_float2(0.f, 0.f);
      }
      __syncwarp();
      mbar_wait_relaxed(&tmem_full[s], (it >> 1) & 1);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + s * 256 + 2 * VALS * fsub;
      float y[2][VALS];
      tmem_ld16_nowait(taddr, y[0]);
      if (NDEC == 3) tmem_ld2_nowait(taddr + 16, y[0] + 16); else tmem_ld8_nowait(taddr + 16, y[0] + 16);
      tmem_ld16_nowait(taddr + VALS, y[1]);
      if (NDEC == 3) tmem_ld2_nowait(taddr + VALS + 16, y[1] + 16); else tmem_ld8_nowait(taddr + VALS + 16, y[1] + 16);
      tmem_wait_ld();
      // the accumulator values are in registers: hand the TMEM buffer back right away
      fence_before_sync();
      mbar_arrive(&tmem_empty[s]);
      float acc[2][4];
#pragma unroll
      for (int ff = 0; ff < 2; ++ff) acc[ff][0] = acc[ff][1] = acc[ff][2] = acc[ff][3] = 0.f;
#pragma unroll
      for (int j = 0; j < MT_SLOTS; ++j) {
#pragma unroll
        for (int ff = 0; ff < 2; ++ff) {
          const float4 c = xf[ff * 6 + j];   // (up, down, up/4, -)
          const float p0 = fmaxf(y[ff][NDEC * j + 0] + bo0, 0.f), p1 = fmaxf(y[ff][NDEC * j + 1] + bo1, 0.f);
          const float p2 = fmaxf(y[ff][NDEC * j + 2] + bo2, 0.f);
          const float p3 = fmaxf(y[ff][NDEC * j + (NDEC == 3 ? 1 : 3)] + bo3, 0.f);   // DSD100: decoder 2 again (separate_dsd.py:228)
          const float tot = (p0 + p1) + (p2 + p3);
          const bool pos = tot > 1.2e-38f;
          float rc;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(tot));
          const float r = pos ? c.x * rc : 0.f;        // up * mask = p * (up / tot)
          const float q = (pos || NDEC == 4) ? 0.f : c.z;   // all-zero bin: 1/4 each (DSD100 rule); 0 (ILD rule)
          acc[ff][0] = fmaf(c.y, acc[ff][0], fmaf(p0, r, q));
          acc[ff][1] = fmaf(c.y, acc[ff][1], fmaf(p1, r, q));
          acc[ff][2] = fmaf(c.y, acc[ff][2], fmaf(p2, r, q));
          acc[ff][3] = fmaf(c.y, acc[ff][3], fmaf(p3, r, q));
        }
      }
#pragma unroll
      for (int ff = 0; ff < 2; ++ff) {
        const int t = tA + ff;
        if (bok && t < a.T) {
          const int64_t o = (int64_t)t * a.ldf + b;
          const float2 x = xs[ff];
          a.S[o] = make_float2(acc[ff][0] * x.x, acc[ff][0] * x.y);
          a.S[o + a.src_stride] = make_float2(acc[ff][1] * x.x, acc[ff][1] * x.y);
          a.S[o + 2 * a.src_stride] = make_float2(acc[ff][2] * x.x, acc[ff][2] * x.y);
          a.S[o + 3 * a.src_stride] = make_float2(acc[ff][3] * x.x, acc[ff][3] * x.y);
        }
      }
      __syncwarp();   // xf is rewritten at the top of the next iteration
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == MT_EPI_WARPS) {
    fence_after_sync();
    tmem_dealloc(tmem_base, MT_TMEM_COLS);
  }
}

// (up, down, up/4, 0) of every (frame, patch slot): the sequential cross-fade of overlapadd_multi
// (separate_dsd.py:139-169) as a per-slot recurrence; frames >= T (padding to whole groups) and slots
// without a patch get (0, 1): they leave the accumulated masks untouched.
__global__ void dsd_xfade_table_kernel(float4* __restrict__ tab, int T, int Tpad, int P, int tc, int overlap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Tpad * MT_SLOTS) return;
  const int t = i / MT_SLOTS, j = i - t * MT_SLOTS;
  const int step = tc - overlap;
  const float inv_ov1 = overlap > 1 ? 1.0f / (float)(overlap - 1) : 0.f;
  int k_lo = t - tc + 1;
  k_lo = k_lo > 0 ? (k_lo + step - 1) / step : 0;
  int k_hi = t / step;
  if (k_hi > P - 1) k_hi = P - 1;
  float up = 0.f, down = 1.f;
  if (t < T && k_lo + j <= k_hi) {
    const int p = t - (k_lo + j) * step;
    up = j == 0 ? 1.f : (float)p * inv_ov1;
    down = j == 0 ? 0.f : (float)(overlap - 1 - p) * inv_ov1;
  }
  tab[i] = make_float4(up, down, 0.25f * up, 0.f);
}

bool dsd_mask_tc_supported(const DsdMaskArgs& a) {
  const int step = a.tc - a.overlap;
  return step > 0 && (a.ndec == 3 || a.ndec == 4) && (a.tc + step - 1) / step <= MT_SLOTS && a.ldg % 4 == 0 && a.ldg >= 52 &&
         a.ldg <= 64 && ((uintptr_t)a.G % 16 == 0);
}

// all F bins; the last 128-bin tile holds only the Nyquist bin (F = 2^k + 1)
template <int NDEC>
static int launch_dsd_mask_tc_t(dcs_ctx* ctx, const DsdMaskArgs& a, cudaStream_t st) {
  using MT = MaskTile<NDEC>;
  DCS_TRY(ensure_smem_attr(dsd_mask_tc_kernel<NDEC>, MT_SMEM));
  const int m_tiles = (a.F + MT_BINS - 1) / MT_BINS;
  const int num_groups = (a.T + MT::FRAMES - 1) / MT::FRAMES;
  int chunks = ctx->num_sms / m_tiles;
  if (chunks < 1) chunks = 1;
  if (chunks > num_groups) chunks = num_groups;
  const int gpc = (num_groups + chunks - 1) / chunks;
  dim3 grid((unsigned)m_tiles, (unsigned)((num_groups + gpc - 1) / gpc));
  // G is frame-major here: [T * 6 slots * NDEC decoders][ldg] (see GemmDesc fm_*)
  alignas(64) CUtensorMap tmG;
  DCS_TRY(tma_encode_2d_f32(&tmG, a.G, (uint64_t)a.ldg, (uint64_t)a.T * MT_SLOTS * NDEC, (uint64_t)a.ldg * 4, MT_COLS));
  const int Tpad = num_groups * MT::FRAMES;
  DCS_TRY(ctx->net[11].ensure((size_t)Tpad * MT_SLOTS * sizeof(float4), st));
  float4* xtab = ctx->net[11].as<float4>();
  dsd_xfade_table_kernel<<<(unsigned)ceil_div64((int64_t)Tpad * MT_SLOTS, 256), 256, 0, st>>>(xtab, a.T, Tpad, a.P, a.tc, a.overlap);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  dsd_mask_tc_kernel<NDEC><<<grid, MT::THREADS, MT_SMEM, st>>>(a, tmG, xtab, gpc, num_groups);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

int launch_dsd_mask_tc(dcs_ctx* ctx, const DsdMaskArgs& a, cudaStream_t st) {
  if (a.T <= 0) return DCS_OK;
  DCS_REQUIRE(dsd_mask_tc_supported(a), "dsd_mask_tc: unsupported shape");
  return a.ndec == 4 ? launch_dsd_mask_tc_t<4>(ctx, a, st) : launch_dsd_mask_tc_t<3>(ctx, a, st);
}

}  // namespace dcs

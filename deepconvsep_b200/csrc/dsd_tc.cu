// dsd_tc.cu -- K3 on the tensor cores: InverseLayer(conv1) for every (patch, decoder) covering a
// frame as ONE tcgen05 GEMM per tile, with bias + ReLU + soft ratio mask + patch cross-fade +
// mixture-phase re-apply fused into the TMEM epilogue.  Same math as dsd.cu (reference:
// separate_dsd.py:212-234, :258-271, :139-169, :304), which remains the path for
// (time_context, overlap) settings with more than 6 patches per frame.
//
// GEMM view (D = A * B^T, fp32-accurate 3xTF32):
//   M = frequency bins  -> TMEM lanes (one epilogue thread per bin; its 18 values per frame are
//                           thread-local, so mask + cross-fade need no shuffles)
//   N = (frame, patch slot, decoder) = 8 x 6 x 3 = 144 columns per tile
//   K = conv1 filters (50, padded to 56 = 7 k-steps)
//   A = W1t tile [128 bins][K]  (weights; split hi/lo into shared memory ONCE per CTA)
//   B = G rows   [144][K]       (decoder activations; loaded, split and staged per tile)
// Persistent CTAs: a CTA owns one 128-bin tile and a contiguous range of 8-frame groups.
//   warps 0-15 epilogue | warp 16 MMA issue + TMEM alloc | warps 17-24 B producers
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
constexpr int MT_FRAMES = 8;             // frames per group
constexpr int MT_SLOTS = 6;              // patch slots per frame
constexpr int MT_COLS = MT_FRAMES * MT_SLOTS * 3;  // 144
constexpr int MT_C1 = 50;
constexpr int MT_KSTEPS = 7;             // ceil(50 / 8)
constexpr int MT_EPI_WARPS = 16;          // 4 per TMEM lane quadrant, 2 frames of a group each
constexpr int MT_PROD_WARPS = 8;
constexpr int MT_PROD = MT_PROD_WARPS * 32;             // 256 producer threads
constexpr int MT_THREADS = (MT_EPI_WARPS + 1 + MT_PROD_WARPS) * 32;  // 800
constexpr int MT_A_SUB = MT_BINS * ROW_BYTES;      // 16 KB: [128][32] fp32
constexpr int MT_B_SUB = MT_COLS * ROW_BYTES;      // 18 KB: [144][32] fp32
constexpr int MT_A_BYTES = 4 * MT_A_SUB;           // hi k0-31, hi k32-63, lo k0-31, lo k32-63
constexpr int MT_B_STAGE = 4 * MT_B_SUB;           // same four planes
constexpr int MT_BAR_OFF = MT_A_BYTES + 2 * MT_B_STAGE;
constexpr int MT_TAB_OFF = MT_BAR_OFF + 128;       // int64 source-row offsets of the 144 B rows
constexpr int MT_XF_OFF = MT_TAB_OFF + 2 * MT_COLS * 8;   // cross-fade coefficient tables, 12 float4 per epilogue warp
constexpr int MT_SMEM = MT_XF_OFF + MT_EPI_WARPS * 12 * 16 + 1024;  // + alignment slack
constexpr uint32_t MT_TMEM_COLS = 512;

__global__ void __launch_bounds__(MT_THREADS, 1)
dsd_mask_tc_kernel(const DsdMaskArgs a, int groups_per_cta, int num_groups) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + MT_A_BYTES;
  uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + MT_BAR_OFF);
  uint64_t* empty_b = full_b + 2;
  uint64_t* tmem_full = empty_b + 2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int64_t* row_src = reinterpret_cast<int64_t*>(smem + MT_TAB_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bin0 = blockIdx.x * MT_BINS;
  const int g_begin = blockIdx.y * groups_per_cta;
  const int g_end = min(num_groups, g_begin + groups_per_cta);
  const int step = a.tc - a.overlap;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_b[s], MT_PROD);
      mbar_init(&empty_b[s], 1);
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], MT_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == MT_EPI_WARPS) tmem_alloc(tmem_slot, MT_TMEM_COLS);
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

  if (warp > MT_EPI_WARPS) {
    // ------------------------------------------------------------------ B producers
    const int pt = tid - (MT_EPI_WARPS + 1) * 32;  // 0..255
    // Each group stages 144 rows x 16 chunks of 16 B (13 real, the rest zero) = 9 chunks per
    // thread: 8 consecutive threads read one row's consecutive chunks (coalesced) and write 8
    // distinct swizzled slots.  The loads of group g+1 are in flight while group g is split,
    // stored and consumed; the source-row table is double buffered (one bar.sync per group).
    constexpr int CPT = MT_COLS * 16 / MT_PROD;  // 9
    auto build_table = [&](int g, int64_t* tab) {
      for (int rr = pt; rr < MT_COLS; rr += MT_PROD) {
        const int f = rr / 18, rem = rr - f * 18, j = rem / 3, d = rem - j * 3;
        const int t = g * MT_FRAMES + f;
        int64_t src = -1;
        if (t < a.T) {
          int k_lo = t - a.tc + 1;
          k_lo = k_lo > 0 ? (k_lo + step - 1) / step : 0;
          int k_hi = t / step;
          if (k_hi > a.P - 1) k_hi = a.P - 1;
          const int k = k_lo + j;
          if (k <= k_hi) src = ((int64_t)(k * 3 + d) * a.tc + (t - k * step)) * a.ldg;
        }
        tab[rr] = src;
      }
    };
    float4 v[CPT];
    auto issue_loads = [&](const int64_t* tab) {
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        const int idx = u * MT_PROD + pt;
        const int rr = idx >> 4, c4 = idx & 15;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < 13) {
          const int64_t src = tab[rr];
          if (src >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(a.G + src) + c4);
        }
      }
    };
    if (g_begin < g_end) {
      build_table(g_begin, row_src);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      issue_loads(row_src);
    }
    for (int g = g_begin; g < g_end; ++g) {
      const int it = g - g_begin, s = it & 1;
      if (g + 1 < g_end) build_table(g + 1, row_src + ((it + 1) & 1) * MT_COLS);
      mbar_wait_relaxed(&empty_b[s], ((it >> 1) & 1) ^ 1);
      uint8_t* st = sB + s * MT_B_STAGE;
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        const int idx = u * MT_PROD + pt;
        const int rr = idx >> 4, c4 = idx & 15;
        if (c4 == 12) { v[u].z = 0.f; v[u].w = 0.f; }  // columns 50, 51 of the padded G row
        float4 hi, lo;
        split4(v[u], hi, lo);
        const uint32_t off = (c4 >> 3) * MT_B_SUB + tile_off(rr, c4 & 7);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + 2 * MT_B_SUB + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(&full_b[s]);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // next table complete, this one no longer read
      if (g + 1 < g_end) issue_loads(row_src + ((it + 1) & 1) * MT_COLS);
    }
  } else if (warp == MT_EPI_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
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
        // corrections first (tiny partial sums), then the 7 main products
#pragma unroll
        for (int j = 0; j < MT_KSTEPS; ++j) {
          const uint32_t ao = (j >> 2) * MT_A_SUB + KSTEP_BYTES * (j & 3), bo = (j >> 2) * MT_B_SUB + KSTEP_BYTES * (j & 3);
          umma_tf32(dcol, make_desc(a_lo + ao), make_desc(b_hi + bo), idesc, j != 0);
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
    const float inv_ov1 = a.overlap > 1 ? 1.0f / (float)(a.overlap - 1) : 0.f;
    float4* xf = reinterpret_cast<float4*>(smem + MT_XF_OFF) + warp * 12;
    for (int g = g_begin; g < g_end; ++g) {
      const int it = g - g_begin, s = it & 1;
      const int tA = g * MT_FRAMES + 2 * fsub;
      // cross-fade coefficients of this warp's 2 frames x 6 patch slots, one per lane 0..11:
      // acc <- down*acc + up*mask, (up, down) = (1, 0) for the first covering patch, the linspace
      // ramp for later ones, (0, 1) for empty slots -- the evaluation below is branch free
      if (lane < 12) {
        const int ff = lane / 6, j = lane - 6 * ff, t = tA + ff;
        int k_lo = t - a.tc + 1;
        k_lo = k_lo > 0 ? (k_lo + step - 1) / step : 0;
        int k_hi = t / step;
        if (k_hi > a.P - 1) k_hi = a.P - 1;
        float up = 0.f, down = 1.f;
        if (t < a.T && k_lo + j <= k_hi) {
          const int p = t - (k_lo + j) * step;
          up = j == 0 ? 1.f : (float)p * inv_ov1;
          down = j == 0 ? 0.f : (float)(a.overlap - 1 - p) * inv_ov1;
        }
        xf[lane] = make_float4(up, down, 0.25f * up, 0.f);
      }
      float2 xs[2];
#pragma unroll
      for (int ff = 0; ff < 2; ++ff) {
        const int t = tA + ff;
        xs[ff] = (bok && t < a.T) ? a.X[(int64_t)t * a.ldf + b] : make_float2(0.f, 0.f);
      }
      __syncwarp();
      mbar_wait_relaxed(&tmem_full[s], (it >> 1) & 1);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + s * 256 + 36 * fsub;
      float y[2][18];
      tmem_ld16_nowait(taddr, y[0]);
      tmem_ld2_nowait(taddr + 16, y[0] + 16);
      tmem_ld16_nowait(taddr + 18, y[1]);
      tmem_ld2_nowait(taddr + 34, y[1] + 16);
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
          const float p0 = fmaxf(y[ff][3 * j + 0] + bo0, 0.f), p1 = fmaxf(y[ff][3 * j + 1] + bo1, 0.f);
          const float p2 = fmaxf(y[ff][3 * j + 2] + bo2, 0.f), p3 = fmaxf(y[ff][3 * j + 1] + bo3, 0.f);
          const float tot = (p0 + p1) + (p2 + p3);
          const bool pos = tot > 1.2e-38f;
          float rc;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(tot));
          const float r = pos ? c.x * rc : 0.f;        // up * mask = p * (up / tot)
          const float q = pos ? 0.f : c.z;             // all-zero bin: 1/4 each
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

bool dsd_mask_tc_supported(const DsdMaskArgs& a) {
  const int step = a.tc - a.overlap;
  return step > 0 && (a.tc + step - 1) / step <= MT_SLOTS && a.ldg % 4 == 0 && a.ldg >= 52 && ((uintptr_t)a.G % 16 == 0);
}

// all F bins; the last 128-bin tile holds only the Nyquist bin (F = 2^k + 1)
int launch_dsd_mask_tc(dcs_ctx* ctx, const DsdMaskArgs& a, cudaStream_t st) {
  if (a.T <= 0) return DCS_OK;
  DCS_REQUIRE(dsd_mask_tc_supported(a), "dsd_mask_tc: unsupported shape");
  static bool attr = false;
  if (!attr) {
    DCS_CUDA(cudaFuncSetAttribute(dsd_mask_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM));
    attr = true;
  }
  const int m_tiles = (a.F + MT_BINS - 1) / MT_BINS;
  const int num_groups = (a.T + MT_FRAMES - 1) / MT_FRAMES;
  int chunks = ctx->num_sms / m_tiles;
  if (chunks < 1) chunks = 1;
  if (chunks > num_groups) chunks = num_groups;
  const int gpc = (num_groups + chunks - 1) / chunks;
  dim3 grid((unsigned)m_tiles, (unsigned)((num_groups + gpc - 1) / gpc));
  dsd_mask_tc_kernel<<<grid, MT_THREADS, MT_SMEM, st>>>(a, gpc, num_groups);
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  return DCS_OK;
}

}  // namespace dcs

// gemm_tma.cu -- the TMA-fed version of the fp32-accurate tensor-core GEMM (gemm_tc.cu):
// C = act(A*B + bias) with tcgen05.mma kind::tf32, 3xTF32 split, accumulators in TMEM.
//
// gemm_tc.cu stages operands through registers (any strided / segmented view works, but four
// producer warps per stage spend their time on address arithmetic, the hi/lo split and two
// shared stores per element).  Here the copy engine does the staging:
//   warp 5   lane 0: per k-block one cp.async.bulk.tensor.2d for the raw fp32 A tile (128 rows x
//            32 floats, SWIZZLE_128B = exactly the K-major operand layout of tc.cuh) and two for
//            the pre-split weight planes; completion is counted on the stage's `full` mbarrier
//   warps 0-3: when a stage lands, derive the LOW operand plane in shared memory
//            (lo = x - trunc_tf32(x), elementwise on the swizzled tile: same byte offsets) and
//            publish it to the async proxy; afterwards the epilogue (thread = output row)
//   warp 4   lane 0: 3 tcgen05.mma per k-step.  The HIGH A operand is the raw tile itself: the
//            tf32 datapath reads the top 19 bits of each fp32 word, i.e. trunc_tf32(x)
//            (checked on the device by tests/test_gpu_gemm.py; DCS_DEBUG_TMA=2 rewrites the
//            tile with explicitly truncated values instead).
// A views this kernel takes (everything else stays on gemm_tc.cu):
//  * "rows" (2-D tensor map): one K segment, offset(m) = (m / m_inner) * a_so + (m % m_inner) * a_si
//    (plain: m * lda) with a 16-byte aligned row pitch.  Row tiles never straddle two outer
//    indices u = m / m_inner (the tile grid is [u][ceil(m_inner / 128)]), so a tile is ONE box:
//    128 rows at column u * a_so + k.  This is InverseLayer(conv2) of the DSD100 net on the
//    zero-padded decoder activations (examples/dsd100/separate_dsd.py:214-217): u = output time
//    position, a_so = one time step; overlapping rows (pitch < K) are its conv2 / bottleneck.
//  * "conv" (4-D tensor map {32 channels, positions, time, patch x decoder}): the 2-D convolutions
//    of the 30-channel nets on channel-padded NHWC activations (examples/ikala/separate_ikala.py:
//    177,186-187; examples/bach10/separate_bach10.py:200,209-246).  One k-block = one filter tap
//    (q, w): the box {32, 128 positions, 1, 1} at (0, pos0 + w, u + q, kd) -- implicit GEMM, no
//    im2col and no per-row address arithmetic anywhere.
// Rows / positions / taps outside the tensor are zero-filled by the TMA unit.
#include <atomic>
#include "common.cuh"
#include "tc.cuh"

namespace dcs {

using namespace tc;

constexpr int TM_BM = 128;
constexpr int TM_SPLIT_WARPS = 4;
constexpr int TM_MMA_WARP = 4, TM_TMA_WARP = 5;
constexpr int TM_THREADS = 6 * 32;

template <int BN, int STAGES>
struct TmSmem {
  static constexpr int A_BYTES = TM_BM * ROW_BYTES;  // 16 KB
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A raw(=hi), A lo, B hi, B lo
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

struct TmaGemmArgs {
  int conv;           // 0: rows mode (2-D map), 1: conv mode (4-D map)
  // A tile = the box {32, pb, tb, kdb} of the map: tile row = pos + pb * (tl + tb * kdl), at most 128.
  // rows mode and wide conv layers: pb = 128, tb = kdb = 1; narrow conv layers (iKala: 40 / 21
  // positions) pack several planes (decoder) or several output times (encoder) into one tile
  int pb, tb, kdb;
  int n_pos;          // rows per (u, kd): positions (conv) / m_inner (rows mode; plain GEMM: M)
  int n_u, n_kd;      // output time positions; (patch, decoder) planes (1 outside the decoders)
  int tiles_pos;      // ceil(n_pos / pb)
  int col_per_u;      // rows mode: a_so (plain GEMM: 0)
  int kw;             // conv mode: filter taps along the position axis (k-blocks per time tap)
  int m_inner;        // row index m = u * m_inner + kd * n_pos + pos
  int Np;             // row offset of the low weight plane in the stacked weight tensor
  int acc_mode;
  int rewrite_hi;     // store trunc_tf32(x) over the raw tile (bring-up cross-check)
  int cvec;           // C rows allow 16-byte stores
  int probe;          // DCS_DEBUG_TMA_PROBE (timing experiments, wrong results): 1 no low-plane work, 2 no MMAs, 4 no B loads,
                      // 8 no A loads, 16 only the main product (1 of 3 MMAs per k-step)
  int prefetch;       // A boxes prefetched into L2 ahead of the stage ring (DCS_DEBUG_TMA_PREFETCH, default 6; 0 = off)
  int k_splits;       // > 1: blockIdx.z owns a slice of the k-blocks and stores raw partial sums
  int ldp;            // row pitch of the partial-sum workspace
  float* partial;
};

// bias / ReLU / store of 16 accumulator columns [n, n+16) of the C row at element offset `roff`.  `seg` (uniform per
// launch): the C columns are segmented (n_seg < N, the decoder dense layers scattering into the padded buffer) -- only
// then is the per-group column offset a division; the plain case costs one add.  (The epilogue of the persistent
// kernels ran 1 224 instructions per warp per 128x32 tile, mostly these divisions: 40 % of all instructions executed
// by Bach10's InverseLayer(conv2), whose tiles are only ~7 k-blocks long -- profiles/r2_notes.md.)
__device__ __forceinline__ void gemm_store16(const GemmDesc& d, bool cvec, bool seg, int64_t roff, int n, const float (&v)[16]) {
  int sq = 0, sr = 0;     // segment / column inside it of the chunk's first column: one division pair per 16 columns
  if (seg) { sq = n / d.n_seg; sr = n - sq * d.n_seg; }
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const int nn = n + 4 * i4;
    if (nn >= d.N) break;
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = v[4 * i4 + e];
    if (d.bias) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (nn + e < d.N) x[e] += __ldg(d.bias + nn + e);
    }
    if (d.relu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
    }
    if (cvec && nn + 4 <= d.N) {   // a group of 4 columns never straddles a C segment (launcher)
      int q = sq, r = sr + 4 * i4;
      while (seg && r >= d.n_seg) { r -= d.n_seg; ++q; }      // at most once or twice: n_seg >= 4
      const int64_t coff = seg ? (int64_t)q * d.n_ss + r : (int64_t)nn;
      *reinterpret_cast<float4*>(d.C + roff + coff) = make_float4(x[0], x[1], x[2], x[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (nn + e < d.N) {
          const int64_t coff = seg ? (int64_t)((nn + e) / d.n_seg) * d.n_ss + ((nn + e) % d.n_seg) : (int64_t)(nn + e);
          d.C[roff + coff] = x[e];
        }
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(TM_THREADS)
gemm_tma_kernel(const GemmDesc d, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const TmaGemmArgs g) {
  using SM = TmSmem<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);   // TMA landed
  uint64_t* split = full + STAGES;                                     // low plane written
  uint64_t* empty = split + STAGES;                                    // MMAs of the stage retired
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kdg = (g.n_kd + g.kdb - 1) / g.kdb;
  const int per_u = g.tiles_pos * n_kdg;
  // tile order: the output times u of one (plane, position tile) are adjacent -- they re-read the same few input rows
  // (a transposed convolution reads each input row once per tap), so the re-reads hit L2 instead of HBM.  (u-major order
  // made Bach10's InverseLayer(conv2) read 30 GB from DRAM for 1.5 GB of input: L2 hit rate 28 %, profiles/r2_notes.md)
  const int n_ug = (g.n_u + g.tb - 1) / g.tb;
  const int ug = blockIdx.x % n_ug, rest = blockIdx.x / n_ug;
  const int kdg = rest / g.tiles_pos;
  const int u = ug * g.tb, kd = kdg * g.kdb;                                   // first output time / plane of the tile
  const int r0 = (rest - kdg * g.tiles_pos) * g.pb;                          // first position / row
  const int n0 = blockIdx.y * BN;
  const uint32_t a_tx = (uint32_t)(g.pb * g.tb * g.kdb) * ROW_BYTES;         // bytes one A box delivers
  int kb_lo = 0, kb_hi = (d.K + KSTAGE - 1) / KSTAGE;
  if (d.kc_rows > 0) {   // taps that only see the zero padding are skipped (GemmDesc)
    const int q_lo = max(0, d.kc_pad - u), q_hi = min(d.kc_taps - 1, d.kc_pad + d.kc_n - 1 - u);
    kb_lo = (d.kc_unit * q_lo) / KSTAGE;
    kb_hi = min(kb_hi, (d.kc_unit * (q_hi + 1) + KSTAGE - 1) / KSTAGE);
    if (kb_hi <= kb_lo) kb_hi = kb_lo + 1;
  }
  if (g.k_splits > 1) {   // the launcher sizes the slices so that none is empty
    const int per = (kb_hi - kb_lo + g.k_splits - 1) / g.k_splits;
    kb_lo += blockIdx.z * per;
    kb_hi = min(kb_hi, kb_lo + per);
  }
  const int num_kb = kb_hi - kb_lo;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], TM_SPLIT_WARPS * 32);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == TM_TMA_WARP && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  constexpr uint32_t TMEM_COLS = 4 * BN;
  const int n_main = g.acc_mode == 2 ? 3 : 1;
  const int corr_acc = g.acc_mode == 0 ? 0 : n_main;
  const int n_main_used = min(n_main, num_kb * (KSTAGE / 8));
  if (warp == TM_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TM_TMA_WARP) {
    if (elect_one()) {
      // ---------------------------------------------------------------- copy engine
      // the activation boxes of the next PF k-blocks are prefetched into L2 while the stage ring (2 stages x 2 CTAs)
      // is busy: a tile's first touch of HBM is then off the TMA -> low plane -> MMA -> commit chain
      const int PF = g.prefetch;
      auto prefetch_a = [&](int kb) {
        if (g.conv) {
          const int q = (kb_lo + kb) / g.kw, w = (kb_lo + kb) - q * g.kw;
          tma_prefetch_4d(&tmA, 0, r0 + w, u + q, kd);
        } else {
          tma_prefetch_2d(&tmA, u * g.col_per_u + (kb_lo + kb) * KSTAGE, r0);
        }
      };
      for (int kb = STAGES; kb < min(num_kb, STAGES + PF); ++kb) prefetch_a(kb);
      int cq = kb_lo / g.kw, cw = kb_lo - cq * g.kw;     // running (time tap, position tap) of the conv view
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        if (PF > 0 && kb + STAGES + PF < num_kb) prefetch_a(kb + STAGES + PF);
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * SM::STAGE_BYTES;
        const int k0 = (kb_lo + kb) * KSTAGE;
        const bool ldA = !(g.probe & 8), ldB = !(g.probe & 4);
        mbar_arrive_expect_tx(&full[s], (ldA ? a_tx : 0) + (ldB ? 2 * SM::B_BYTES : 0));
        if (ldA) {
          if (g.conv) {
            tma_load_4d(st, &tmA, &full[s], 0, r0 + cw, u + cq, kd);
            if (++cw == g.kw) { cw = 0; ++cq; }
          } else {
            tma_load_2d(st, &tmA, &full[s], u * g.col_per_u + k0, r0);
          }
        }
        if (ldB) {
          tma_load_2d(st + 2 * SM::A_BYTES, &tmB, &full[s], k0, n0);
          tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &tmB, &full[s], k0, g.Np + n0);
        }
      }
    }
  } else if (warp == TM_MMA_WARP) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      // one thread, ~12 MMAs per k-block: everything the loop needs per MMA is one 32-bit add (descriptor low word)
      // -- no divisions, no 64-bit descriptor rebuilds; the accumulator rotation is a counter
      constexpr uint32_t idesc = make_idesc_tf32(TM_BM, BN);
      constexpr uint32_t ST_D = SM::STAGE_BYTES >> 4, A_D = SM::A_BYTES >> 4, B_D = SM::B_BYTES >> 4;
      const uint32_t lo0 = desc_lo(smem_u32(smem));
      const uint32_t t_corr = tmem_base + corr_acc * BN;
      const bool p_ab = !(g.probe & (2 | 16)), p_main = !(g.probe & 2);
      int ma = 0, ks = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t par = (kb / STAGES) & 1;
        mbar_wait(&full[s], par);
        fence_after_sync();
        const uint32_t a_hi = lo0 + s * ST_D, a_lo = a_hi + A_D, b_hi = a_hi + 2 * A_D, b_lo = b_hi + B_D;
        // the two products that only need what the copy engine delivered go first; the one with
        // the derived low plane of A follows when the writers are done (off the critical path)
#pragma unroll
        for (int j = 0; j < KSTAGE / 8; ++j, ++ks) {
          if (p_ab) umma_tf32(t_corr, desc_of(a_hi + KSTEP_DESC * j), desc_of(b_lo + KSTEP_DESC * j), idesc, ks != 0);
          if (p_main) umma_tf32(tmem_base + ma * BN, desc_of(a_hi + KSTEP_DESC * j), desc_of(b_hi + KSTEP_DESC * j), idesc,
                                corr_acc == 0 ? 1 : (ks >= n_main));
          ma = ma + 1 == n_main ? 0 : ma + 1;
        }
        mbar_wait(&split[s], par);
        fence_after_sync();
#pragma unroll
        for (int j = 0; j < KSTAGE / 8; ++j)
          if (p_ab) umma_tf32(t_corr, desc_of(a_lo + KSTEP_DESC * j), desc_of(b_hi + KSTEP_DESC * j), idesc, 1);
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ------------------------------------------------------------------ low-plane writers
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full[s], (kb / STAGES) & 1);
      float4* raw = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES + SM::A_BYTES);
      constexpr int CHUNKS = SM::A_BYTES / 16 / (TM_SPLIT_WARPS * 32);   // 8 per thread
      float4 x[CHUNKS];
      if (g.probe & 1) { mbar_arrive(&split[s]); continue; }
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) x[i] = raw[i * (TM_SPLIT_WARPS * 32) + tid];
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) {
        float4 h, l;
        split4(x[i], h, l);
        lo[i * (TM_SPLIT_WARPS * 32) + tid] = l;
        if (g.rewrite_hi) raw[i * (TM_SPLIT_WARPS * 32) + tid] = h;
      }
      fence_proxy_async();
      mbar_arrive(&split[s]);
    }
    // ------------------------------------------------------------------ epilogue
    // tile row -> (position, output time, plane) -> GEMM row index
    const int pos = r0 + tid % g.pb, t2 = tid / g.pb, tl = t2 % g.tb, kdl = t2 / g.tb;
    const int64_t m64 = (int64_t)(u + tl) * g.m_inner + (int64_t)(kd + kdl) * g.n_pos + pos;
    const bool m_ok = kdl < g.kdb && kd + kdl < g.n_kd && u + tl < g.n_u && pos < g.n_pos && m64 < d.M;
    const int m = m_ok ? (int)m64 : 0;
    mbar_wait_relaxed(tmem_full, 0);
    fence_after_sync();
    const int64_t roff = gemm_c_row_offset(d, m);
    const bool st_ok = m_ok && roff >= 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int j = 0; j < BN / 16; ++j) {
      if (n0 + 16 * j >= d.N) break;  // warp-uniform
      float v[16];
      tmem_ld16(taddr + 16 * j, v);
      for (int a = 1; a < n_main_used; ++a) {
        float w[16];
        tmem_ld16(taddr + a * BN + 16 * j, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
      }
      if (corr_acc) {
        float w[16];
        tmem_ld16(taddr + corr_acc * BN + 16 * j, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
      }
      if (st_ok && g.k_splits > 1) {   // raw partial sums; splitk_reduce adds bias / activation
        float4* dst = reinterpret_cast<float4*>(g.partial + ((int64_t)blockIdx.z * d.M + m) * g.ldp + n0 + 16 * j);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
          if (n0 + 16 * j + 4 * i4 < g.ldp) dst[i4] = make_float4(v[4 * i4], v[4 * i4 + 1], v[4 * i4 + 2], v[4 * i4 + 3]);
      } else if (st_ok) {
        gemm_store16(d, g.cvec != 0, d.n_seg < d.N, roff, n0 + 16 * j, v);
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == TM_MMA_WARP) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- persistent variant ------------------------------------------------------------------------
// Same tiles, same operand views, same accumulation plan -- but ONE CTA per SM walks a strided list of tiles with
//   * a stage ring that runs on across tile boundaries (no fill / drain per tile),
//   * two TMEM accumulator sets, so the epilogue of tile i runs under the main loop of tile i+1,
//   * dedicated epilogue warps (the low-plane writers never leave the ring).
// The one-tile-per-CTA kernel above pays launch + barrier init + TMEM alloc + pipeline fill + drain + epilogue for
// every 128-row tile; when a tile is only a handful of k-blocks (the transposed convolutions after tap clipping:
// 7 of 20 taps for Bach10, 13 of 25 for DSD100) that skeleton IS the run time (profiles/r2_notes.md: 245 760
// tiles x ~10 us / 296 resident CTAs = 8.3 of the 9.4 ms of Bach10's InverseLayer(conv2)).
//   warps 0-3 epilogue | warp 4 MMA issue + TMEM alloc | warp 5 copy engine | warps 6-9 low plane
constexpr int TP_THREADS = 10 * 32;

template <int BN, int STAGES>
__global__ void __launch_bounds__(TP_THREADS, 1)
gemm_tma_persist_kernel(const GemmDesc d, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const TmaGemmArgs g, int tiles_x, int tiles_y) {
  using SM = TmSmem<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* split = full + STAGES;
  uint64_t* empty = split + STAGES;
  uint64_t* tmem_full = empty + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kdg = (g.n_kd + g.kdb - 1) / g.kdb;
  const int per_u = g.tiles_pos * n_kdg;
  const int64_t total_tiles = (int64_t)tiles_x * tiles_y;
  const uint32_t a_tx = (uint32_t)(g.pb * g.tb * g.kdb) * ROW_BYTES;
  const int n_main = g.acc_mode == 2 ? 3 : 1;
  const int corr_acc = g.acc_mode == 0 ? 0 : n_main;
  constexpr uint32_t SET_COLS = 4 * BN;

  // tile id -> (first output time / plane / position of the tile, first output column, k-block range)
  struct Tile { int u, kd, r0, n0, kb_lo, num_kb; };
  auto decode = [&](int64_t t) {
    Tile x;
    const int bx = (int)(t / tiles_y);
    x.n0 = (int)(t - (int64_t)bx * tiles_y) * BN;
    const int n_ug = (g.n_u + g.tb - 1) / g.tb;     // u fastest: see gemm_tma_kernel
    const int ug = bx % n_ug, rest = bx / n_ug;
    const int kdg = rest / g.tiles_pos;
    x.u = ug * g.tb;
    x.kd = kdg * g.kdb;
    x.r0 = (rest - kdg * g.tiles_pos) * g.pb;
    int kb_lo = 0, kb_hi = (d.K + KSTAGE - 1) / KSTAGE;
    if (d.kc_rows > 0) {
      const int q_lo = max(0, d.kc_pad - x.u), q_hi = min(d.kc_taps - 1, d.kc_pad + d.kc_n - 1 - x.u);
      kb_lo = (d.kc_unit * q_lo) / KSTAGE;
      kb_hi = min(kb_hi, (d.kc_unit * (q_hi + 1) + KSTAGE - 1) / KSTAGE);
      if (kb_hi <= kb_lo) kb_hi = kb_lo + 1;
    }
    x.kb_lo = kb_lo;
    x.num_kb = kb_hi - kb_lo;
    return x;
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], TM_SPLIT_WARPS * 32);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 5 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 4) tmem_alloc(tmem_slot, 2 * SET_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    if (elect_one()) {
      // ---------------------------------------------------------------- copy engine
      int it = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const Tile x = decode(t);
        for (int kb = 0; kb < x.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * SM::STAGE_BYTES;
          const int k0 = (x.kb_lo + kb) * KSTAGE;
          mbar_arrive_expect_tx(&full[s], a_tx + 2 * SM::B_BYTES);
          if (g.conv) {
            const int q = (x.kb_lo + kb) / g.kw, w = (x.kb_lo + kb) - q * g.kw;
            tma_load_4d(st, &tmA, &full[s], 0, x.r0 + w, x.u + q, x.kd);
          } else {
            tma_load_2d(st, &tmA, &full[s], x.u * g.col_per_u + k0, x.r0);
          }
          tma_load_2d(st + 2 * SM::A_BYTES, &tmB, &full[s], k0, x.n0);
          tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &tmB, &full[s], k0, g.Np + x.n0);
        }
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_tf32(TM_BM, BN);
      constexpr uint32_t ST_D = SM::STAGE_BYTES >> 4, A_D = SM::A_BYTES >> 4, B_D = SM::B_BYTES >> 4;
      const uint32_t lo0 = desc_lo(smem_u32(smem));
      int it = 0, ti = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
        const Tile x = decode(t);
        const int set = ti & 1;
        const uint32_t acc = tmem_base + set * SET_COLS;
        mbar_wait(&tmem_empty[set], ((ti >> 1) & 1) ^ 1);
        fence_after_sync();
        int ma = 0, ks = 0;
        const uint32_t t_corr = acc + corr_acc * BN;
        for (int kb = 0; kb < x.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t par = (it / STAGES) & 1;
          mbar_wait(&full[s], par);
          fence_after_sync();
          const uint32_t a_hi = lo0 + s * ST_D, a_lo = a_hi + A_D, b_hi = a_hi + 2 * A_D, b_lo = b_hi + B_D;
#pragma unroll
          for (int j = 0; j < KSTAGE / 8; ++j, ++ks) {
            umma_tf32(t_corr, desc_of(a_hi + KSTEP_DESC * j), desc_of(b_lo + KSTEP_DESC * j), idesc, ks != 0);
            umma_tf32(acc + ma * BN, desc_of(a_hi + KSTEP_DESC * j), desc_of(b_hi + KSTEP_DESC * j), idesc,
                      corr_acc == 0 ? 1 : (ks >= n_main));
            ma = ma + 1 == n_main ? 0 : ma + 1;
          }
          mbar_wait(&split[s], par);
          fence_after_sync();
#pragma unroll
          for (int j = 0; j < KSTAGE / 8; ++j)
            umma_tf32(t_corr, desc_of(a_lo + KSTEP_DESC * j), desc_of(b_hi + KSTEP_DESC * j), idesc, 1);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[set]);
      }
    }
  } else if (warp > 5) {
    // ------------------------------------------------------------------ low-plane writers
    const int pt = tid - 6 * 32;
    int it = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile x = decode(t);
      for (int kb = 0; kb < x.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        float4* raw = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES + SM::A_BYTES);
        constexpr int CHUNKS = SM::A_BYTES / 16 / (TM_SPLIT_WARPS * 32);   // 8 per thread
        float4 v[CHUNKS];
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) v[i] = raw[i * (TM_SPLIT_WARPS * 32) + pt];
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
          float4 h, l;
          split4(v[i], h, l);
          lo[i * (TM_SPLIT_WARPS * 32) + pt] = l;
        }
        fence_proxy_async();
        mbar_arrive(&split[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (thread = tile row = TMEM lane)
    int ti = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const Tile x = decode(t);
      const int set = ti & 1;
      const int n_main_used = min(n_main, x.num_kb * (KSTAGE / 8));
      const int pos = x.r0 + tid % g.pb, t2 = tid / g.pb, tl = t2 % g.tb, kdl = t2 / g.tb;
      const int64_t m64 = (int64_t)(x.u + tl) * g.m_inner + (int64_t)(x.kd + kdl) * g.n_pos + pos;
      const bool m_ok = kdl < g.kdb && x.kd + kdl < g.n_kd && x.u + tl < g.n_u && pos < g.n_pos && m64 < d.M;
      const int m = m_ok ? (int)m64 : 0;
      const int64_t roff = gemm_c_row_offset(d, m);
      const bool st_ok = m_ok && roff >= 0;
      mbar_wait_relaxed(&tmem_full[set], (ti >> 1) & 1);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + set * SET_COLS;
      const int n0 = x.n0;
#pragma unroll 1
      for (int j = 0; j < BN / 16; ++j) {
        const bool live = n0 + 16 * j < d.N;  // warp-uniform
        float v[16];
        if (live) {
          // all accumulators of the chunk in flight, one wait (each tcgen05.ld + wait round trip is ~100+ cycles)
          float w1[16], w2[16], wc[16];
          tmem_ld16_nowait(taddr + 16 * j, v);
          if (n_main_used > 1) tmem_ld16_nowait(taddr + BN + 16 * j, w1);
          if (n_main_used > 2) tmem_ld16_nowait(taddr + 2 * BN + 16 * j, w2);
          if (corr_acc) tmem_ld16_nowait(taddr + corr_acc * BN + 16 * j, wc);
          tmem_wait_ld();
          if (n_main_used > 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w1[i];
          }
          if (n_main_used > 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w2[i];
          }
          if (corr_acc) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += wc[i];
          }
        }
        if (j == BN / 16 - 1) {   // every column of this set has been read: the MMA warp may start tile i+2 in it
          fence_before_sync();
          mbar_arrive(&tmem_empty[set]);
        }
        if (live && st_ok) gemm_store16(d, g.cvec != 0, d.n_seg < d.N, roff, n0 + 16 * j, v);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 2 * SET_COLS);
  }
}

// ---- persistent variant with the A operand in tensor memory ------------------------------------
// What bounded the kernels above is shared-memory bandwidth: per k-block the copy engine writes the tile, the low-plane
// warps read it and write the low plane back, and each of the twelve MMAs fetches 4 KB of A and BN/32 KB of B again.
// Here the four warps that already pull the raw tile through registers write BOTH planes to tensor memory instead
// (tcgen05.st, thread = row = TMEM lane) and the MMAs take A from there: shared memory carries the TMA fill, one
// read of the raw tile and the B fetches only (136 -> 72 KB per k-block at BN = 64).
constexpr int TA_THREADS = 14 * 32;   // warps 0-3 and 10-13 epilogue, 4 MMA, 5 copy engine, 6-9 A writers

template <int BN, int STAGES>
struct TmSmemAtm {
  static constexpr int A_BYTES = TM_BM * ROW_BYTES;  // 16 KB raw tile
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;   // A raw, B hi, B lo
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TA_THREADS, 1)
gemm_tma_atm_kernel(const GemmDesc d, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const TmaGemmArgs g, int tiles_x, int tiles_y) {
  using SM = TmSmemAtm<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);    // TMA landed (raw A tile + both B planes)
  uint64_t* empty = full + STAGES;                                      // MMAs that read the stage's B planes retired
  uint64_t* a_ready = empty + STAGES;        // [4] both A planes of a k-block are in tensor memory
  uint64_t* a_free = a_ready + 4;            // [4] the MMAs that read them retired
  uint64_t* tmem_full = a_free + 4;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kdg = (g.n_kd + g.kdb - 1) / g.kdb;
  const int per_u = g.tiles_pos * n_kdg;
  const int64_t total_tiles = (int64_t)tiles_x * tiles_y;
  const uint32_t a_tx = (uint32_t)(g.pb * g.tb * g.kdb) * ROW_BYTES;
  // 512 TMEM columns: two accumulator sets + a two-deep ring of A operand planes (hi, lo: 2 x 32 columns each).
  // BN = 64: 2 x 3 accumulators (2 main + 1 correction); BN = 32: 2 x 4 (3 main + 1 correction)
  constexpr int NACC = BN == 64 ? 3 : 4;
  const int n_main = g.acc_mode == 2 ? NACC - 1 : 1;
  const int corr_acc = g.acc_mode == 0 ? 0 : n_main;
  constexpr uint32_t SET_COLS = NACC * BN;
  constexpr uint32_t A_COL0 = 2 * SET_COLS;
  constexpr int ASLOTS = (512 - (int)A_COL0) / 64;      // A ring depth in tensor memory: 2 (BN 64) or 4 (BN 32)

  // tile id -> (first output time / plane / position of the tile, first output column, k-block range)
  struct Tile { int u, kd, r0, n0, kb_lo, num_kb; };
  auto decode = [&](int64_t t) {
    Tile x;
    const int bx = (int)(t / tiles_y);
    x.n0 = (int)(t - (int64_t)bx * tiles_y) * BN;
    const int n_ug = (g.n_u + g.tb - 1) / g.tb;     // u fastest: see gemm_tma_kernel
    const int ug = bx % n_ug, rest = bx / n_ug;
    const int kdg = rest / g.tiles_pos;
    x.u = ug * g.tb;
    x.kd = kdg * g.kdb;
    x.r0 = (rest - kdg * g.tiles_pos) * g.pb;
    int kb_lo = 0, kb_hi = (d.K + KSTAGE - 1) / KSTAGE;
    if (d.kc_rows > 0) {
      const int q_lo = max(0, d.kc_pad - x.u), q_hi = min(d.kc_taps - 1, d.kc_pad + d.kc_n - 1 - x.u);
      kb_lo = (d.kc_unit * q_lo) / KSTAGE;
      kb_hi = min(kb_hi, (d.kc_unit * (q_hi + 1) + KSTAGE - 1) / KSTAGE);
      if (kb_hi <= kb_lo) kb_hi = kb_lo + 1;
    }
    x.kb_lo = kb_lo;
    x.num_kb = kb_hi - kb_lo;
    return x;
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&a_ready[s], TM_SPLIT_WARPS * 32);
      mbar_init(&a_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 256);
    }
    fence_barrier_init();
  }
  if (warp == 5 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    if (elect_one()) {
      // ---------------------------------------------------------------- copy engine
      int it = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const Tile x = decode(t);
        for (int kb = 0; kb < x.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * SM::STAGE_BYTES;
          const int k0 = (x.kb_lo + kb) * KSTAGE;
          mbar_arrive_expect_tx(&full[s], a_tx + 2 * SM::B_BYTES);
          if (g.conv) {
            const int q = (x.kb_lo + kb) / g.kw, w = (x.kb_lo + kb) - q * g.kw;
            tma_load_4d(st, &tmA, &full[s], 0, x.r0 + w, x.u + q, x.kd);
          } else {
            tma_load_2d(st, &tmA, &full[s], x.u * g.col_per_u + k0, x.r0);
          }
          tma_load_2d(st + SM::A_BYTES, &tmB, &full[s], k0, x.n0);
          tma_load_2d(st + SM::A_BYTES + SM::B_BYTES, &tmB, &full[s], k0, g.Np + x.n0);
        }
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_tf32(TM_BM, BN);
      constexpr uint32_t ST_D = SM::STAGE_BYTES >> 4, A_D = SM::A_BYTES >> 4, B_D = SM::B_BYTES >> 4;
      const uint32_t lo0 = desc_lo(smem_u32(smem));
      int it = 0, ti = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
        const Tile x = decode(t);
        const int set = ti & 1;
        const uint32_t acc = tmem_base + set * SET_COLS;
        mbar_wait(&tmem_empty[set], ((ti >> 1) & 1) ^ 1);
        fence_after_sync();
        int ma = 0, ks = 0;
        const uint32_t t_corr = acc + corr_acc * BN;
        for (int kb = 0; kb < x.num_kb; ++kb, ++it) {
          const int s = it % STAGES, as = it % ASLOTS;
          mbar_wait(&full[s], (it / STAGES) & 1);
          mbar_wait(&a_ready[as], (it / ASLOTS) & 1);
          fence_after_sync();
          const uint32_t b_hi = lo0 + s * ST_D + A_D, b_lo = b_hi + B_D;
          const uint32_t ta_hi = tmem_base + A_COL0 + as * 64, ta_lo = ta_hi + 32;
          // corrections into their own accumulator, the main product rotating over the others
#pragma unroll
          for (int j = 0; j < KSTAGE / 8; ++j, ++ks) {
            umma_tf32_ts(t_corr, ta_hi + 8 * j, desc_of(b_lo + KSTEP_DESC * j), idesc, ks != 0);
            umma_tf32_ts(t_corr, ta_lo + 8 * j, desc_of(b_hi + KSTEP_DESC * j), idesc, 1);
            umma_tf32_ts(acc + ma * BN, ta_hi + 8 * j, desc_of(b_hi + KSTEP_DESC * j), idesc, corr_acc == 0 ? 1 : (ks >= n_main));
            ma = ma + 1 == n_main ? 0 : ma + 1;
          }
          umma_commit(&empty[s]);
          umma_commit(&a_free[as]);
        }
        umma_commit(&tmem_full[set]);
      }
    }
  } else if (warp > 5 && warp < 10) {
    // ------------------------------------------------------------------ A operand -> tensor memory
    // thread = tile row = TMEM lane (a warp may only touch the lane quadrant warp % 4): reads its 128-byte row of the
    // raw tile (eight swizzled 16-byte chunks), writes hi = trunc_tf32(x) and lo = x - hi as 2 x 32 columns
    // (two groups of four warps taking alternate k-blocks were tried: slower, 0.124 vs 0.116 ms on DSD100's convT2)
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0;
    int it = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile x = decode(t);
      for (int kb = 0; kb < x.num_kb; ++kb, ++it) {
        const int s = it % STAGES, as = it % ASLOTS;
        mbar_wait(&full[s], (it / STAGES) & 1);
        const uint8_t* raw = smem + s * SM::STAGE_BYTES;
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(raw + tile_off(row, c));
          float4 h, l;
          split4(v, h, l);
          hi[4 * c] = h.x; hi[4 * c + 1] = h.y; hi[4 * c + 2] = h.z; hi[4 * c + 3] = h.w;
          lo[4 * c] = l.x; lo[4 * c + 1] = l.y; lo[4 * c + 2] = l.z; lo[4 * c + 3] = l.w;
        }
        mbar_wait(&a_free[as], ((it / ASLOTS) & 1) ^ 1);
        fence_after_sync();
        tmem_st32(lane_base + as * 64, hi);
        tmem_st32(lane_base + as * 64 + 32, lo);
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&a_ready[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (thread = tile row = TMEM lane)
    // eight warps: warps 0-3 and 10-13 pair up on the four lane quadrants, each taking every other 16-column chunk --
    // with tiles of ~7 k-blocks the epilogue of a tile, not its main loop, set the pace
    const int half = warp >= 10 ? 1 : 0;
    const int erow = (warp & 3) * 32 + lane;
    int ti = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const Tile x = decode(t);
      const int set = ti & 1;
      const int n_main_used = min(n_main, x.num_kb * (KSTAGE / 8));
      const int pos = x.r0 + erow % g.pb, t2 = erow / g.pb, tl = t2 % g.tb, kdl = t2 / g.tb;
      const int64_t m64 = (int64_t)(x.u + tl) * g.m_inner + (int64_t)(x.kd + kdl) * g.n_pos + pos;
      const bool m_ok = kdl < g.kdb && x.kd + kdl < g.n_kd && x.u + tl < g.n_u && pos < g.n_pos && m64 < d.M;
      const int m = m_ok ? (int)m64 : 0;
      const int64_t roff = gemm_c_row_offset(d, m);
      const bool st_ok = m_ok && roff >= 0;
      mbar_wait_relaxed(&tmem_full[set], (ti >> 1) & 1);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + set * SET_COLS;
      const int n0 = x.n0;
#pragma unroll 1
      for (int j = 0; j < BN / 16; ++j) {
        const bool live = n0 + 16 * j < d.N && (j & 1) == half;  // warp-uniform
        float v[16];
        if (live) {
          // all accumulators of the chunk in flight, one wait (each tcgen05.ld + wait round trip is ~100+ cycles)
          float w1[16], w2[16], wc[16];
          tmem_ld16_nowait(taddr + 16 * j, v);
          if (n_main_used > 1) tmem_ld16_nowait(taddr + BN + 16 * j, w1);
          if (n_main_used > 2) tmem_ld16_nowait(taddr + 2 * BN + 16 * j, w2);
          if (corr_acc) tmem_ld16_nowait(taddr + corr_acc * BN + 16 * j, wc);
          tmem_wait_ld();
          if (n_main_used > 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w1[i];
          }
          if (n_main_used > 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w2[i];
          }
          if (corr_acc) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += wc[i];
          }
        }
        if (j == BN / 16 - 1) {   // every column of this set has been read: the MMA warp may start tile i+2 in it
          fence_before_sync();
          mbar_arrive(&tmem_empty[set]);
        }
        if (live && st_ok) gemm_store16(d, g.cvec != 0, d.n_seg < d.N, roff, n0 + 16 * j, v);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  // the driver entry point is fetched through the runtime: libdcs.so links cudart statically and
  // does not link libcuda
  static const EncodeTiledFn fn = [] {   // C++11 magic static: initialised once, thread-safe
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return (EncodeTiledFn)p;
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

// fp32 tensor of `rank` dimensions (dims[0] innermost, strides in BYTES for dims 1..), box = 32
// floats x box_rows along dim 1 x box2 x box3, 128-byte swizzle, zero fill outside
static int encode_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      uint32_t box_rows, bool quiet, uint32_t box2 = 1, uint32_t box3 = 1) {
  EncodeTiledFn fn = encode_tiled_fn();
  DCS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gd[4], gs[3];
  cuuint32_t box[4] = {(cuuint32_t)KSTAGE, box_rows, box2, box3}, estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) gd[i] = dims[i];
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gd, gs, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && quiet) return DCS_ECUDA;
  DCS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu pitch %llu box %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0], box_rows);
  return DCS_OK;
}

int tma_encode_2d_f32(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows}, strides[1] = {pitch_bytes};
  return encode_map(map, base, 2, dims, strides, box_rows, false);
}

int tma_encode_3d_f32(CUtensorMap* map, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box_rows) {
  const uint64_t dims[3] = {d0, d1, d2}, strides[2] = {stride1_bytes, stride2_bytes};
  return encode_map(map, base, 3, dims, strides, box_rows, false);
}

// Overlapping rows (pitch < K: a convolution over time read in place) make a tensor whose row
// pitch is smaller than its row extent.  The copy engine only does address arithmetic and a
// per-dimension bounds check, so it works; should a driver refuse to encode such a map, the GEMM
// stays on the register-staged kernel (remembered here).
static std::atomic<bool> g_overlap_rejected{false};   // a property of the driver: process-wide is right

// how the copy engine can address the A view of `d`
struct TmaView {
  int mode = 0;                 // 0: not at all, 1: rows (2-D), 2: conv (4-D)
  int kind = 0;                 // bring-up mask bit (DCS_DEBUG_TMA_MASK)
  int rows_per_u = 0, n_u = 1, col_per_u = 0, n_kd = 1, kw = 1;
  int pb = TM_BM, tb = 1, kdb = 1;   // A box (TmaGemmArgs)
  int rank = 2;
  uint64_t dims[4] = {1, 1, 1, 1}, strides[3] = {0, 0, 0};
  bool overlap = false;
};

static TmaView classify(const GemmDesc& d) {
  TmaView v;
  if ((uintptr_t)d.A % 16 != 0 || d.M <= 0) return v;
  if (d.win_stride > 0) {                                 // ---- window mode (strided conv1 of the 30-channel nets)
    // 4-D map {32 window floats (stride 4 B), positions (stride win_stride), channel plane, frame}; one k-block = one
    // channel plane: in the conv-mode kernel the "time tap" coordinate walks the planes (q = k-block, u = 0) and the
    // "plane" coordinate is the frame.  Frames >= a_valid_rows / m_inner are the copy engine's zero fill.
    if (d.win_stride % 4 != 0 || d.k_seg != KSTAGE || d.K % KSTAGE != 0 || d.m_inner <= 0 || d.a_so % 4 != 0 || d.k_ss % 4 != 0 ||
        d.kc_rows != 0 || d.a_valid_rows % d.m_inner != 0)
      return v;
    const int nseg = d.K / KSTAGE;
    v.rank = 4;
    v.dims[0] = KSTAGE; v.dims[1] = (uint64_t)d.m_inner; v.dims[2] = (uint64_t)nseg; v.dims[3] = (uint64_t)(d.a_valid_rows / d.m_inner);
    v.strides[0] = (uint64_t)d.win_stride * 4;
    v.strides[1] = (uint64_t)(nseg > 1 ? d.k_ss : d.a_so) * 4;
    v.strides[2] = (uint64_t)d.a_so * 4;
    v.n_u = 1; v.kw = 1;
    v.rows_per_u = d.m_inner;
    v.n_kd = (int)ceil_div64(d.M, d.m_inner);
    if (v.dims[3] == 0) return v;
    v.overlap = true;        // windows overlap (pitch 16 B < 128 B): a driver that refused such a map would mean fallback
    v.kind = 32;
    v.mode = 2;
    return v;
  }
  if (d.k_seg >= d.K) {                                   // ---- rows mode
    if (d.m_inner2 != 1) return v;
    const bool plain = d.m_inner == 1;
    const int64_t pitch = plain ? d.a_so : d.a_si;
    if (pitch <= 0 || pitch % 4 != 0) return v;
    if (plain) {
      v.overlap = pitch < d.K;
      if (v.overlap && g_overlap_rejected) return v;
      v.kind = v.overlap ? 8 : (d.a_valid_rows < d.M ? 2 : 1);
      v.rows_per_u = d.M;
      v.dims[0] = (uint64_t)d.K; v.dims[1] = (uint64_t)std::min(d.M, d.a_valid_rows);
    } else {
      if (d.a_valid_rows < d.M || (d.kc_rows != 0 && d.kc_rows != d.m_inner)) return v;
      v.n_u = (int)ceil_div64(d.M, d.m_inner);
      if ((int64_t)(v.n_u - 1) * d.a_so + d.K > pitch) return v;       // every row stays inside its pitch
      v.kind = 4;
      v.rows_per_u = d.m_inner; v.col_per_u = (int)d.a_so;
      v.dims[0] = (uint64_t)pitch; v.dims[1] = (uint64_t)std::min<int64_t>(d.m_inner, d.M);
    }
    v.strides[0] = (uint64_t)pitch * 4;
    v.mode = 1;
    return v;
  }
  // ---- conv mode: K = taps x 32 channels, one tap per k-block, time tap stride = time step
  const int64_t pos_stride = d.m_inner2 > 1 ? d.a_s2 : d.a_si;
  if (d.k_seg % KSTAGE != 0 || d.K % d.k_seg != 0 || d.k_ss != d.a_so || pos_stride != KSTAGE || d.a_so % KSTAGE != 0) return v;
  if (d.a_valid_rows < d.M || d.m_inner <= 1 || (d.kc_rows != 0 && (d.kc_rows != d.m_inner || d.kc_unit != d.k_seg))) return v;
  v.n_u = (int)ceil_div64(d.M, d.m_inner);
  const int taps_t = d.K / d.k_seg;
  v.kw = d.k_seg / KSTAGE;
  v.rank = 4;
  v.dims[0] = KSTAGE;
  v.dims[1] = (uint64_t)(d.a_so / KSTAGE);                 // positions per time row (padding included)
  v.strides[0] = KSTAGE * 4;
  v.strides[1] = (uint64_t)d.a_so * 4;
  if (d.m_inner2 > 1) {                                    // rows (u, kd, pos): transposed conv on the padded planes
    if (d.m_inner % d.m_inner2 != 0 || d.a_si % d.a_so != 0) return v;
    v.rows_per_u = d.m_inner2; v.n_kd = d.m_inner / d.m_inner2;
    v.dims[2] = (uint64_t)(d.a_si / d.a_so);               // time rows per plane
    v.dims[3] = (uint64_t)v.n_kd;
    v.strides[2] = (uint64_t)d.a_si * 4;
  } else {                                                 // rows (u, pos): forward conv on one long plane
    v.rows_per_u = d.m_inner; v.n_kd = 1;
    v.dims[2] = (uint64_t)(v.n_u + taps_t - 1);
    v.dims[3] = 1;
    v.strides[2] = v.strides[1] * v.dims[2];
  }
  if ((uint64_t)(v.rows_per_u + v.kw - 1) > v.dims[1]) return v;   // taps must stay inside the time row
  if (v.rows_per_u < TM_BM) {   // narrow layer: fill the 128 tile rows with several planes / output times
    v.pb = v.rows_per_u;
    const int group = TM_BM / v.pb;
    if (d.m_inner2 > 1) v.kdb = std::max(1, std::min(group, v.n_kd));
    else if (d.kc_rows == 0) v.tb = std::max(1, std::min(group, v.n_u));
  }
  v.kind = 16;
  v.mode = 2;
  return v;
}

bool gemm_tma_eligible(const GemmDesc& d, int mask) {
  const TmaView v = classify(d);
  return v.mode != 0 && (mask & v.kind);
}

// tensor maps of the stacked [hi; lo] weight planes (each Np x Kp) for the three tile widths, made
// when the weight is created -- single-threaded -- so that contexts on several host threads can
// share a model without racing on the lazily filled cache.  A failure here is not an error: the
// launch path encodes on first use (and reports).
void tc_weight_encode_maps(TcWeight* w) {
  const uint64_t dims[2] = {(uint64_t)w->Kp, (uint64_t)2 * w->Np}, strides[1] = {(uint64_t)w->Kp * 4};
  const uint32_t rows[3] = {32, 64, 128};
  for (int slot = 0; slot < 3; ++slot)
    w->tmap_ok[slot] = encode_map(&w->tmap[slot], w->hi, 2, dims, strides, rows[slot], true) == DCS_OK;
}

template <int BN, int STAGES>
static int launch_tma(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st) {
  using SM = TmSmem<BN, STAGES>;
  DCS_TRY(ensure_smem_attr(gemm_tma_kernel<BN, STAGES>, SM::TOTAL));
  constexpr int slot = BN == 32 ? 0 : (BN == 64 ? 1 : 2);
  if (!w.tmap_ok[slot]) {   // normally done by tc_weight_encode_maps at weight creation
    const uint64_t dims[2] = {(uint64_t)w.Kp, (uint64_t)2 * w.Np}, strides[1] = {(uint64_t)w.Kp * 4};
    DCS_TRY(encode_map(&w.tmap[slot], w.hi, 2, dims, strides, BN, false));
    w.tmap_ok[slot] = true;
  }
  const TmaView v = classify(d);
  DCS_REQUIRE(v.mode != 0, "tma gemm: operand view not supported");
  TmaGemmArgs g;
  g.conv = v.mode == 2;
  g.pb = v.pb; g.tb = v.tb; g.kdb = v.kdb;
  g.n_pos = v.rows_per_u;
  g.n_u = v.n_u; g.n_kd = v.n_kd;
  g.tiles_pos = (int)ceil_div64(v.rows_per_u, v.pb);
  g.col_per_u = v.col_per_u;
  g.kw = v.kw;
  g.m_inner = d.m_inner == 1 ? d.M : d.m_inner;
  g.Np = w.Np;
  g.acc_mode = ctx->tc_acc_mode;
  g.rewrite_hi = ctx->tma_mode == 2;
  g.prefetch = ctx->tma_prefetch;
  g.probe = ctx->tma_probe;
  g.cvec = ((uintptr_t)d.C % 16 == 0) && d.c_so % 4 == 0 && d.c_si % 4 == 0 && d.c_s2 % 4 == 0 && d.c_col0 % 4 == 0 &&
           (d.n_seg >= d.N || (d.n_seg % 4 == 0 && d.n_ss % 4 == 0));
  alignas(64) CUtensorMap tmA;
  if (v.overlap) {
    if (encode_map(&tmA, d.A, v.rank, v.dims, v.strides, (uint32_t)v.pb, true, (uint32_t)v.tb, (uint32_t)v.kdb) != DCS_OK) {
      g_overlap_rejected = true;
      return DCS_TMA_FALLBACK;
    }
  } else {
    DCS_TRY(encode_map(&tmA, d.A, v.rank, v.dims, v.strides, (uint32_t)v.pb, false, (uint32_t)v.tb, (uint32_t)v.kdb));
  }
  const int64_t gx = ceil_div64(v.n_u, v.tb) * ceil_div64(v.n_kd, v.kdb) * g.tiles_pos;
  DCS_REQUIRE(gx <= 0x7fffffff, "tma gemm: M=%d too large", d.M);
  const int64_t tiles = gx * ceil_div64(d.N, BN);
  // fewer tiles than CTA slots (two per SM) and a long K: split K so that every SM has work
  // (conv1 / conv2 of the DSD100 net: 122 tiles of 25-33 k-blocks); fixed-order reduction
  int splits = 1;
  const int num_kb = (d.K + KSTAGE - 1) / KSTAGE;
  if (d.kc_rows == 0 && tiles < 2 * ctx->num_sms && num_kb >= 8) {
    splits = (int)std::min<int64_t>(2 * ctx->num_sms / tiles, num_kb / 4);
    if (splits > 1) {
      const int per = (num_kb + splits - 1) / splits;
      splits = (num_kb + per - 1) / per;
    }
    if (splits < 2) splits = 1;
  }
  g.k_splits = splits;
  g.ldp = (d.N + 3) / 4 * 4;
  g.partial = nullptr;
  if (splits > 1) {
    DCS_TRY(ctx->net[8].ensure((size_t)splits * d.M * g.ldp * sizeof(float), st));
    g.partial = ctx->net[8].as<float>();
  }
  // many short tiles: one persistent CTA per SM with a continuous stage ring and double-buffered accumulators
  // (DCS_DEBUG_TMA_PERSIST=0: always the one-tile-per-CTA kernel)
  bool launched = false;
  if constexpr (BN <= 64) {
    // measured (profiles/r2_notes.md): the persistent kernel wins where a tile is a handful of k-blocks and there are
    // many of them -- the tap-clipped transposed convolutions (DSD100 0.153 -> 0.139 ms, Bach10 9.7 -> 7.4 ms); it
    // loses on long-K tiles (iKala's 200-tap transposed conv: two interleaved CTAs per SM hide more) and on the
    // store-heavy decoder dense layer (eight epilogue warps per SM beat four)
    const bool short_tiles = num_kb <= 32 && (d.N <= 64 || ctx->tma_persist_wide);   // (num_kb: the un-clipped count)
    if (ctx->tma_persist && splits == 1 && short_tiles && tiles >= (int64_t)ctx->tma_persist * ctx->num_sms) {
      const int tiles_y = (int)ceil_div64(d.N, BN);
      const unsigned ctas = (unsigned)std::min<int64_t>(tiles, ctx->num_sms);
      if (ctx->tma_atm) {     // A operand in tensor memory (DCS_DEBUG_TMA_ATM=0: both planes through shared memory)
        constexpr int AST = 6;                   // 6 x 32 KB (BN 64) / 6 x 24 KB (BN 32) of stages
        using ASM_ = TmSmemAtm<BN, AST>;
        DCS_TRY(ensure_smem_attr(gemm_tma_atm_kernel<BN, AST>, ASM_::TOTAL));
        gemm_tma_atm_kernel<BN, AST><<<ctas, TA_THREADS, ASM_::TOTAL, st>>>(d, tmA, w.tmap[slot], g, (int)gx, tiles_y);
      } else {
        constexpr int PST = BN == 32 ? 5 : 4;      // 5 x 40 KB / 4 x 48 KB of stages: one CTA per SM
        using PSM = TmSmem<BN, PST>;
        DCS_TRY(ensure_smem_attr(gemm_tma_persist_kernel<BN, PST>, PSM::TOTAL));
        gemm_tma_persist_kernel<BN, PST><<<ctas, TP_THREADS, PSM::TOTAL, st>>>(d, tmA, w.tmap[slot], g, (int)gx, tiles_y);
      }
      launched = true;
    }
  }
  if (!launched) {
    dim3 grid((unsigned)gx, (unsigned)ceil_div64(d.N, BN), (unsigned)splits);
    gemm_tma_kernel<BN, STAGES><<<grid, TM_THREADS, SM::TOTAL, st>>>(d, tmA, w.tmap[slot], g);
  }
  DCS_CHECK_LAUNCH();
  ctx->launches++;
  if (splits > 1) DCS_TRY(launch_splitk_reduce(ctx, d, g.partial, g.ldp, splits, st));
  if (ctx->tma_sync) {   // bring-up: attribute an asynchronous fault to the launch that caused it
    const cudaError_t e = cudaStreamSynchronize(st);
    DCS_REQUIRE(e == cudaSuccess, "tma gemm M=%d N=%d K=%d m_inner=%d a_so=%lld a_si=%lld mode=%d BN=%d: %s", d.M, d.N, d.K,
                d.m_inner, (long long)d.a_so, (long long)d.a_si, v.mode, BN, cudaGetErrorString(e));
  }
  return DCS_OK;
}

int launch_gemm_tma(dcs_ctx* ctx, const GemmDesc& d, const TcWeight& w, cudaStream_t st) {
  if (d.M <= 0 || d.N <= 0) return DCS_OK;
  DCS_REQUIRE(d.K == w.K && d.N == w.N, "tma gemm: weight is %dx%d, GEMM wants K=%d N=%d", w.K, w.N, d.K, d.N);
  if (d.N > 64 && ctx->tma_wide) return launch_tma<128, 3>(ctx, d, w, st);
  if (d.N <= 32) return ctx->tma_stages == 4 ? launch_tma<32, 4>(ctx, d, w, st) : launch_tma<32, 2>(ctx, d, w, st);
  // two stages = 97 KB and 256 TMEM columns: two CTAs per SM, one's epilogue under the other's main loop
  return ctx->tma_stages == 4 ? launch_tma<64, 4>(ctx, d, w, st) : launch_tma<64, 2>(ctx, d, w, st);
}

}  // namespace dcs

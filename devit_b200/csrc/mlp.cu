// Fused MLP block of one transformer layer (include/devit_b200.h, devit_mlp_fused):
//     x += gelu( LN(x) W1^T + b1 ) W2^T + b2            models/de_vit.py:35-47 + :115
// as ONE kernel: the hidden activation never leaves the SM pair.
//
// A CTA pair (cta_group::2) owns 256 token rows.  Per pair-tile:
//   * TMA loads the bf16 copy of the residual stream for its rows once (Y, 128 x 384 per CTA,
//     K-major, 128B swizzle) and streams W1 / W2 in 64-neuron chunks through two 2-slot rings;
//     each CTA stages HALF of every weight chunk (the pair's MMA reads both halves).
//   * GEMM1 chunk c: acc1[c&1] (128 x 64 fp32 per CTA, TMEM columns 384 + 64 (c&1)) = Y W1_c^T.
//   * 16 epilogue warps turn acc1 into H_c = gelu(rstd (acc - mean c1) + c2)  (LayerNorm folded
//     into the epilogue, see devit_gemm_args.ln_stats) and write it back INTO TMEM as packed
//     bf16 pairs over columns they have already consumed.
//   * GEMM2 chunk c: acc2 (128 x 384 fp32, TMEM columns [0, 384)) += H_c W2_c^T with the A operand
//     read from tensor memory (two N = D/2 UMMAs per K = 16 step).  GEMM1 of chunk c+1 is issued
//     before GEMM2 of chunk c, so the tensor pipe works while the GELU of chunk c runs.
//   * Final epilogue: acc2 + b2 + residual -> x (fp32).  At the end of a tile Y and both weight
//     rings are dead: their 192 KB (D = 384) become 48 slots of [32 rows x 32 fp32], the WHOLE
//     residual tile, three slots per epilogue warp, TMA-prefetched as soon as the last GEMM1
//     (Y, W1 ring) / last GEMM2 (W2 ring) retired.  Each warp adds, stores through TMA and
//     also emits the bf16 copy and the partial row sums the next layer's LayerNorm-folded QKV
//     GEMM consumes.  The loaders resume only when every slot has been drained (all_free).
// TMEM: D + 2 x 64 columns (all 512 at D = 384).  smem at D = 384: Y 96 KB + W1 ring 48 KB +
// W2 ring 48 KB + 32 KB bf16 staging.  Instantiated for D = 384 (dedeit) and D = 256 (cct_7).
//
// PROJ variant (devit_mlp_args.o != NULL): the attention-output projection and its residual add
// (models/de_vit.py:81-82, :114) run in the SAME kernel in front of the MLP:
//     x1 = x + o Wp^T + bp ;  x = x1 + gelu( LN(x1) W1^T + b1 ) W2^T + b2
// so the residual stream makes ONE round trip through HBM per layer instead of two and neither
// x1, its bf16 copy nor its row statistics ever exist in global memory.  Between fused layers the
// stream travels as two bf16 planes, x = hi + lo (x_lo_in / x_lo_out: an SM stores ~32 B/clk, so
// 4 bytes per element instead of fp32 + bf16 copy = 6 shorten the write-bound final epilogue; hi
// is the next QKV GEMM's operand); the first layer reads and the last one writes fp32.
// Per pair-tile:
//   * the attention output tile O (128 x 64h bf16 per CTA) is TMA-loaded into the Y buffer and
//     Wp streams through the W2 ring, one 64-wide head chunk at a time;
//   * preload: every epilogue warp brings the residual (fp32, or hi + lo planes) of its 32 rows x
//     D/4 columns into its two private 4 KB slots by TMA (requested while the PREVIOUS tile's stores drain; the slots
//     lie over the weight rings and the staging buffer, never over Y, so the next tile's O is
//     loaded as soon as the last GEMM1 has retired), adds bp
//     and writes x + bp INTO acc2 (tcgen05.st), so the HBM-latency-bound part of the residual add
//     happens before the tensor core needs the tile and off the GEMM0 -> GEMM1 path;
//   * GEMM0: acc2 += O Wp^T (A and B from shared memory, same N = D/2 UMMA pairs as GEMM2),
//     giving x1 in acc2;
//   * epilogue 0 (all sixteen epilogue warps, D/4 columns each, 32 at a time): read x1 back,
//     accumulate the row's (sum, sum^2) and write bf16(x1) into the Y buffer in the K-major
//     128B-swizzled layout GEMM1 reads.  The four partial row sums of a row meet in shared memory
//     (one named barrier), giving the exact LayerNorm statistics of x1;
//   * the hidden-chunk loop is unchanged (GEMM2 accumulates on top of x1); the final epilogue
//     adds b2, has no residual to fetch, and stores hi / lo planes (or fp32 x + a bf16 copy written
//     straight from registers) through the same two slots.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace devit {

// warps: 0 = Y + W1 loads, 1 = MMA issue, 2..17 = epilogue (four per TMEM lane quarter: every
// hidden chunk is split 4 x 16 columns, because GEMM1 of chunk c+1 has to wait for the GELU of
// chunk c-1 -- two accumulator buffers -- so the GELU LATENCY of a chunk is what bounds the loop;
// all sixteen also run the final epilogue, D/4 output columns each), 18 = W2 loads
constexpr int kMlpThreads = 19 * 32;
constexpr int kW2Warp = 18;

// Geometry for a model width D (384: dedeit, 256: cct_7): everything below is per CTA.
template <int D>
struct MlpCfg {
  static_assert(D == 256 || D == 384, "fused MLP is laid out for D = 256 or 384");
  static constexpr int kAtoms = D / 64;             // K-atoms of Y / of a W1 chunk
  static constexpr int kYBytes = kAtoms * 16384;    // [128 rows x 128 B] per atom
  static constexpr int kW1Slot = kAtoms * 4096;     // [32 rows x 128 B] per atom (half chunk)
  static constexpr int kHalfN = D / 2;              // GEMM2 is issued as two N = D/2 UMMAs
  static constexpr int kW2Rows = D / 4;             // rows of one N-half staged by this CTA
  static constexpr int kW2Slot = 2 * kW2Rows * 128;
  static constexpr int kOffW1 = kYBytes;
  static constexpr int kOffW2 = kOffW1 + 2 * kW1Slot;
  static constexpr int kOffXb = kOffW2 + 2 * kW2Slot;  // bf16-copy staging: 16 x [32 rows x 64 B]
  // PROJ: the 4 KB row-statistics exchange of epilogue 0 lives at the bottom of the staging
  // buffer (idle until the final epilogue)
  static constexpr int kOffStat = kOffXb;
  static constexpr int kXbBytes = 16 * 2048;
  // PROJ: two 4 KB fp32 slots per epilogue warp (store staging of the final epilogue, then the
  // next tile's residual) laid over the weight rings and the staging buffer -- NOT over Y, so
  // the next tile's attention output can be loaded as soon as the last GEMM1 has retired
  static constexpr int kOffPSlots = kOffW1;
  static constexpr int kPSlotBytes = 32 * 4096;
  static constexpr int kOffBar = (kOffXb + kXbBytes) > (kOffPSlots + kPSlotBytes)
                                     ? (kOffXb + kXbBytes) : (kOffPSlots + kPSlotBytes);
  static constexpr int kSmem = kOffBar + 1024 + 1024;
  static constexpr int kAcc1Col = D;                // acc2 = TMEM columns [0, D), acc1 behind it
  static constexpr int kColsPerWarp = D / 4;        // final epilogue: output columns per warp
  static constexpr int kSlots = kColsPerWarp / 32;  // 32-column residual slots per warp
  // Y + both weight rings are exactly the 16 * kSlots slots of the fp32 residual tile
  static_assert(kYBytes + 2 * kW1Slot + 2 * kW2Slot == 16 * kSlots * 4096, "slot carve-out");
  static_assert(kSmem <= 227 * 1024, "fused MLP shared memory budget");
  static_assert(D + 128 <= 512, "TMEM budget");
};

extern long long* g_attn_trace;  // devit_debug_set_trace buffer (shared)
#ifdef DEVIT_GEMM_TRACE
#define MLP_TRACE(slot_, idx_)                                                          \
  do {                                                                                  \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (idx_) < 512)                        \
      p.trace[(slot_) * 512 + (idx_)] = clock64();                                      \
  } while (0)
#else
#define MLP_TRACE(slot_, idx_) do { } while (0)
#endif

struct MlpParams {
  long long* trace;
  int stagger;  // clocks by which some clusters delay their first tile (see the kernel)
  int stagger_mode;  // 0: every other cluster; 1: only the clusters with a tile less than the rest
  int M, F_ld, num_chunks;
  const float* c1;
  const float* c2;
  const float* b2;
  const float* ln_stats;
  int ln_parts;
  float ln_inv_dim, ln_eps;
  __nv_bfloat16* xb_out;
  float* stats_out;
  const float* bp;   // PROJ: attention-output projection bias [D]
  int proj_chunks;   // PROJ: kept heads (64-wide K chunks of O / Wp)
  // PROJ: the residual stream as two bf16 planes, x = hi + lo (hi doubles as the next QKV GEMM's
  // operand): an SM stores at most ~32 B/clk, and fp32 x + its bf16 copy are 6 bytes per element
  // against 4 for hi + lo (profiles/r2_ubench_store_bw.txt)
  int split_in;      // residual read from the hi / lo planes (tmXHi / tmXLoIn) instead of fp32 x
  __nv_bfloat16* lo_out;  // != NULL: write hi (xb_out) + lo planes instead of fp32 x + bf16 copy
};

__device__ __forceinline__ void umma_bf16_ts_cg2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32-byte global store (STG.256, sm_100): one full sector per lane and instruction
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2,
                                              uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6,
                                              uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

__device__ __forceinline__ int mlp_off(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }

// tmY: the bf16 copy of x (plain variant) or the attention output O (PROJ); tmWp: Wp (PROJ only)
template <int D, bool PROJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmXB, const __grid_constant__ CUtensorMap tmWp,
                 const __grid_constant__ CUtensorMap tmXHi, const __grid_constant__ CUtensorMap tmXLoIn,
                 const __grid_constant__ CUtensorMap tmXLoOut, const __grid_constant__ MlpParams p) {
  using Cfg = MlpCfg<D>;
  constexpr int kAtoms = Cfg::kAtoms, kYBytes = Cfg::kYBytes, kW1Slot = Cfg::kW1Slot,
                kW2Slot = Cfg::kW2Slot, kOffW1 = Cfg::kOffW1, kOffW2 = Cfg::kOffW2,
                kOffXb = Cfg::kOffXb, kAcc1Col = Cfg::kAcc1Col, kHalfN = Cfg::kHalfN,
                kW2Rows = Cfg::kW2Rows, kColsPerWarp = Cfg::kColsPerWarp, kSlots = Cfg::kSlots;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* y_full = bars + 0;
  uint64_t* y_empty = bars + 1;     // last GEMM1 of the tile retired (Y is dead)
  uint64_t* y_free = bars + 2;      // "all_free": every final-epilogue slot drained (16 arrivals)
  uint64_t* w1_full = bars + 3;     // [2]
  uint64_t* w1_empty = bars + 5;    // [2]
  uint64_t* w2_full = bars + 7;     // [2]
  uint64_t* w2_empty = bars + 9;    // [2]
  uint64_t* acc1_full = bars + 11;  // [2]
  uint64_t* h_ready = bars + 13;    // [2]  (leader: 32 warp arrivals)
  uint64_t* acc2_full = bars + 15;
  uint64_t* acc2_empty = bars + 16;  // (leader: 32 warp arrivals)
  uint64_t* rfull = bars + 17;       // [16 warps][3 slots max]
  uint64_t* p_full = bars + 65;      // PROJ: GEMM0 retired (acc2 = x1, O is dead)
  uint64_t* y_ready = bars + 66;     // PROJ: bf16(x1) written to Y (leader: 32 warp arrivals)
  uint64_t* x_done = bars + 67;      // PROJ: this CTA's residual slots are drained (16 arrivals):
                                     //       Y and the rings may be loaded for this tile
  uint64_t* x_loaded = bars + 68;    // PROJ: x + bp preloaded into acc2 (leader: 32 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 69);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cta_rank = __shfl_sync(0xffffffffu, static_cast<int>(cluster_ctarank()), 0);
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_pairs = (p.M + 255) / 256;
  const int NC = p.num_chunks;

  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmXB);
    if (PROJ && p.split_in) {
      tma_prefetch_desc(&tmXHi);
      tma_prefetch_desc(&tmXLoIn);
    }
    if (PROJ && p.lo_out) tma_prefetch_desc(&tmXLoOut);
    mbar_init(y_full, 1);
    mbar_init(y_empty, 1);
    mbar_init(y_free, 16);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&h_ready[i], 32);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 32);
    for (int i = 0; i < 48; ++i) mbar_init(&rfull[i], 1);
    mbar_init(p_full, 1);
    mbar_init(y_ready, 32);
    mbar_init(x_done, 16);
    mbar_init(x_loaded, 32);
    if (PROJ) tma_prefetch_desc(&tmWp);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  griddep_wait();  // xb / x / ln_stats are outputs of the previous kernels

  // Every cluster alternates a tensor-bound phase (hidden chunks, no HBM traffic to speak of)
  // with an HBM-bound one (final epilogue + refill: ~0.6 MB per CTA).  Started together, all
  // 148 SMs stream at the same time and compute at the same time; delaying every other cluster
  // by about half a tile lets one half compute while the other half owns the HBM.
  // Mode 1 (DEVIT_MLP_STAGGER_MODE=1): delay only the clusters that own one tile fewer than the
  // busiest ones -- they have a whole tile time of slack (198 pair-tiles over 74 clusters: 24 such
  // clusters; 99 tiles at N = 8: 49).  Measured 104.8 vs 106.4 us at 30 k clocks, inside the
  // run-to-run spread (profiles/r2_ab_projmlp_sm.txt): mode 0 stays the default.
  const bool delayed = p.stagger_mode == 1
      ? (num_pairs > num_clusters && cluster_id >= num_pairs % num_clusters &&
         num_pairs % num_clusters != 0)
      : (cluster_id & 1) != 0;
  if (p.stagger > 0 && delayed) {
    const long long t0 = clock64();
    while (clock64() - t0 < p.stagger) __nanosleep(500);
  }

  // width of chunk c (multiple of 16; only the last chunk can be narrower than 64)
  auto chunk_n = [&](int c) -> int {
    const int n = p.F_ld - c * 64;
    return n < 64 ? n : 64;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ loads: Y (or O) and the W1 ring
    uint32_t g1 = 0;
    int it = 0;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters, ++it) {
      const int m0 = (pt * 2 + cta_rank) * 128;
      if constexpr (PROJ) {
        // Y is dead once the previous tile's last GEMM1 has retired (the fp32 slots do not use it)
        if (it >= 1) mbar_wait_warp(y_empty, (it - 1) & 1);
      } else if (it >= 1) {
        mbar_wait_warp(y_free, (it - 1) & 1);
      }
      if (elect_one()) {
        const uint32_t bar = mapa_u32(smem_u32(y_full), 0);
        if constexpr (PROJ) {
          // the attention output tile, one 64-column atom per kept head
          if (leader) mbar_expect_tx(y_full, 2 * p.proj_chunks * 16384);
          for (int a = 0; a < p.proj_chunks; ++a)
            tma_load_2d_cg2(smem + a * 16384, &tmY, bar, a * 64, m0);
        } else {
          if (leader) mbar_expect_tx(y_full, 2 * kYBytes);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a) tma_load_2d_cg2(smem + a * 16384, &tmY, bar, a * 64, m0);
        }
      }
      // PROJ: the rings host this tile's residual slots until the preload has drained them
      if constexpr (PROJ) mbar_wait_warp(x_done, it & 1);
      for (int c = 0; c < NC; ++c, ++g1) {
        const int s = g1 & 1;
        mbar_wait_warp(&w1_empty[s], ((g1 >> 1) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t bar = mapa_u32(smem_u32(&w1_full[s]), 0);
          if (leader) mbar_expect_tx(&w1_full[s], 2 * kW1Slot);
          const int row0 = c * 64 + cta_rank * (chunk_n(c) / 2);
          uint8_t* dst = smem + kOffW1 + s * kW1Slot;
#pragma unroll
          for (int a = 0; a < kAtoms; ++a) tma_load_2d_cg2(dst + a * 4096, &tmW1, bar, a * 64, row0);
        }
      }
    }
  } else if (warp == kW2Warp) {
    // ------------------------------------------------------------ loads: the W2 ring
    uint32_t gw = 0;  // ring uses (PROJ: the Wp head chunks of a tile go through it first)
    int it = 0;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters, ++it) {
      // ring memory served as slots of the final epilogue (and, PROJ, of this tile's preload)
      if constexpr (PROJ) mbar_wait_warp(x_done, it & 1);
      else if (it >= 1) mbar_wait_warp(y_free, (it - 1) & 1);
      const int n_pre = PROJ ? p.proj_chunks : 0;
      for (int c = -n_pre; c < NC; ++c, ++gw) {
        const int s = gw & 1;
        mbar_wait_warp(&w2_empty[s], ((gw >> 1) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t bar = mapa_u32(smem_u32(&w2_full[s]), 0);
          if (leader) mbar_expect_tx(&w2_full[s], 2 * kW2Slot);
          uint8_t* dst = smem + kOffW2 + s * kW2Slot;
          // output columns [D/2 hh + D/4 rank, +D/4) of W2 (Wp), K = neurons [64c, 64c + 64)
          // (K = columns of head chunk c + n_pre)
          const CUtensorMap* tm = c < 0 ? &tmWp : &tmW2;
          const int k0 = (c < 0 ? c + n_pre : c) * 64;
          tma_load_2d_cg2(dst, tm, bar, k0, cta_rank * kW2Rows);
          tma_load_2d_cg2(dst + kW2Rows * 128, tm, bar, k0, kHalfN + cta_rank * kW2Rows);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue (leader only)
    if (leader) {
      uint32_t g1 = 0, g2 = 0, gw = 0;  // W1 ring / hidden chunks (GEMM2) / W2 ring uses
      int it = 0;
      const uint32_t idesc2 = make_idesc(kFmtBF16, 256, kHalfN, 0, 0);
      auto gemm2 = [&](int cc, bool first_of_tile) {
        const int b = g2 & 1;   // acc1 / H buffer of this hidden chunk
        const int ws = gw & 1;  // W2 ring slot
        MLP_TRACE(3, g2);
        mbar_wait_warp(&h_ready[b], (g2 >> 1) & 1);
        MLP_TRACE(4, g2);
        mbar_wait_warp(&w2_full[ws], (gw >> 1) & 1);
        MLP_TRACE(5, g2);
        if (!PROJ && first_of_tile && it >= 1) mbar_wait_warp(acc2_empty, (it - 1) & 1);
        tc_fence_after();
        const uint32_t sw = smem_u32(smem + kOffW2 + ws * kW2Slot);
        const int ksteps = chunk_n(cc) / 16;
        if (elect_one()) {
          for (int j = 0; j < ksteps; ++j) {
            // H of neurons [16j, 16j + 16) of the chunk sits in columns [16j, 16j + 8)
            const uint32_t a_t = tmem_base + kAcc1Col + 64 * b + 16 * j;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint64_t db = make_sw128_desc(sw + hh * (kW2Rows * 128), 1024, 16) + 2 * j;
              // PROJ: acc2 already holds x1, every GEMM2 accumulates
              umma_bf16_ts_cg2(tmem_base + kHalfN * hh, a_t, db, idesc2,
                               (!PROJ && first_of_tile && j == 0) ? 0u : 1u);
            }
          }
          umma_commit_cg2(&w2_empty[ws], 3);
        }
        MLP_TRACE(6, g2);
        ++g2;
        ++gw;
      };
      for (int pt = cluster_id; pt < num_pairs; pt += num_clusters, ++it) {
        MLP_TRACE(7, it);
        mbar_wait_warp(y_full, it & 1);
        MLP_TRACE(8, it);
        if constexpr (PROJ) {
          // GEMM0: acc2 (= x + bp, preloaded by the epilogue warps of both CTAs) += O Wp^T, one
          // 64-wide head chunk per W2-ring slot
          mbar_wait_warp(x_loaded, it & 1);
          MLP_TRACE(0, 400 + it * 8 + 7);
          for (int a = 0; a < p.proj_chunks; ++a, ++gw) {
            const int ws = gw & 1;
            mbar_wait_warp(&w2_full[ws], (gw >> 1) & 1);
            MLP_TRACE(1, 400 + it * 8 + a);
            tc_fence_after();
            const uint64_t da = make_sw128_desc(smem_u32(smem) + a * 16384, 1024, 16);
            const uint32_t sw = smem_u32(smem + kOffW2 + ws * kW2Slot);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const uint64_t db = make_sw128_desc(sw + hh * (kW2Rows * 128), 1024, 16);
                  umma_bf16_cg2(tmem_base + kHalfN * hh, da + 2 * k, db + 2 * k, idesc2, 1u);
                }
              }
              umma_commit_cg2(&w2_empty[ws], 3);
            }
          }
          if (elect_one()) umma_commit_cg2(p_full, 3);
          MLP_TRACE(18, it);
          // epilogue 0 of both CTAs: bf16(x1) in Y
          mbar_wait_warp(y_ready, it & 1);
          tc_fence_after();
          MLP_TRACE(19, it);
        }
        for (int c = 0; c < NC; ++c) {
          const int s = g1 & 1;
          MLP_TRACE(0, g1);
          mbar_wait_warp(&w1_full[s], (g1 >> 1) & 1);
          MLP_TRACE(1, g1);
          tc_fence_after();
          const uint32_t idesc1 = make_idesc(kFmtBF16, 256, chunk_n(c), 0, 0);
          const uint32_t sy = smem_u32(smem);
          const uint32_t sw = smem_u32(smem + kOffW1 + s * kW1Slot);
          if (elect_one()) {
#pragma unroll
            for (int a = 0; a < kAtoms; ++a) {
              const uint64_t da = make_sw128_desc(sy + a * 16384, 1024, 16);
              const uint64_t db = make_sw128_desc(sw + a * 4096, 1024, 16);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_cg2(tmem_base + kAcc1Col + 64 * s, da + 2 * k, db + 2 * k, idesc1,
                              (a | k) ? 1u : 0u);
            }
            umma_commit_cg2(&w1_empty[s], 3);
            umma_commit_cg2(&acc1_full[s], 3);
            if (c == NC - 1) umma_commit_cg2(y_empty, 3);
          }
          MLP_TRACE(2, g1);
          ++g1;
          if (c >= 1) gemm2(c - 1, c == 1);
        }
        gemm2(NC - 1, NC == 1);
        if (elect_one()) umma_commit_cg2(acc2_full, 3);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;        // TMEM lane quarter
    const int sub = (warp - 2) >> 2;     // 16-column slice of every hidden chunk; final epilogue:
                                         // output columns [96 sub, +96)
    const int ew = sub * 4 + quarter;    // slot owner index (0..15)
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    // slots 3 ew .. 3 ew + 2 of the 48: [0,24) in Y, [24,36) in the W1 ring, [36,48) in the W2
    // ring (contiguous in shared memory, so the slot address is simply 4 KB x index)
    uint8_t* slots = smem + ew * (kSlots * 4096);
    // slots that lie in the W2 ring are usable only once the last GEMM2 has retired
    const bool slots_in_w2 = (ew + 1) * kSlots * 4096 > kOffW2;
    uint64_t* rbar = rfull + ew * 3;
    uint8_t* xb_stg = smem + kOffXb + ew * 2048;  // [32 rows x 32 bf16], rows of 64 B, 64B swizzle
    const uint32_t h_ready_leader0 = mapa_u32(smem_u32(&h_ready[0]), 0);
    const uint32_t h_ready_leader1 = mapa_u32(smem_u32(&h_ready[1]), 0);
    const uint32_t acc2_empty_leader = mapa_u32(smem_u32(acc2_empty), 0);
    const uint32_t y_ready_leader = mapa_u32(smem_u32(y_ready), 0);
    const uint32_t x_loaded_leader = mapa_u32(smem_u32(x_loaded), 0);
    // PROJ: this warp's two private fp32 slots.  Final epilogue: chunk j is staged in slot
    // j & 1.  Residual of the next tile: chunk j goes to slot rs(j), chosen so that the slot whose
    // store was committed FIRST is re-armed first (3 chunks: stores used slots 0,1,0 -> residual
    // chunks use 1,0,1; 2 chunks: 0,1 -> 0,1).
    uint8_t* pslot = smem + Cfg::kOffPSlots + ew * 2 * 4096;
    auto rs = [](int j) -> int { return (j + (kSlots == 3 ? 1 : 0)) & 1; };
    uint32_t pphase = 0;  // parity bits of rbar[0], rbar[1]
    // request residual chunk j of tile `m_tile` into slot rs(j) (the slot must be drained)
    auto request_resid = [&](int m_tile, int j) {
      if (elect_one()) {
        uint8_t* dst = pslot + rs(j) * 4096;
        mbar_expect_tx(&rbar[rs(j)], 4096);
        if (p.split_in) {  // hi plane -> first 2 KB, lo plane -> second 2 KB ([32 x 64 B], SW64)
          tma_load_2d(dst, &tmXHi, &rbar[rs(j)], sub * kColsPerWarp + j * 32, m_tile + quarter * 32);
          tma_load_2d(dst + 2048, &tmXLoIn, &rbar[rs(j)], sub * kColsPerWarp + j * 32,
                      m_tile + quarter * 32);
        } else {
          tma_load_2d(dst, &tmX, &rbar[rs(j)], sub * kColsPerWarp + j * 32,
                      m_tile + quarter * 32);
        }
      }
      __syncwarp();
    };
    uint32_t ecnt = 0;
    int it = 0;
    for (int pt = cluster_id; pt < num_pairs; pt += num_clusters, ++it) {
      const int m0 = (pt * 2 + cta_rank) * 128;
      const int row0 = m0 + quarter * 32;
      const int row = row0 + lane;
      const bool has_next = pt + num_clusters < num_pairs;
      const int m_next = ((pt + num_clusters) * 2 + cta_rank) * 128;
      // ---- this row's LayerNorm statistics (folded into the fc1 epilogue)
      float rstd, nmr;
      if constexpr (PROJ) {
        float2* stat_s = reinterpret_cast<float2*>(smem + Cfg::kOffStat);
        const int rr = quarter * 32 + lane;  // row inside the CTA's 128
        // ---- preload: acc2 <- x + bp.  The residual chunks were requested into this warp's own
        //      slots while the previous tile's stores drained (here for the CTA's first tile).
        if (it == 0) {
          request_resid(m0, 0);
          if (kSlots >= 2) request_resid(m0, 1);
        }
        if (warp == 2) MLP_TRACE(12, it);
#pragma unroll 1
        for (int j = 0; j < kSlots; ++j) {
          const int col0 = sub * kColsPerWarp + j * 32;
          const int sl = rs(j);
          uint32_t r[32];
          float* v = reinterpret_cast<float*>(r);
          mbar_wait_warp(&rbar[sl], (pphase >> sl) & 1u);
          pphase ^= 1u << sl;
          const uint8_t* bsl = pslot + sl * 4096;
          if (p.split_in) {
            // x = hi + lo: this lane's row is 64 B of each plane, 16-byte chunks 64B-swizzled
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int so = lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
              const uint4 h4 = *reinterpret_cast<const uint4*>(bsl + so);
              const uint4 l4 = *reinterpret_cast<const uint4*>(bsl + 2048 + so);
              const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[8 * g + 2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
                v[8 * g + 2 * e + 1] = __uint_as_float(hw[e] & 0xffff0000u) +
                                       __uint_as_float(lw[e] & 0xffff0000u);
              }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bp + col0) + g);
              v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 t = *reinterpret_cast<const float4*>(bsl + mlp_off(lane, g));
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bp + col0) + g);
              v[4 * g] = t.x + b.x; v[4 * g + 1] = t.y + b.y;
              v[4 * g + 2] = t.z + b.z; v[4 * g + 3] = t.w + b.w;
            }
          }
          if (j + 2 < kSlots) {  // the slot has been read by every lane: fetch chunk j + 2
            __syncwarp();
            fence_proxy_async_smem();
            request_resid(m0, j + 2);
          }
          tmem_st_x32(tmem_base + lane_off + col0, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(x_done);  // slots drained: the loaders may fill Y and the rings
          if (leader) mbar_arrive(x_loaded);
          else mbar_arrive_cluster(x_loaded_leader);
        }
        if (warp == 2) MLP_TRACE(13, it);
        // ---- epilogue 0: x1 = acc2 after GEMM0: row sums -> shared memory, bf16(x1) -> Y
        //      (K-major, 128B swizzle)
        mbar_wait_warp(p_full, it & 1);
        if (warp == 2) MLP_TRACE(15, it * 4);
        tc_fence_after();
        float st1 = 0.f, st2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < kSlots; ++j) {
          const int col0 = sub * kColsPerWarp + j * 32;
          uint32_t r[32];
          tmem_ld_x32(tmem_base + lane_off + col0, r);
          tmem_ld_wait();
          const float* v = reinterpret_cast<const float*>(r);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            st1 += v[k];
            st2 = fmaf(v[k], v[k], st2);
          }
          uint8_t* ya = smem + (col0 >> 6) * 16384 + rr * 128;
          const int cb = (col0 & 63) >> 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 t;
            t.x = pack_bf16x2(v[8 * g], v[8 * g + 1]);
            t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
            t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
            t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
            *reinterpret_cast<uint4*>(ya + (((cb + g) ^ (rr & 7)) << 4)) = t;
          }
        }
        stat_s[sub * 128 + rr] = make_float2(st1, st2);
        fence_proxy_async_smem();  // Y as the tensor core (async proxy) will read it
        tc_fence_before();
        named_bar_sync(1, 16 * 32);  // the sixteen epilogue warps of this CTA
        if (lane == 0) {
          if (leader) mbar_arrive(y_ready);
          else mbar_arrive_cluster(y_ready_leader);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 t = stat_s[q * 128 + rr];
          s1 += t.x;
          s2 += t.y;
        }
        const float mean = s1 * p.ln_inv_dim;
        const float var = fmaxf(fmaf(s2, p.ln_inv_dim, -mean * mean), 0.f);
        rstd = rsqrtf(var + p.ln_eps);
        nmr = -rstd * mean;
        if (warp == 2) MLP_TRACE(14, it);
      } else {
        float s1 = 0.f, s2 = 0.f;
        if (row < p.M) {
          for (int q = 0; q < p.ln_parts; ++q) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(p.ln_stats) +
                                   static_cast<long long>(q) * p.M + row);
            s1 += t.x;
            s2 += t.y;
          }
        }
        const float mean = s1 * p.ln_inv_dim;
        const float var = fmaxf(fmaf(s2, p.ln_inv_dim, -mean * mean), 0.f);
        rstd = rsqrtf(var + p.ln_eps);
        nmr = -rstd * mean;
        // The residual of this tile is needed only in the final epilogue, and its slots do not
        // exist until then: pull it into L2 now, so that the late TMA loads are L2 hits.
        if (row0 < p.M && elect_one()) {
#pragma unroll
          for (int s = 0; s < kSlots; ++s) tma_prefetch_l2_2d(&tmX, sub * kColsPerWarp + s * 32, row0);
        }
      }
      // ---- hidden chunks: acc1 -> H (bf16, in TMEM).  Slice `sub` of chunk c: neurons
      //      [64c + 16 sub, +16) = acc1 columns [16 sub, +16) -> H columns [16 sub, +8).
      for (int c = 0; c < NC; ++c, ++ecnt) {
        const int b = ecnt & 1;
        if constexpr (PROJ) {
          // next tile's residual and attention output -> L2, late enough to still be there when
          // the TMA loads ask for them (~half a chunk loop ahead)
          if (c == NC / 2 && has_next && elect_one()) {
#pragma unroll
            for (int s2 = 0; s2 < kSlots; ++s2) {
              if (p.split_in) {  // the residual comes as hi / lo planes
                tma_prefetch_l2_2d(&tmXHi, sub * kColsPerWarp + s2 * 32, m_next + quarter * 32);
                tma_prefetch_l2_2d(&tmXLoIn, sub * kColsPerWarp + s2 * 32, m_next + quarter * 32);
              } else {
                tma_prefetch_l2_2d(&tmX, sub * kColsPerWarp + s2 * 32, m_next + quarter * 32);
              }
            }
            if (sub == 0 && quarter < 2)
              for (int a = quarter; a < p.proj_chunks; a += 2)
                tma_prefetch_l2_2d(&tmY, a * 64, m_next);
          }
        }
        // fold constants of this slice, fetched and combined BEFORE the accumulator is awaited
        float t[16];
        {
          const int col0 = c * 64 + 16 * sub;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), bs = cs;
            if (col0 + 4 * g < p.F_ld) {
              cs = __ldg(reinterpret_cast<const float4*>(p.c1 + col0) + g);
              bs = __ldg(reinterpret_cast<const float4*>(p.c2 + col0) + g);
            }
            t[4 * g] = fmaf(nmr, cs.x, bs.x);
            t[4 * g + 1] = fmaf(nmr, cs.y, bs.y);
            t[4 * g + 2] = fmaf(nmr, cs.z, bs.z);
            t[4 * g + 3] = fmaf(nmr, cs.w, bs.w);
          }
        }
        if (warp == 2) MLP_TRACE(9, ecnt);
        mbar_wait_warp(&acc1_full[b], (ecnt >> 1) & 1);
        if (warp == 2) MLP_TRACE(10, ecnt);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + lane_off + kAcc1Col + 64 * b + 16 * sub;
        uint32_t r[16];
        tmem_ld_x16(t_acc, r);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float v0 = gelu_erf_fast(fmaf(rstd, __uint_as_float(r[2 * g]), t[2 * g]));
          const float v1 = gelu_erf_fast(fmaf(rstd, __uint_as_float(r[2 * g + 1]), t[2 * g + 1]));
          pk[g] = pack_bf16x2(v0, v1);
        }
        tmem_st_x8(t_acc, pk);  // over the first 8 of the 16 columns this thread just consumed
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&h_ready[b]);
          else mbar_arrive_cluster(b ? h_ready_leader1 : h_ready_leader0);
        }
        if (warp == 2) MLP_TRACE(11, ecnt);
      }
      // ---- final epilogue: x += acc2 + b2, bf16 copy, partial row sums (96 columns per warp)
      //      (PROJ: acc2 already holds x1 + the MLP branch; the dead Y / ring memory stages the
      //      stores, and each slot is re-armed with the NEXT tile's residual as soon as its store
      //      has been read)
      if constexpr (!PROJ) {
        if (slots_in_w2) mbar_wait_warp(acc2_full, it & 1);
        else mbar_wait_warp(y_empty, it & 1);
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < kSlots; ++s) {
            mbar_expect_tx(&rbar[s], 4096);
            tma_load_2d(slots + s * 4096, &tmX, &rbar[s], sub * kColsPerWarp + s * 32, row0);
          }
        }
        // ... and the next tile's Y (loaded only after every slot has been drained) likewise
        if (warp == 2 && pt + num_clusters < num_pairs && elect_one()) {
          const int m_next = ((pt + num_clusters) * 2 + cta_rank) * 128;
#pragma unroll
          for (int a = 0; a < kAtoms; ++a) tma_prefetch_l2_2d(&tmY, a * 64, m_next);
        }
      }
      if (!PROJ && warp == 2) MLP_TRACE(12, it);
      mbar_wait_warp(acc2_full, it & 1);
      if (!PROJ && warp == 2) MLP_TRACE(13, it);
      if (PROJ && warp == 2) MLP_TRACE(16, it * 4 + 3);
      tc_fence_after();
      float st1 = 0.f, st2 = 0.f;
#pragma unroll 1
      for (int j = 0; j < kSlots; ++j) {
        const int col0 = sub * kColsPerWarp + j * 32;
        uint32_t r[32];
        tmem_ld_x32(tmem_base + lane_off + col0, r);
        tmem_ld_wait();
        if (!PROJ && j == kSlots - 1) {
          // acc2 is in registers: release it before the memory work of the last chunk
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(acc2_empty);
            else mbar_arrive_cluster(acc2_empty_leader);
          }
        }
        float* v = reinterpret_cast<float*>(r);
        uint8_t* bsl = PROJ ? pslot + (j & 1) * 4096 : slots + j * 4096;
        if constexpr (PROJ) {
          // two slots: chunk j reuses the slot of chunk j - 2 once that store has been read
          if (j >= 2) {
            if (elect_one()) bulk_wait_read<1>();
            __syncwarp();
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.b2 + col0) + g);
          v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
        }
        if constexpr (!PROJ) {
          if (warp == 2) MLP_TRACE(15, it * 4 + j);
          mbar_wait_warp(&rbar[j], it & 1);
          if (warp == 2) MLP_TRACE(16, it * 4 + j);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(bsl + mlp_off(lane, g));
            v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
          }
        }
        if (!(PROJ && p.lo_out)) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(bsl + mlp_off(lane, g)) =
                make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        if constexpr (PROJ) {
          if (p.lo_out) {
            // x -> hi = bf16(x), lo = bf16(x - hi): two [32 x 64 B] planes staged in this slot
            // (64B swizzle) and stored with one TMA store each: 4 bytes per element leave the SM
            // instead of 6 (fp32 x + bf16 copy)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = v[8 * g + 2 * e], b = v[8 * g + 2 * e + 1];
                hw[e] = pack_bf16x2(a, b);
                lw[e] = pack_bf16x2(a - __uint_as_float(hw[e] << 16),
                                    b - __uint_as_float(hw[e] & 0xffff0000u));
              }
              const int so = lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
              *reinterpret_cast<uint4*>(bsl + so) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(bsl + 2048 + so) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            fence_proxy_async_smem();
            if (elect_one()) {
              tma_store_2d(&tmXB, bsl, col0, row0);
              tma_store_2d(&tmXLoOut, bsl + 2048, col0, row0);
              bulk_commit();
            }
            __syncwarp();
          } else {
          // bf16 copy straight from registers: this lane's 32 values are 64 contiguous bytes of
          // its row = two full 32-byte sectors (STG.256): no staging tile.
          if (p.xb_out && row < p.M) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(p.xb_out + static_cast<long long>(row) * D + col0);
#pragma unroll
            for (int g = 0; g < 2; ++g)
              st_global_256(dst + 8 * g, pack_bf16x2(v[16 * g], v[16 * g + 1]),
                            pack_bf16x2(v[16 * g + 2], v[16 * g + 3]),
                            pack_bf16x2(v[16 * g + 4], v[16 * g + 5]),
                            pack_bf16x2(v[16 * g + 6], v[16 * g + 7]),
                            pack_bf16x2(v[16 * g + 8], v[16 * g + 9]),
                            pack_bf16x2(v[16 * g + 10], v[16 * g + 11]),
                            pack_bf16x2(v[16 * g + 12], v[16 * g + 13]),
                            pack_bf16x2(v[16 * g + 14], v[16 * g + 15]));
          }
          fence_proxy_async_smem();
          if (elect_one()) {
            tma_store_2d(&tmX, bsl, col0, row0);
            bulk_commit();
          }
          __syncwarp();
          }
        } else {
          if (j > 0 && p.xb_out) {
            // the previous chunk's stores have read their staging tile
            if (elect_one()) bulk_wait_read<0>();
            __syncwarp();
          }
          if (p.xb_out) {
            // bf16 copy through a 64B-swizzled staging tile + TMA store (scattered 16-byte global
            // stores from 16 warps were the slowest part of this phase).
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t;
              t.x = pack_bf16x2(v[8 * g], v[8 * g + 1]);
              t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
              t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
              t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
              *reinterpret_cast<uint4*>(xb_stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = t;
            }
          }
          fence_proxy_async_smem();
          if (elect_one()) {
            tma_store_2d(&tmX, bsl, col0, row0);
            if (p.xb_out) tma_store_2d(&tmXB, xb_stg, col0, row0);
            bulk_commit();
          }
          __syncwarp();
        }
        if (p.stats_out) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            st1 += v[k];
            st2 = fmaf(v[k], v[k], st2);
          }
        }
        if (warp == 2) MLP_TRACE(17, it * 4 + j);
      }
      // partial row sums of this warp's D/4 columns: part index = sub (4 parts per row)
      if (p.stats_out && row < p.M)
        reinterpret_cast<float2*>(p.stats_out)[static_cast<long long>(sub) * p.M + row] =
            make_float2(st1, st2);
      // the loaders may reuse Y and the rings once every warp's stores have drained its slots
      // (PROJ: ... and the next tile's preload has consumed the residual re-armed into them)
      if constexpr (PROJ) {
        // re-arm the slots with the next tile's residual as soon as their stores have been read
        // (one bulk group per chunk, oldest first): chunk 0 -> rs(0) after all but the newest
        // group, chunk 1 -> rs(1) after the newest; a third chunk follows in the preload
        if (elect_one()) bulk_wait_read<1>();
        __syncwarp();
        if (has_next) request_resid(m_next, 0);
        if (elect_one()) bulk_wait_read<0>();
        __syncwarp();
        if (has_next && kSlots >= 2) request_resid(m_next, 1);
      } else {
        if (elect_one()) {
          bulk_wait_read<0>();
          mbar_arrive(y_free);
        }
        __syncwarp();
      }
      if (!PROJ && warp == 2) MLP_TRACE(14, it);
      if (PROJ && warp == 2) MLP_TRACE(15, it * 4 + 3);
    }
    if (elect_one()) bulk_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

}  // namespace devit

using namespace devit;

extern "C" int devit_mlp_fused(const devit_mlp_args* a, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  DEVIT_REQUIRE(a != nullptr, "devit_mlp_fused: null args");
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(a->dim == 256 || a->dim == 384,
                "devit_mlp_fused: dim %d unsupported (built for 256 and 384)", a->dim);
  const int Dm = a->dim;
  const bool proj = a->o != nullptr;
  DEVIT_REQUIRE(a->m > 0 && a->hidden_ld >= 16 && a->hidden_ld % 16 == 0,
                "devit_mlp_fused: need m > 0 and hidden_ld a positive multiple of 16");
  DEVIT_REQUIRE(a->w1 && a->c1 && a->c2 && a->w2 && a->b2 && a->x,
                "devit_mlp_fused: null pointer");
  if (proj) {
    DEVIT_REQUIRE(a->w_proj && a->b_proj, "devit_mlp_fused: o given without w_proj / b_proj");
    DEVIT_REQUIRE(a->proj_k >= 64 && a->proj_k % 64 == 0 && a->proj_k <= Dm,
                  "devit_mlp_fused: proj_k %d must be a multiple of 64 in [64, dim]", a->proj_k);
    DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(a->b_proj) % 16 == 0,
                  "devit_mlp_fused: b_proj must be 16-byte aligned");
  } else {
    DEVIT_REQUIRE(a->xb && a->ln_stats, "devit_mlp_fused: null pointer");
    DEVIT_REQUIRE(a->ln_parts >= 1, "devit_mlp_fused: ln_parts must be >= 1");
  }
  DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(a->c1) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(a->c2) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(a->b2) % 16 == 0 &&
                    (!a->xb_out || reinterpret_cast<uintptr_t>(a->xb_out) % 32 == 0),
                "devit_mlp_fused: c1 / c2 / b2 must be 16-byte, xb_out 32-byte aligned");
  static bool attr_done[64] = {};
  int dev = 0;
  DEVIT_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    DEVIT_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<384, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       MlpCfg<384>::kSmem));
    DEVIT_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<256, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       MlpCfg<256>::kSmem));
    DEVIT_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<384, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       MlpCfg<384>::kSmem));
    DEVIT_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<256, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       MlpCfg<256>::kSmem));
    attr_done[dev & 63] = true;
  }
  CUtensorMap tY, tW1, tW2, tX, tXB, tWp, tXHi, tXLoIn, tXLoOut;
  if (proj) {
    rc = encode_tmap_2d(&tY, a->o, 2, a->proj_k, a->m, a->proj_k, 64, 128, false);
    if (rc) return rc;
    rc = encode_tmap_2d(&tWp, a->w_proj, 2, a->proj_k, Dm, a->proj_k, 64, Dm / 4, true);
    if (rc) return rc;
  } else {
    rc = encode_tmap_2d(&tY, a->xb, 2, Dm, a->m, Dm, 64, 128, false);
    if (rc) return rc;
  }
  rc = encode_tmap_2d(&tW1, a->w1, 2, Dm, a->hidden_ld, Dm, 64, 32, true);
  if (rc) return rc;
  rc = encode_tmap_2d(&tW2, a->w2, 2, a->hidden_ld, Dm, a->hidden_ld, 64, Dm / 4, true);
  if (rc) return rc;
  if (!proj) tWp = tW2;
  rc = encode_tmap_2d(&tX, a->x, 4, Dm, a->m, Dm, 32, 32, false);
  if (rc) return rc;
  tXB = tX;
  if (a->xb_out) {
    rc = encode_tmap_2d_sw64(&tXB, a->xb_out, 2, Dm, a->m, Dm, 32, 32);
    if (rc) return rc;
  }
  tXHi = tXLoIn = tXLoOut = tXB;
  if (a->x_lo_in) {
    DEVIT_REQUIRE(proj && a->xb, "devit_mlp_fused: x_lo_in needs the fused projection (o) and xb "
                  "(the hi plane)");
    rc = encode_tmap_2d_sw64(&tXHi, a->xb, 2, Dm, a->m, Dm, 32, 32);
    if (rc) return rc;
    rc = encode_tmap_2d_sw64(&tXLoIn, a->x_lo_in, 2, Dm, a->m, Dm, 32, 32);
    if (rc) return rc;
  }
  if (a->x_lo_out) {
    DEVIT_REQUIRE(proj && a->xb_out, "devit_mlp_fused: x_lo_out needs the fused projection (o) "
                  "and xb_out (the hi plane)");
    rc = encode_tmap_2d_sw64(&tXLoOut, a->x_lo_out, 2, Dm, a->m, Dm, 32, 32);
    if (rc) return rc;
  }
  MlpParams p;
  p.M = a->m;
  p.F_ld = a->hidden_ld;
  p.num_chunks = (a->hidden_ld + 63) / 64;
  p.c1 = a->c1;
  p.c2 = a->c2;
  p.b2 = a->b2;
  p.ln_stats = a->ln_stats;
  p.ln_parts = a->ln_parts;
  p.ln_inv_dim = 1.0f / static_cast<float>(Dm);
  p.ln_eps = a->ln_eps;
  p.xb_out = static_cast<__nv_bfloat16*>(a->xb_out);
  p.stats_out = a->stats_out;
  p.bp = a->b_proj;
  p.proj_chunks = proj ? a->proj_k / 64 : 0;
  p.split_in = a->x_lo_in ? 1 : 0;
  p.lo_out = static_cast<__nv_bfloat16*>(a->x_lo_out);
  p.trace = g_attn_trace;
  {
    static int stagger = -1;  // DEVIT_MLP_STAGGER=<clocks> (0 = off)
    if (stagger < 0) {
      const char* e = getenv("DEVIT_MLP_STAGGER");
      stagger = e ? atoi(e) : 10000;  // measured best of {0, 10k, 18k, 26k}: -1.5 .. -3.5 %
    }
    p.stagger = stagger;
    static int mode_cache = kEnvUnread;
    p.stagger_mode = env_int("DEVIT_MLP_STAGGER_MODE", 0, &mode_cache);
  }
  const int num_pairs = (a->m + 255) / 256;
  int clusters = num_sms() / 2;
  if (clusters > num_pairs) clusters = num_pairs;
  {
    ProfScope ps(kTagMlpFused, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(kMlpThreads);
    cfg.dynamicSmemBytes = Dm == 384 ? MlpCfg<384>::kSmem : MlpCfg<256>::kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    auto kern = Dm == 384 ? (proj ? mlp_fused_kernel<384, true> : mlp_fused_kernel<384, false>)
                          : (proj ? mlp_fused_kernel<256, true> : mlp_fused_kernel<256, false>);
    DEVIT_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tY, tW1, tW2, tX, tXB, tWp, tXHi, tXLoIn, tXLoOut,
                                     p));
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

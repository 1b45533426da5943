// Persistent, warp-specialised tcgen05 GEMM with a fused epilogue (see include/devit_b200.h,
// devit_gemm).  out = epilogue(sum_s A_s * B_s^T), A and B K-major, fp32 accumulation in TMEM.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 2..9 = epilogue (two warps per TMEM lane quarter, alternating column chunks).
// Pipelines:
//   smem ring : kStages slots of {A k-block 16 KB | B k-block}, full[s] / empty[s] mbarriers
//               between the TMA producer and the MMA issuer.
//   TMEM ring : two accumulator stages (tmem_full / tmem_empty) so the epilogue of tile i
//               overlaps the MMAs of tile i+1.
//   stores    : registers -> 128B-swizzled shared memory -> TMA store (coalesced, edge-clipped);
//               an fp32 residual tile is TMA-prefetched chunk by chunk into the same staging
//               buffers, summed in place and stored from there.
//   tile loop : static round-robin, n fastest so co-resident CTAs share A rows in L2.
// Tile = 128 x BN (BN in {128,192,256}); a k-block is one 128-byte swizzle atom along K
// (64 bf16 or 32 tf32) = 4 UMMAs.  The last N tile issues a narrower UMMA (N rounded up to
// 16) so ragged shrunk widths do not pay for a full tile.
// CL = 2 (cta_group::2): a CTA pair computes a 256 x BN tile with one MMA stream issued by the
// leader; each CTA stages its own 128 A rows and HALF of the B rows.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace devit {

constexpr int kBlockM = 128;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxLnParts = 12;  // partial row sums a LayerNorm-folded GEMM can combine
constexpr int kMaxKSegs = 24;  // 8 logical segments x 3 passes in the 3xTF32 mode

struct KSeg {
  int a_row_off, a_k_off, b_k_off, k_blocks;
  int a_plane, b_plane;
};

struct GemmKParams {
  int M, N;
  int num_segs;
  int total_kb;  // sum of k_blocks over the segments
  KSeg segs[kMaxKSegs];
  void* out;
  long long ldo;
  int out_kind;
  long long out_plane_stride;
  const float* bias;
  const float* resid;
  long long ldr;
  int resid_period;  // > 0: the residual of row m is row (m % resid_period) of a wrapped table
  const float* rowbias;
  long long ld_rowbias;
  int act;
  float alpha;
  int rowmap_period, rowmap_stride, rowmap_off;
  // LayerNorm folding (see devit_gemm_args)
  const float* ln_stats;   // consumer: [ln_parts][M][2]
  int ln_parts;
  float ln_inv_dim, ln_eps;
  const float* ln_colsum;  // consumer: c1[n]
  float* stats_out;        // producer: [2 * N / 128][M][2]
  int xb_out;              // producer: 1 = also TMA-store the bf16 copy (tensor map tmO1)
  int vec_ok;
  int tma_epi;  // 1: coalesced epilogue (smem staging + TMA store, residual via R-slots)
  int dbg;      // DEVIT_GEMM_DBG bit0: no epilogue work, bit1: no TMA loads, bit2: no MMAs
  long long* trace;  // optional clock64 trace buffer (CTA 0 only), see devit_debug_set_trace
};

static long long* g_trace = nullptr;

#ifdef DEVIT_GEMM_TRACE
#define DEVIT_TRACE(slot_, idx_)                                                  \
  do {                                                                            \
    if (p.trace && blockIdx.x == 0 && (idx_) < 512)                               \
      p.trace[(slot_) * 512 + (idx_)] = clock64();                                \
  } while (0)
#else
#define DEVIT_TRACE(slot_, idx_) do { (void)(idx_); } while (0)
#endif

constexpr int kMaxResidentKb = 6;  // B-resident mode: K <= 6 k-blocks (384 bf16)

// BRES ("B resident"): for K <= 384 the CTA (pair) keeps ONE n-tile's weight k-blocks in shared
// memory for the whole kernel and walks down the m-tiles of that n-tile, so only the activation
// tile is streamed: half the operand bytes per tile of the ordinary mode, which matters because
// these K = 384 GEMMs are bound by the bytes an SM pulls through L2.
template <int BN, int CL, bool RES = false, bool BRES = false>
struct GemmCfg {
  static constexpr int kStageA = kBlockM * 128;
  static constexpr int kStageB = (BN / CL) * 128;  // a CTA pair splits B's rows half / half
  static constexpr int kStageBytes = BRES ? kStageA : kStageA + kStageB;
  static constexpr int kResidentBytes = BRES ? kMaxResidentKb * kStageB : 0;
  // one [32 x 128 B] staging buffer per warp; the residual variant keeps a slot for every
  // 32x32 fp32 chunk of the tile so the NEXT tile's residual is in flight during this one
  // (+ one bf16 staging buffer per warp for the optional bf16 copy of the output)
  static constexpr int kEpiBytes = RES ? kBlockM * BN * 4 + kEpiWarps * 4096 : kEpiWarps * 4096;
  // per-warp copies of the tile's bias and (LayerNorm folding) column sums
  static constexpr int kBiasBytes = 2 * kEpiWarps * BN * 4;
  static constexpr int kBarBytes = (2 * 8 + 32 + 4) * 8 + 16 + 32;
  static constexpr int kBudget =
      227 * 1024 - 1024 - kEpiBytes - kBiasBytes - kBarBytes - kResidentBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kAccStride = BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kRingBytes = kStages * kStageBytes + kResidentBytes;
  static constexpr int kSmemBytes = kRingBytes + kEpiBytes + kBiasBytes + kBarBytes + 1024;
  static_assert(kStages >= (RES ? 2 : 3), "not enough shared memory for the operand ring");
};

// byte offset of 16-byte chunk j of row r inside a [rows x 128 B] 128B-swizzled buffer
__device__ __forceinline__ int stg_off(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }

// -------------------------------------------------------------------- direct (fallback) path
// Applies the epilogue to `cnt` consecutive columns held in v[] and stores them with plain
// global accesses (used for the row-remapped patch embedding and unaligned / tiny outputs).
template <bool FULL>
__device__ __forceinline__ void epilogue_store(const GemmKParams& p, float* v, int cnt,
                                               long long row_out, int rb_row, int col0) {
  if (p.bias) {
    if (FULL) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = __ldg(b4 + j);
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
      }
    } else {
      for (int j = 0; j < cnt; ++j) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (p.act == DEVIT_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (p.act == DEVIT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.rowbias) {
    const float* rb = p.rowbias + (long long)rb_row * p.ld_rowbias + col0;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = __ldg(reinterpret_cast<const float4*>(rb) + j);
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
      }
    } else {
      for (int j = 0; j < cnt; ++j) v[j] += __ldg(rb + j);
    }
  }
  if (p.resid) {
    const float* rs = p.resid + row_out * p.ldr + col0;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = *(reinterpret_cast<const float4*>(rs) + j);
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
      }
    } else {
      for (int j = 0; j < cnt; ++j) v[j] += rs[j];
    }
  }
  if (p.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  }
  if (p.out_kind == DEVIT_OUT_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row_out * p.ldo + col0;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 t;
        t.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
        t.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        t.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        t.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        *(reinterpret_cast<uint4*>(o) + j) = t;
      }
    } else {
      for (int j = 0; j < cnt; ++j) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else if (p.out_kind == DEVIT_OUT_F32) {
    float* o = reinterpret_cast<float*>(p.out) + row_out * p.ldo + col0;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *(reinterpret_cast<float4*>(o) + j) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      for (int j = 0; j < cnt; ++j) o[j] = v[j];
    }
  } else {  // split fp32: hi plane + lo plane
    float* o = reinterpret_cast<float*>(p.out) + row_out * p.ldo + col0;
    float* ol = o + p.out_plane_stride;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float h0 = tf32_hi(v[4 * j]), h1 = tf32_hi(v[4 * j + 1]), h2 = tf32_hi(v[4 * j + 2]),
              h3 = tf32_hi(v[4 * j + 3]);
        *(reinterpret_cast<float4*>(o) + j) = make_float4(h0, h1, h2, h3);
        *(reinterpret_cast<float4*>(ol) + j) =
            make_float4(v[4 * j] - h0, v[4 * j + 1] - h1, v[4 * j + 2] - h2, v[4 * j + 3] - h3);
      }
    } else {
      for (int j = 0; j < cnt; ++j) {
        float h = tf32_hi(v[j]);
        o[j] = h;
        ol[j] = v[j] - h;
      }
    }
  }
}

// bias (from the warp's shared-memory copy) + activation on 32 accumulator columns
__device__ __forceinline__ void bias_act32(const GemmKParams& p, float* v, const float* bias_s) {
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(bias_s + 4 * j);  // warp broadcast
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
  }
  if (p.act == DEVIT_ACT_GELU_ERF) {
    if (p.out_kind == DEVIT_OUT_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
  } else if (p.act == DEVIT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

// LayerNorm folding: v = rstd * acc + ((-rstd * mean) * c1 + c2), then the activation
__device__ __forceinline__ void ln_bias_act32(const GemmKParams& p, float* v, const float* cs,
                                              const float* bias_s, float rstd, float nmr) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 c = *reinterpret_cast<const float4*>(cs + 4 * j);      // warp broadcast
    const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * j);  // (zeros if no bias)
    v[4 * j] = fmaf(rstd, v[4 * j], fmaf(nmr, c.x, b.x));
    v[4 * j + 1] = fmaf(rstd, v[4 * j + 1], fmaf(nmr, c.y, b.y));
    v[4 * j + 2] = fmaf(rstd, v[4 * j + 2], fmaf(nmr, c.z, b.z));
    v[4 * j + 3] = fmaf(rstd, v[4 * j + 3], fmaf(nmr, c.w, b.w));
  }
  if (p.act == DEVIT_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
  }
}

template <int BN, int KIND, int CL, bool RES, bool BRES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
            const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
            const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
            const __grid_constant__ CUtensorMap tmR, const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN, CL, RES, BRES>;
  static_assert(!(RES && BRES), "the residual and B-resident variants are exclusive");
  constexpr int kElem = KIND == 0 ? 2 : 4;
  constexpr int kBlockK = 128 / kElem;  // elements per k-block (one swizzle atom)
  constexpr int kStages = Cfg::kStages;

  // Declared 1024-byte aligned (128B-swizzle atoms) and used directly: pointer arithmetic
  // through uintptr_t would make the compiler lose the shared address space and turn every
  // LDS/STS of the epilogue into a slower generic LD/ST.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* epi_smem = smem + Cfg::kRingBytes;
  uint8_t* bres_smem = smem + kStages * Cfg::kStageBytes;  // resident weight k-blocks (BRES)
  float* bias_smem = reinterpret_cast<float*>(epi_smem + Cfg::kEpiBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes + Cfg::kBiasBytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* rfull_bar = empty_bar + 8;  // [8 warps][4 slots]: residual chunk has landed
  uint64_t* bres_full = rfull_bar;      // BRES (never together with RES): resident B has landed
  uint64_t* tmem_full = rfull_bar + 32;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  // Role dispatch must be WARP-UNIFORM in the compiler's eyes (shfl from lane 0): the async
  // instructions (UTMALDG / UTCHMMA / UTCBAR / UTMASTG) take uniform-register operands, and in
  // code the compiler considers divergent every one of them is wrapped in an elect+broadcast
  // "waterfall" loop that costs more than the instruction itself.  So whole warps run the role
  // loops and a single elected lane issues.
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cta_rank =
      CL > 1 ? __shfl_sync(0xffffffffu, static_cast<int>(cluster_ctarank()), 0) : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;
  const bool leader = cta_rank == 0;

  griddep_launch_dependents();  // the next kernel's prologue may overlap this kernel
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (KIND == 1) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 32; ++s) mbar_init(&rfull_bar[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps * CL);  // one arrival per epilogue warp (of the pair)
    }
    if (p.tma_epi) {
      tma_prefetch_desc(&tmO0);
      if (p.resid) tma_prefetch_desc(&tmR);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CL == 1) {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    } else {
      tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_cg2();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int total_kb = p.total_kb;
  griddep_wait();  // operands / residual come from (and outputs may alias buffers of) earlier kernels

  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_units = ((num_m + CL - 1) / CL) * num_n;  // unit = CL m-blocks x one n-tile
  // This cluster's units: unit_first, unit_first + unit_step, ...  Ordinary mode: round-robin
  // over all units (n fastest, so co-resident CTAs share A rows in L2).  BRES: the cluster owns
  // n-tile (cluster_id % num_n) and every G-th m-block of it (G = clusters on that n-tile).
  int unit_first = cluster_id, unit_step = num_clusters;
  if (BRES) {
    const int n_fixed = cluster_id % num_n;
    const int group = (num_clusters - n_fixed + num_n - 1) / num_n;
    unit_first = (cluster_id / num_n) * num_n + n_fixed;
    unit_step = group * num_n;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      int stage = 0;
      uint32_t phase = 0;
      int tr_p = 0;
      if (BRES && unit_first < num_units) {
        // the n-tile's weight k-blocks, once: they stay in shared memory for the whole kernel
        const int n0 = (unit_first % num_n) * BN;
        const int n_valid = (p.N - n0) < BN ? (p.N - n0) : BN;
        const int n_cur = n_valid >= BN ? BN : ((n_valid + 16 * CL - 1) & ~(16 * CL - 1));
        const KSeg sg = p.segs[0];
        if (elect_one()) {
          if (CL == 1) {
            mbar_expect_tx(bres_full, sg.k_blocks * Cfg::kStageB);
            for (int kb = 0; kb < sg.k_blocks; ++kb)
              tma_load_2d(bres_smem + kb * Cfg::kStageB, &tmB0, bres_full,
                          sg.b_k_off + kb * kBlockK, n0);
          } else {
            const uint32_t bar = mapa_u32(smem_u32(bres_full), 0);
            if (leader) mbar_expect_tx(bres_full, 2 * sg.k_blocks * Cfg::kStageB);
            for (int kb = 0; kb < sg.k_blocks; ++kb)
              tma_load_2d_cg2(bres_smem + kb * Cfg::kStageB, &tmB0, bar,
                              sg.b_k_off + kb * kBlockK, n0 + cta_rank * (n_cur / 2));
          }
        }
      }
      for (int unit = unit_first; unit < num_units; unit += unit_step) {
        const int m0 = ((unit / num_n) * CL + cta_rank) * kBlockM;
        const int n0 = (unit % num_n) * BN;
        const int n_valid = (p.N - n0) < BN ? (p.N - n0) : BN;
        const int n_cur = n_valid >= BN ? BN : ((n_valid + 16 * CL - 1) & ~(16 * CL - 1));
        for (int s = 0; s < p.num_segs; ++s) {
          const KSeg sg = p.segs[s];
          const CUtensorMap* ma = (KIND == 1 && sg.a_plane) ? &tmA1 : &tmA0;
          const CUtensorMap* mb = (KIND == 1 && sg.b_plane) ? &tmB1 : &tmB0;
          for (int kb = 0; kb < sg.k_blocks; ++kb) {
            DEVIT_TRACE(0, tr_p);
            mbar_wait_warp(&empty_bar[stage], phase ^ 1);
            DEVIT_TRACE(1, tr_p);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kStageA;
            if (elect_one()) {
              if (CL == 1) {
                mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                tma_load_2d(sa, ma, &full_bar[stage], sg.a_k_off + kb * kBlockK,
                            sg.a_row_off + m0);
                if (!BRES) tma_load_2d(sb, mb, &full_bar[stage], sg.b_k_off + kb * kBlockK, n0);
              } else {
                // both CTAs' bytes complete on the LEADER's barrier (it issues the MMAs)
                const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                tma_load_2d_cg2(sa, ma, full_leader, sg.a_k_off + kb * kBlockK,
                                sg.a_row_off + m0);
                if (!BRES)
                  tma_load_2d_cg2(sb, mb, full_leader, sg.b_k_off + kb * kBlockK,
                                  n0 + cta_rank * (n_cur / 2));
              }
            }
            DEVIT_TRACE(2, tr_p);
            ++tr_p;
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (leader) {
      int stage = 0;
      uint32_t kphase_bits = 0;  // per-slot parity of full_bar: it completes on K-uses only
      int acc = 0;
      uint32_t acc_phase = 0;
      int tr_m = 0, tr_t = 0;
      if (BRES && unit_first < num_units) {
        mbar_wait_warp(bres_full, 0);  // the resident weight k-blocks (of both CTAs of a pair)
        tc_fence_after();
      }
      for (int unit = unit_first; unit < num_units; unit += unit_step) {
        const int n0 = (unit % num_n) * BN;
        const int n_valid = (p.N - n0) < BN ? (p.N - n0) : BN;
        const int n_cur = n_valid >= BN ? BN : ((n_valid + 16 * CL - 1) & ~(16 * CL - 1));
        const uint32_t idesc =
            make_idesc(KIND == 0 ? kFmtBF16 : kFmtTF32, kBlockM * CL, n_cur, 0, 0);
        DEVIT_TRACE(7, tr_t);
        if (CL == 1) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        else mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
        __syncwarp();
        tc_fence_after();
        DEVIT_TRACE(8, tr_t);
        ++tr_t;
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < total_kb; ++kb) {
          DEVIT_TRACE(3, tr_m);
          mbar_wait_warp(&full_bar[stage], (kphase_bits >> stage) & 1u);
          kphase_bits ^= 1u << stage;
          tc_fence_after();
          DEVIT_TRACE(4, tr_m);
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = BRES ? smem_u32(bres_smem + kb * Cfg::kStageB) : sa + Cfg::kStageA;
          const uint64_t da = make_sw128_desc(sa, 1024, 16);
          const uint64_t db = make_sw128_desc(sb, 1024, 16);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              if (CL == 1) {
                if (KIND == 0) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, accumulate | k);
                else umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, accumulate | k);
              } else {
                if (KIND == 0)
                  umma_bf16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, accumulate | k);
                else
                  umma_tf32_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, accumulate | k);
              }
            }
            // slot reusable (in both CTAs of a pair) once these MMAs retire
            if (CL == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_cg2(&empty_bar[stage], 3);
          }
          accumulate = 1;
          DEVIT_TRACE(6, tr_m);
          ++tr_m;
          if (++stage == kStages) stage = 0;
        }
        // accumulator complete -> epilogue warps (of both CTAs of a pair)
        if (elect_one()) {
          if (CL == 1) umma_commit(&tmem_full[acc]);
          else umma_commit_cg2(&tmem_full[acc], 3);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;       // 0..7
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32)
    const int half = ew >> 2;      // this warp handles column chunks c with (c & 1) == half
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t rphase_bits = 0;      // parity of this warp's two residual-buffer barriers
    uint8_t* stg = epi_smem + ew * 4096;  // (non-residual variants)
    float* bias_s = bias_smem + ew * BN;
    float* cs_s = bias_smem + (kEpiWarps + ew) * BN;  // LayerNorm-folding column sums c1
    int tr_e = 0;
    // ---- residual variant: slot bookkeeping (see the RES epilogue below)
    constexpr int kResSlots = BN / 64;  // this warp's chunks: c = half * kResSlots + j
    uint8_t* res_slots = epi_smem + ew * (kResSlots * 4096);
    uint8_t* xb_stg = epi_smem + kBlockM * BN * 4 + ew * 4096;  // bf16 copy staging (RES)
    uint64_t* res_bar = rfull_bar + ew * 4;
    int cur_cnt = 0, nxt_cnt = 0, nxt_row0 = 0, nxt_n0 = 0;
    // number of this warp's chunks (= slots used) in tile u, and the tile's coordinates
    auto res_tile = [&](int u, int* row0_, int* n0_) -> int {
      *n0_ = (u % num_n) * BN;
      *row0_ = ((u / num_n) * CL + cta_rank) * kBlockM + quarter * 32;
      if (*row0_ >= p.M) return 0;
      const int nv = (p.N - *n0_) < BN ? (p.N - *n0_) : BN;
      const int nch = (nv + 31) >> 5;
      const int mine = nch - half * kResSlots;
      return mine < 0 ? 0 : (mine > kResSlots ? kResSlots : mine);
    };
    // row of the residual tensor that holds the residual of output row r (a 32-row box never
    // wraps: periodic tables repeat their first 31 rows at the end)
    auto res_row = [&](int r) -> int { return p.resid_period > 0 ? r % p.resid_period : r; };
    auto res_issue = [&](int j, int row0_, int n0_) {
      if (elect_one()) {
        mbar_expect_tx(&res_bar[j], 4096);
        tma_load_2d(res_slots + j * 4096, &tmR, &res_bar[j], n0_ + (half * kResSlots + j) * 32,
                    res_row(row0_));
      }
    };
    // ---- per-tile epilogue constants, requested one tile ahead into registers
    float pf_bias[BN / 32], pf_cs[BN / 32];
    float2 pf_st[kMaxLnParts];
    auto epi_prefetch = [&](int u) {
      const int n0_ = (u % num_n) * BN;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
          const int col = n0_ + i * 32 + lane;
          pf_bias[i] = col < p.N ? __ldg(p.bias + col) : 0.f;
        }
      }
      if (!RES && p.ln_stats) {
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
          const int col = n0_ + i * 32 + lane;
          pf_cs[i] = col < p.N ? __ldg(p.ln_colsum + col) : 0.f;
        }
        const int m = ((u / num_n) * CL + cta_rank) * kBlockM + quarter * 32 + lane;
#pragma unroll
        for (int q = 0; q < kMaxLnParts; ++q)
          pf_st[q] = (q < p.ln_parts && m < p.M)
                         ? __ldg(reinterpret_cast<const float2*>(p.ln_stats) +
                                 static_cast<long long>(q) * p.M + m)
                         : make_float2(0.f, 0.f);
      }
    };
    if (p.tma_epi && unit_first < num_units) epi_prefetch(unit_first);
    if constexpr (RES) {
      if (unit_first < num_units) {
        cur_cnt = res_tile(unit_first, &nxt_row0, &nxt_n0);
        for (int j = 0; j < cur_cnt; ++j) res_issue(j, nxt_row0, nxt_n0);
      }
    }
    for (int unit = unit_first; unit < num_units; unit += unit_step) {
      const int m0 = ((unit / num_n) * CL + cta_rank) * kBlockM;
      const int n0 = (unit % num_n) * BN;
      const int row0 = m0 + quarter * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             acc * Cfg::kAccStride;
      const int n_valid = (p.N - n0) < BN ? (p.N - n0) : BN;
      const bool live = row0 < p.M;  // warp-uniform
      if (p.tma_epi) {
        // The tile's bias (and, LayerNorm folding, column sums + this lane's row statistics)
        // were requested one tile ago (pf_* registers): store them to this warp's shared-memory
        // copies, then request the next tile's so their latency never sits in front of the
        // accumulator read.
        float ln_rstd = 1.f, ln_nmr = 0.f;  // v = rstd * acc + (-rstd * mean) * c1 + c2
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < BN / 32; ++i) bias_s[i * 32 + lane] = pf_bias[i];
        }
        if (!RES && p.ln_stats) {
#pragma unroll
          for (int i = 0; i < BN / 32; ++i) cs_s[i * 32 + lane] = pf_cs[i];
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int q = 0; q < kMaxLnParts; ++q) {
            s1 += pf_st[q].x;
            s2 += pf_st[q].y;
          }
          const float mean = s1 * p.ln_inv_dim;
          const float var = fmaxf(fmaf(s2, p.ln_inv_dim, -mean * mean), 0.f);
          ln_rstd = rsqrtf(var + p.ln_eps);
          ln_nmr = -ln_rstd * mean;
        }
        if (unit + unit_step < num_units) epi_prefetch(unit + unit_step);
        if constexpr (RES) {
          // residual chunks of the NEXT tile whose slots this tile does not use can fly now
          nxt_cnt = 0;
          if (unit + unit_step < num_units)
            nxt_cnt = res_tile(unit + unit_step, &nxt_row0, &nxt_n0);
          for (int j = cur_cnt; j < nxt_cnt; ++j) res_issue(j, nxt_row0, nxt_n0);
        }
        if (ew == 0 && lane == 0) DEVIT_TRACE(9, tr_e);
        mbar_wait_warp(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (ew == 0 && lane == 0) DEVIT_TRACE(10, tr_e);
        if (p.out_kind == DEVIT_OUT_BF16) {
          // ---- bf16 output: 64 columns (128 B per row) per chunk
          const int nchunk = (n_valid + 63) >> 6;
          if (live && !(p.dbg & 1)) {
#pragma unroll 1
            for (int c = half; c < nchunk; c += 2) {
              uint32_t r0[32], r1[32];
              const bool tr = ew == 0 && lane == 0 && c == 0;
              if (tr) DEVIT_TRACE(12, tr_e);
              tmem_ld_x32(t_row + c * 64, r0);
              tmem_ld_x32(t_row + c * 64 + 32, r1);
              tmem_ld_wait();
              if (tr) DEVIT_TRACE(13, tr_e);
              float* v0 = reinterpret_cast<float*>(r0);
              float* v1 = reinterpret_cast<float*>(r1);
              if (p.ln_stats) {
                ln_bias_act32(p, v0, cs_s + c * 64, bias_s + c * 64, ln_rstd, ln_nmr);
                ln_bias_act32(p, v1, cs_s + c * 64 + 32, bias_s + c * 64 + 32, ln_rstd, ln_nmr);
              } else {
                bias_act32(p, v0, bias_s + c * 64);
                bias_act32(p, v1, bias_s + c * 64 + 32);
              }
              if (p.alpha != 1.0f) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { v0[j] *= p.alpha; v1[j] *= p.alpha; }
              }
              if (tr) DEVIT_TRACE(14, tr_e);
              if (elect_one()) bulk_wait_read<0>();  // previous store of this warp has drained
              __syncwarp();
              if (tr) DEVIT_TRACE(15, tr_e);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 t;
                t.x = pack_bf16x2(v0[8 * g], v0[8 * g + 1]);
                t.y = pack_bf16x2(v0[8 * g + 2], v0[8 * g + 3]);
                t.z = pack_bf16x2(v0[8 * g + 4], v0[8 * g + 5]);
                t.w = pack_bf16x2(v0[8 * g + 6], v0[8 * g + 7]);
                *reinterpret_cast<uint4*>(stg + stg_off(lane, g)) = t;
                t.x = pack_bf16x2(v1[8 * g], v1[8 * g + 1]);
                t.y = pack_bf16x2(v1[8 * g + 2], v1[8 * g + 3]);
                t.z = pack_bf16x2(v1[8 * g + 4], v1[8 * g + 5]);
                t.w = pack_bf16x2(v1[8 * g + 6], v1[8 * g + 7]);
                *reinterpret_cast<uint4*>(stg + stg_off(lane, 4 + g)) = t;
              }
              if (tr) DEVIT_TRACE(16, tr_e);
              fence_proxy_async_smem();
              __syncwarp();
              if (tr) DEVIT_TRACE(17, tr_e);
              if (elect_one()) {
                tma_store_2d(&tmO0, stg, n0 + c * 64, row0);
                bulk_commit();
              }
              if (tr) DEVIT_TRACE(18, tr_e);
            }
          }
        } else {
          // ---- fp32 output, 32 columns (128 B per row) per chunk
          const bool split = p.out_kind == DEVIT_OUT_F32_SPLIT;
          const int nchunk = (n_valid + 31) >> 5;
          if constexpr (RES) {
            // + fp32 residual (x += ...).  Warp (quarter, half) owns the kResSlots consecutive
            // chunks c = half * kResSlots + j of its 32 rows; slot j holds the residual of chunk
            // c, TMA-prefetched one whole tile ahead.  The sum is written back into the slot and
            // TMA-stored from there; as soon as that store has read the slot, the same chunk of
            // the next tile is requested into it.  Optionally (LayerNorm folding, producer side)
            // the warp also emits the bf16 copy of its 64 columns and their partial row sums.
            float st1 = 0.f, st2 = 0.f;
#pragma unroll 1
            for (int j = 0; j < cur_cnt; ++j) {
              const int c = half * kResSlots + j;
              uint32_t r[32];
              tmem_ld_x32(t_row + c * 32, r);
              tmem_ld_wait();
              float* v = reinterpret_cast<float*>(r);
              bias_act32(p, v, bias_s + c * 32);
              mbar_wait_warp(&res_bar[j], (rphase_bits >> j) & 1u);
              rphase_bits ^= 1u << j;
              uint8_t* b = res_slots + j * 4096;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 t = *reinterpret_cast<const float4*>(b + stg_off(lane, g));
                v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
              }
              if (p.alpha != 1.0f) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] *= p.alpha;
              }
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(b + stg_off(lane, g)) =
                    make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
              if (p.xb_out) {  // BN == 128: j in {0, 1} -> 16-byte groups 4j .. 4j+3 of the row
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  uint4 t;
                  t.x = pack_bf16x2(v[8 * g], v[8 * g + 1]);
                  t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                  t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
                  t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                  *reinterpret_cast<uint4*>(xb_stg + stg_off(lane, 4 * j + g)) = t;
                }
              }
              if (p.stats_out) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                  st1 += v[k];
                  st2 = fmaf(v[k], v[k], st2);
                }
              }
              fence_proxy_async_smem();
              if (elect_one()) {
                tma_store_2d(&tmO0, b, n0 + c * 32, row0);
                if (p.xb_out && j == kResSlots - 1)
                  tma_store_2d(&tmO1, xb_stg, n0 + half * 64, row0);
                bulk_commit();
                if (j > 0) {
                  bulk_wait_read<1>();  // the previous chunk's store has drained its slot
                  if (j - 1 < nxt_cnt) {
                    mbar_expect_tx(&res_bar[j - 1], 4096);
                    tma_load_2d(b - 4096, &tmR, &res_bar[j - 1],
                                nxt_n0 + (half * kResSlots + j - 1) * 32, res_row(nxt_row0));
                  }
                }
              }
              __syncwarp();
            }
            if (cur_cnt > 0 && elect_one()) {
              bulk_wait_read<0>();
              if (cur_cnt - 1 < nxt_cnt) {
                const int j = cur_cnt - 1;
                mbar_expect_tx(&res_bar[j], 4096);
                tma_load_2d(res_slots + j * 4096, &tmR, &res_bar[j],
                            nxt_n0 + (half * kResSlots + j) * 32, res_row(nxt_row0));
              }
            }
            if (p.stats_out && cur_cnt > 0 && row0 + lane < p.M) {
              // part index = 2 * (n0 / 128) + half   (BN == 128 when stats are requested)
              float2* so = reinterpret_cast<float2*>(p.stats_out) +
                           static_cast<long long>(2 * (n0 / BN) + half) * p.M + row0 + lane;
              *so = make_float2(st1, st2);
            }
            cur_cnt = nxt_cnt;
          } else if (live) {
#pragma unroll 1
            for (int c = half; c < nchunk; c += 2) {
              uint32_t r[32];
              tmem_ld_x32(t_row + c * 32, r);
              tmem_ld_wait();
              float* v = reinterpret_cast<float*>(r);
              bias_act32(p, v, bias_s + c * 32);
              if (p.alpha != 1.0f) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
              }
              if (elect_one()) bulk_wait_read<0>();  // this warp's previous store has drained
              __syncwarp();
              if (!split) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  *reinterpret_cast<float4*>(stg + stg_off(lane, g)) =
                      make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                  tma_store_2d(&tmO0, stg, n0 + c * 32, row0);
                  bulk_commit();
                }
              } else {  // hi plane, then lo plane through the same buffer (parity mode)
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  *reinterpret_cast<float4*>(stg + stg_off(lane, g)) =
                      make_float4(tf32_hi(v[4 * g]), tf32_hi(v[4 * g + 1]),
                                  tf32_hi(v[4 * g + 2]), tf32_hi(v[4 * g + 3]));
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                  tma_store_2d(&tmO0, stg, n0 + c * 32, row0);
                  bulk_commit();
                  bulk_wait_read<0>();
                }
                __syncwarp();
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  *reinterpret_cast<float4*>(stg + stg_off(lane, g)) = make_float4(
                      v[4 * g] - tf32_hi(v[4 * g]), v[4 * g + 1] - tf32_hi(v[4 * g + 1]),
                      v[4 * g + 2] - tf32_hi(v[4 * g + 2]), v[4 * g + 3] - tf32_hi(v[4 * g + 3]));
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                  tma_store_2d(&tmO1, stg, n0 + c * 32, row0);
                  bulk_commit();
                }
              }
            }
          }
        }
      } else {
        // ---- direct path (row-remapped patch embedding, unaligned / tiny outputs)
        mbar_wait_warp(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const int m = row0 + lane;
        const bool row_ok = m < p.M;
        long long row_out = m;
        int rb_row = 0;
        if (p.rowmap_period > 0) {
          const int q = m / p.rowmap_period;
          const int r = m - q * p.rowmap_period;
          rb_row = p.rowmap_off + r;
          row_out = (long long)q * p.rowmap_stride + rb_row;
        }
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;
          uint32_t r[32];
          __syncwarp();
          tmem_ld_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            const int cnt = p.N - col0;
            if (cnt >= 32 && p.vec_ok)
              epilogue_store<true>(p, v, 32, row_out, rb_row, col0);
            else
              epilogue_store<false>(p, v, cnt < 32 ? cnt : 32, row_out, rb_row, col0);
          }
        }
      }
      if (ew == 0 && lane == 0) {
        DEVIT_TRACE(11, tr_e);
        ++tr_e;
      }
      // every lane's TMEM reads are complete (tcgen05.wait::ld) and ordered before the warp
      // barrier; one lane then releases the accumulator (a remote arrive per THREAD made the
      // pair's hand-off the bottleneck: 256 serialized cluster transactions per tile)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1 || leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.tma_epi && elect_one()) bulk_wait_all<0>();  // stores complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody leaves while the peer may still signal / be written
  if (warp == 1) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int KIND, int CL, bool RES = false, bool BRES = false>
static int launch_gemm(const CUtensorMap* tm, const GemmKParams& p, cudaStream_t stream,
                       int tag) {
  using Cfg = GemmCfg<BN, CL, RES, BRES>;
  static bool attr_done[64] = {};  // per device; benign race: the attribute set is idempotent
  int dev = 0;
  DEVIT_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    DEVIT_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, KIND, CL, RES, BRES>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_done[dev & 63] = true;
  }
  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_units = ((num_m + CL - 1) / CL) * ((p.N + BN - 1) / BN);
  int clusters = num_sms() / CL;
  if (clusters > num_units) clusters = num_units;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  {
    ProfScope ps(tag, stream);
    DEVIT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, KIND, CL, RES, BRES>, tm[0], tm[1], tm[2], tm[3],
                                     tm[4], tm[5], tm[6], p));
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

template <int BN, int KIND, bool RES = false>
static int launch_gemm_cl(int cl, const CUtensorMap* tm, const GemmKParams& p,
                          cudaStream_t stream, int tag) {
  if (cl == 2) return launch_gemm<BN, KIND, 2, RES>(tm, p, stream, tag);
  return launch_gemm<BN, KIND, 1, RES>(tm, p, stream, tag);
}

// The large-M GEMMs of this model are bound by the bytes each SM pulls through L2 per tile:
// 128 rows of A plus BN / cl rows of B per k-block.  Pick the tile width that minimises that
// (ragged last tiles are cheap: TMA clips them and the UMMA is issued narrower).
static int pick_block_n(int n, int cl) {
  int best = 128, best_cost = 1 << 30;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int cost = ((n + bn - 1) / bn) * (kBlockM + bn / cl);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace devit

// Debug: device buffer of 20 x 512 int64 receiving clock64 stamps of CTA 0 (NULL disables).
namespace devit { extern long long* g_attn_trace; }
extern "C" int devit_debug_set_trace(long long* device_buf) {
  devit::g_trace = device_buf;
  devit::g_attn_trace = device_buf;
  return DEVIT_OK;
}

extern "C" int devit_gemm(const devit_gemm_args* a, void* stream_v) {
  using namespace devit;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  DEVIT_REQUIRE(a != nullptr, "devit_gemm: null args");
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(a->precision == DEVIT_BF16 || a->precision == DEVIT_FP32,
                "devit_gemm: bad precision %d", a->precision);
  DEVIT_REQUIRE(a->m > 0 && a->n > 0, "devit_gemm: empty problem m=%d n=%d", a->m, a->n);
  DEVIT_REQUIRE(a->a && a->b && a->out, "devit_gemm: null operand");
  DEVIT_REQUIRE(a->num_segs >= 1 && a->num_segs <= 8, "devit_gemm: num_segs %d not in [1,8]",
                a->num_segs);
  DEVIT_REQUIRE(a->out_kind >= DEVIT_OUT_BF16 && a->out_kind <= DEVIT_OUT_F32_SPLIT,
                "devit_gemm: bad out_kind %d", a->out_kind);
  DEVIT_REQUIRE(!(a->resid && a->out_kind == DEVIT_OUT_BF16 &&
                  static_cast<const void*>(a->resid) == a->out),
                "devit_gemm: resid may alias out only for fp32 outputs");
  const int kind = a->precision == DEVIT_BF16 ? 0 : 1;
  const int elem = kind == 0 ? 2 : 4;
  const int block_k = 128 / elem;

  GemmKParams p;
  p.M = a->m;
  p.N = a->n;
  p.num_segs = 0;
  p.total_kb = 0;
  for (int s = 0; s < a->num_segs; ++s) {
    const devit_gemm_seg& g = a->segs[s];
    DEVIT_REQUIRE(g.k_len > 0 && g.a_k_off >= 0 && g.b_k_off >= 0 && g.a_row_off >= 0,
                  "devit_gemm: bad segment %d", s);
    DEVIT_REQUIRE(g.a_k_off % (16 / elem) == 0 && g.b_k_off % (16 / elem) == 0,
                  "devit_gemm: segment %d K offsets must be 16-byte aligned", s);
    const bool ragged = g.k_len % block_k != 0;
    DEVIT_REQUIRE(!ragged || (g.a_k_off + g.k_len == a->a_cols && g.b_k_off + g.k_len == a->b_cols),
                  "devit_gemm: segment %d has a ragged K (%d) that does not end at the operand "
                  "edge",
                  s, g.k_len);
    const int kbs = (g.k_len + block_k - 1) / block_k;
    if (kind == 0) {
      p.segs[p.num_segs++] = KSeg{g.a_row_off, g.a_k_off, g.b_k_off, kbs, 0, 0};
      p.total_kb += kbs;
    } else {  // 3xTF32: small cross terms first, then hi*hi
      p.segs[p.num_segs++] = KSeg{g.a_row_off, g.a_k_off, g.b_k_off, kbs, 1, 0};
      p.segs[p.num_segs++] = KSeg{g.a_row_off, g.a_k_off, g.b_k_off, kbs, 0, 1};
      p.segs[p.num_segs++] = KSeg{g.a_row_off, g.a_k_off, g.b_k_off, kbs, 0, 0};
      p.total_kb += 3 * kbs;
    }
  }
  p.out = a->out;
  p.ldo = a->ldo;
  p.out_kind = a->out_kind;
  p.out_plane_stride = a->out_plane_stride;
  p.bias = a->bias;
  p.resid = a->resid;
  p.resid_period = a->resid_period;
  DEVIT_REQUIRE(a->resid_period >= 0 && !(a->resid_period > 0 && !a->resid),
                "devit_gemm: resid_period needs resid");
  DEVIT_REQUIRE(!(a->resid_period > 0 && static_cast<const void*>(a->resid) == a->out),
                "devit_gemm: a periodic residual table cannot alias the output");
  p.ldr = a->ldr;
  p.rowbias = a->rowbias;
  p.ld_rowbias = a->ld_rowbias;
  p.act = a->act;
  p.alpha = a->alpha;
  p.rowmap_period = a->rowmap_period;
  p.rowmap_stride = a->rowmap_stride;
  p.rowmap_off = a->rowmap_off;
  DEVIT_REQUIRE(!(p.rowbias && p.rowmap_period <= 0),
                "devit_gemm: rowbias needs rowmap_period > 0");
  // ---- LayerNorm folding
  const bool ln_consumer = a->ln_stats != nullptr;
  const bool ln_producer = a->out_bf16 != nullptr || a->stats_out != nullptr;
  if (ln_consumer) {
    DEVIT_REQUIRE(kind == 0 && a->out_kind == DEVIT_OUT_BF16 && !a->resid,
                  "devit_gemm: ln_stats needs bf16 operands and a bf16 output without residual");
    DEVIT_REQUIRE(a->ln_colsum && a->bias && a->ln_parts >= 1 && a->ln_parts <= kMaxLnParts &&
                      a->ln_dim > 0,
                  "devit_gemm: ln_stats needs ln_colsum, bias, 1 <= ln_parts <= %d and ln_dim > 0",
                  kMaxLnParts);
    DEVIT_REQUIRE(reinterpret_cast<uintptr_t>(a->ln_stats) % 8 == 0,
                  "devit_gemm: ln_stats must be 8-byte aligned");
  }
  if (ln_producer) {
    DEVIT_REQUIRE(a->out_kind == DEVIT_OUT_F32 && a->resid && a->n % 128 == 0 &&
                      a->rowmap_period <= 0,
                  "devit_gemm: out_bf16 / stats_out need an fp32 residual output with n %% 128 == 0");
    DEVIT_REQUIRE(!a->out_bf16 || (reinterpret_cast<uintptr_t>(a->out_bf16) % 16 == 0 &&
                                   (a->ld_out_bf16 * 2) % 16 == 0),
                  "devit_gemm: out_bf16 must be 16-byte aligned with a 16-byte row pitch");
    DEVIT_REQUIRE(!a->stats_out || reinterpret_cast<uintptr_t>(a->stats_out) % 8 == 0,
                  "devit_gemm: stats_out must be 8-byte aligned");
  }
  p.ln_stats = a->ln_stats;
  p.ln_parts = a->ln_parts;
  p.ln_inv_dim = ln_consumer ? 1.0f / static_cast<float>(a->ln_dim) : 0.f;
  p.ln_eps = a->ln_eps;
  p.ln_colsum = a->ln_colsum;
  p.stats_out = a->stats_out;
  p.xb_out = a->out_bf16 ? 1 : 0;
  // 16-byte vector requirements of the direct path
  const int out_elem = a->out_kind == DEVIT_OUT_BF16 ? 2 : 4;
  bool vec = (reinterpret_cast<uintptr_t>(a->out) % 16 == 0) && ((a->ldo * out_elem) % 16 == 0);
  if (a->out_kind == DEVIT_OUT_F32_SPLIT) vec = vec && ((a->out_plane_stride * 4) % 16 == 0);
  if (a->bias) vec = vec && (reinterpret_cast<uintptr_t>(a->bias) % 16 == 0);
  if (a->resid)
    vec = vec && (reinterpret_cast<uintptr_t>(a->resid) % 16 == 0) && ((a->ldr * 4) % 16 == 0);
  if (a->rowbias)
    vec = vec && (reinterpret_cast<uintptr_t>(a->rowbias) % 16 == 0) &&
          ((a->ld_rowbias * 4) % 16 == 0);
  p.vec_ok = vec ? 1 : 0;

  const int tag = (a->profile_tag >= 0 && a->profile_tag < 8) ? a->profile_tag : 0;
  // CTA pairs (cta_group::2, 256-row tiles) pay off once there are enough m-blocks
  // CTA pairs halve the weight-tile bytes each SM pulls through L2 (the binding resource at
  // K = 384); measured on B200 they win for every large-M shape except the short-K residual
  // GEMM (proj); sweeps in profiles/r1_gemm_sweep_*.txt, pipeline traces in profiles/r1_trace_qkv_*.txt
  int cl = a->cluster_m;
  const int total_k = [&] { int t = 0; for (int s = 0; s < a->num_segs; ++s) t += a->segs[s].k_len; return t; }();
  if (cl == 0) {
    cl = (a->m >= 64 * kBlockM) ? 2 : 1;
    if (a->resid && total_k < 512) cl = 1;
    static int env_cl = kEnvUnread;
    if (env_int("DEVIT_GEMM_CLUSTER", 0, &env_cl)) cl = env_cl;
  }
  DEVIT_REQUIRE(cl == 1 || cl == 2, "devit_gemm: cluster_m %d unsupported", cl);
  int bn = a->block_n ? a->block_n : pick_block_n(a->n, cl);
  if (!a->block_n) {
    static int env_bn = kEnvUnread;
    if (env_int("DEVIT_GEMM_BN", 0, &env_bn)) bn = env_bn;
  }
  DEVIT_REQUIRE(bn == 128 || bn == 192 || bn == 256, "devit_gemm: block_n %d unsupported", bn);

  // ---- coalesced (TMA) epilogue eligibility
  bool tma_epi = p.rowmap_period <= 0 && !p.rowbias &&
                 (reinterpret_cast<uintptr_t>(a->out) % 16 == 0) && ((a->ldo * out_elem) % 16 == 0);
  if (a->out_kind == DEVIT_OUT_BF16) tma_epi = tma_epi && !a->resid;
  if (a->out_kind == DEVIT_OUT_F32_SPLIT)
    tma_epi = tma_epi && !a->resid && ((a->out_plane_stride * 4) % 16 == 0);
  if (a->resid)
    tma_epi = tma_epi && (reinterpret_cast<uintptr_t>(a->resid) % 16 == 0) &&
              ((a->ldr * 4) % 16 == 0);
  {
    static int env_no_tma = kEnvUnread;  // debug: 1 = never, 2 = not for residual GEMMs
    const int v = env_int("DEVIT_GEMM_NO_TMA_EPI", 0, &env_no_tma);
    if (v == 1 || (v == 2 && a->resid)) tma_epi = false;
  }
  DEVIT_REQUIRE(!(ln_consumer || ln_producer) || tma_epi,
                "devit_gemm: LayerNorm folding needs 16-byte aligned outputs (TMA epilogue)");
  if (tma_epi && a->resid) {  // residual variant: BN <= 192
    if (!a->block_n && cl == 2 && a->n % 128 == 0) bn = 128;
    if (bn == 256) bn = 192;
    if (ln_producer) bn = 128;  // one 64-column bf16 box + one stats part per epilogue warp
  }
  CUtensorMap ta0, ta1, tb0, tb1;
  rc = encode_tmap_2d(&ta0, a->a, elem, a->a_cols, a->a_rows, a->lda, block_k, kBlockM, false);
  if (rc) return rc;
  rc = encode_tmap_2d(&tb0, a->b, elem, a->b_cols, a->b_rows, a->ldb, block_k, bn / cl, true);
  if (rc) return rc;
  ta1 = ta0;
  tb1 = tb0;
  if (kind == 1) {
    DEVIT_REQUIRE(a->a_plane_stride > 0 && a->b_plane_stride > 0,
                  "devit_gemm: DEVIT_FP32 operands need plane strides");
    rc = encode_tmap_2d(&ta1, static_cast<const float*>(a->a) + a->a_plane_stride, elem,
                        a->a_cols, a->a_rows, a->lda, block_k, kBlockM, false);
    if (rc) return rc;
    rc = encode_tmap_2d(&tb1, static_cast<const float*>(a->b) + a->b_plane_stride, elem,
                        a->b_cols, a->b_rows, a->ldb, block_k, bn / cl, true);
    if (rc) return rc;
  }

  p.tma_epi = tma_epi ? 1 : 0;
  p.dbg = 0;
  p.trace = g_trace;
  static int env_dbg = kEnvUnread;
  p.dbg = env_int("DEVIT_GEMM_DBG", 0, &env_dbg);
  CUtensorMap tm[7];
  tm[0] = ta0; tm[1] = ta1; tm[2] = tb0; tm[3] = tb1;
  tm[4] = ta0; tm[5] = ta0; tm[6] = ta0;  // placeholders when the direct epilogue is used
  if (tma_epi) {
    const int box_cols = a->out_kind == DEVIT_OUT_BF16 ? 64 : 32;
    rc = encode_tmap_2d(&tm[4], a->out, out_elem, a->n, a->m, a->ldo, box_cols, 32, false);
    if (rc) return rc;
    tm[5] = tm[4];
    if (a->out_kind == DEVIT_OUT_F32_SPLIT) {
      rc = encode_tmap_2d(&tm[5], static_cast<float*>(a->out) + a->out_plane_stride, 4, a->n,
                          a->m, a->ldo, 32, 32, false);
      if (rc) return rc;
    }
    if (a->out_bf16) {
      rc = encode_tmap_2d(&tm[5], a->out_bf16, 2, a->n, a->m, a->ld_out_bf16, 64, 32, false);
      if (rc) return rc;
    }
    tm[6] = tm[4];
    if (a->resid) {
      const int r_rows = a->resid_period > 0 ? a->resid_period + 31 : a->m;
      rc = encode_tmap_2d(&tm[6], a->resid, 4, a->n, r_rows, a->ldr, 32, 32, false);
      if (rc) return rc;
    }
  }

  DEVIT_REQUIRE(a->resid_period == 0 || tma_epi,
                "devit_gemm: resid_period needs the TMA epilogue (aligned fp32 output, no rowmap)");
  if (tma_epi && a->resid) {
    // residual variant: whole-tile residual slots in shared memory (BN <= 192, single CTA)
    if (kind == 0) {
      if (bn == 128) return launch_gemm_cl<128, 0, true>(cl, tm, p, stream, tag);
      return launch_gemm_cl<192, 0, true>(cl, tm, p, stream, tag);
    }
    if (bn == 128) return launch_gemm_cl<128, 1, true>(cl, tm, p, stream, tag);
    return launch_gemm_cl<192, 1, true>(cl, tm, p, stream, tag);
  }
  // B-resident mode: one K segment of <= 6 k-blocks, CTA pairs, enough m-blocks per n-tile
  {
    static int env_bres = kEnvUnread;  // DEVIT_GEMM_BRES=0 switches it off (comparison)
    const int num_n = (a->n + bn - 1) / bn;
    const int num_mp = ((a->m + kBlockM - 1) / kBlockM + 1) / 2;
    const int clusters = num_sms() / 2;
    if (kind == 0 && cl == 2 && (bn == 192 || bn == 256) && p.num_segs == 1 &&
        p.total_kb <= kMaxResidentKb && !a->resid && tma_epi && num_n <= clusters &&
        num_mp >= 4 * clusters / num_n && env_int("DEVIT_GEMM_BRES", 1, &env_bres)) {
      if (bn == 192) return launch_gemm<192, 0, 2, false, true>(tm, p, stream, tag);
      return launch_gemm<256, 0, 2, false, true>(tm, p, stream, tag);
    }
  }
  if (kind == 0) {
    if (bn == 128) return launch_gemm_cl<128, 0>(cl, tm, p, stream, tag);
    if (bn == 192) return launch_gemm_cl<192, 0>(cl, tm, p, stream, tag);
    return launch_gemm_cl<256, 0>(cl, tm, p, stream, tag);
  } else {
    if (bn == 128) return launch_gemm_cl<128, 1>(cl, tm, p, stream, tag);
    if (bn == 192) return launch_gemm_cl<192, 1>(cl, tm, p, stream, tag);
    return launch_gemm_cl<256, 1>(cl, tm, p, stream, tag);
  }
}

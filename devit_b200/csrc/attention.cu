// Fused multi-head attention for ViT-sized sequences (tokens <= 256, head_dim 64), see
// include/devit_b200.h devit_attention.  Replaces models/de_vit.py:70-74.
//
// bf16 path (attn_bf16_kernel): one CTA per (query tile of 128 rows, kept head, image).
//   TMA pulls the Q tile and the head's whole K and V straight out of the QKV-GEMM output
//   (3-D tensor map [image][token][3*h*64]; rows past `tokens` are zero-filled), so no
//   head-major copy exists anywhere.  S = Q K^T is ONE accumulation chain of 4 UMMAs into
//   TMEM (128 x KVP fp32); because every key of the head fits in that tile the softmax is
//   single pass (no online rescale): 128 threads each own one TMEM lane = one query row,
//   read it twice (max, then exp2/sum), and write P as bf16 into shared memory in the
//   K-major 128B-swizzled layout the tensor core expects.  O = P V runs as KVP/16 UMMAs with
//   V consumed in place as an MN-major operand, accumulating over the dead S columns.
//   The epilogue scales by 1/rowsum and stores 128 contiguous bytes per row into
//   out[token, head*64 ..], which is exactly the A operand of the proj GEMM.
//   ~92 KB smem and 256 TMEM columns per CTA -> two CTAs per SM, so one CTA's softmax
//   overlaps the other's loads and MMAs.
// fp32 path (attn_f32_kernel): CUDA-core fp32 for the parity mode.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace devit {

constexpr int kAttnThreads = 160;  // 4 softmax warps + 1 control warp

template <int KVP>
struct AttnCfg {
  static constexpr int kAtoms = (KVP + 63) / 64;
  static constexpr int kQBytes = 128 * 128;
  static constexpr int kKVBytes = KVP * 128;
  static constexpr int kPBytes = kAtoms * 16384;
  static constexpr int kOffK = kQBytes;
  static constexpr int kQKEnd = kQBytes + kKVBytes;
  static constexpr int kOffV = ((kQKEnd > kPBytes ? kQKEnd : kPBytes) + 1023) / 1024 * 1024;
  static constexpr int kOffBar = kOffV + kKVBytes;
  static constexpr int kSmemBytes = kOffBar + 64 + 1024;
  static constexpr int kTmemCols = KVP <= 64 ? 64 : (KVP <= 128 ? 128 : 256);
};

template <int KVP>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_bf16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 __nv_bfloat16* __restrict__ out, int tokens, int heads, float scale_log2e) {
  using Cfg = AttnCfg<KVP>;
  // Declared 1024-byte aligned (128B-swizzle atoms) and used directly: pointer arithmetic
  // through uintptr_t would make the compiler lose the shared address space and turn every
  // LDS/STS of the epilogue into a slower generic LD/ST.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + Cfg::kOffK;
  uint8_t* sP = smem;  // overlays Q and K once S has been computed
  uint8_t* sV = smem + Cfg::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  // warp-uniform role id (see gemm.cu: keeps the async instructions free of waterfall loops)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int mtile = blockIdx.x;
  const int head = blockIdx.y;
  const int img = blockIdx.z;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 4) {
    {
      // ---- loads
      if (elect_one()) {
        mbar_expect_tx(bar_qk, Cfg::kQBytes + Cfg::kKVBytes);
        tma_load_3d(sQ, &tmQ, bar_qk, head * 64, mtile * 128, img);
        tma_load_3d(sK, &tmKV, bar_qk, (heads + head) * 64, 0, img);
        mbar_expect_tx(bar_v, Cfg::kKVBytes);
        tma_load_3d(sV, &tmKV, bar_v, (2 * heads + head) * 64, 0, img);
      }
      // ---- S = Q K^T   (M=128, N=KVP, K=64: 4 UMMAs of K=16)
      mbar_wait_warp(bar_qk, 0);
      tc_fence_after();
      {
        const uint32_t idesc = make_idesc(kFmtBF16, 128, KVP, 0, 0);
        const uint64_t dq = make_sw128_desc(smem_u32(sQ), 1024, 16);
        const uint64_t dk = make_sw128_desc(smem_u32(sK), 1024, 16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dq + 2 * k, dk + 2 * k, idesc, k > 0);
          umma_commit(bar_s);
        }
      }
      // ---- O = P V     (M=128, N=64, K=KVP: KVP/16 UMMAs; V is MN-major, 16 keys = 2 KB)
      mbar_wait_warp(bar_p, 0);
      tc_fence_after();
      mbar_wait_warp(bar_v, 0);
      tc_fence_after();
      {
        const uint32_t idesc = make_idesc(kFmtBF16, 128, 64, 0, 1);
        const uint32_t p0 = smem_u32(sP);
        const uint32_t v0 = smem_u32(sV);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KVP / 16; ++k) {
            const uint64_t dp = make_sw128_desc(p0 + (k >> 2) * 16384 + (k & 3) * 32, 1024, 16);
            const uint64_t dv = make_sw128_desc(v0 + k * 2048, 1024, 1024);
            umma_bf16(tmem_base, dp, dv, idesc, k > 0);
          }
          umma_commit(bar_o);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue warps
    const int row_in_tile = warp * 32 + lane;
    const int row = mtile * 128 + row_in_tile;
    const bool warp_live = (mtile * 128 + warp * 32) < tokens;  // any valid row in this warp
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr int kFull = KVP / 32;
    constexpr int kTail = KVP % 32;  // 0 or 16

    mbar_wait_warp(bar_s, 0);
    tc_fence_after();
    float inv_sum = 0.f;
    if (warp_live) {
      // pass 1: row max over the valid keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kFull; ++c) {
        uint32_t r[32];
        tmem_ld_x32(t_row + c * 32, r);
        tmem_ld_wait();
        if ((c + 1) * 32 <= tokens) {  // whole chunk valid: no per-element test
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      }
      if (kTail) {
        uint32_t r[16];
        tmem_ld_x16(t_row + kFull * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (kFull * 32 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      const float moff = mx * scale_log2e;
      // pass 2: p = exp2(s*c - max*c), row sum, P -> smem (bf16, K-major, 128B swizzle)
      float sum = 0.f;
      uint8_t* prow = sP + row_in_tile * 128;
      const int sw = row_in_tile & 7;
#pragma unroll 1
      for (int c = 0; c < kFull; ++c) {
        uint32_t r[32];
        tmem_ld_x32(t_row + c * 32, r);
        tmem_ld_wait();
        float pv[32];
        if ((c + 1) * 32 <= tokens) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2e, -moff));
            sum += e;
            pv[j] = e;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2e, -moff));
            e = (c * 32 + j < tokens) ? e : 0.f;
            sum += e;
            pv[j] = e;
          }
        }
        uint8_t* atom = prow + (c >> 1) * 16384;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 t;
          t.x = pack_bf16x2(pv[8 * g], pv[8 * g + 1]);
          t.y = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
          t.z = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
          t.w = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(atom + ((chunk ^ sw) << 4)) = t;
        }
      }
      if (kTail) {
        uint32_t r[16];
        tmem_ld_x16(t_row + kFull * 32, r);
        tmem_ld_wait();
        float pv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float e = fast_exp2(fmaf(__uint_as_float(r[j]), scale_log2e, -moff));
          e = (kFull * 32 + j < tokens) ? e : 0.f;
          sum += e;
          pv[j] = e;
        }
        uint8_t* atom = prow + (kFull >> 1) * 16384;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 t;
          t.x = pack_bf16x2(pv[8 * g], pv[8 * g + 1]);
          t.y = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
          t.z = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
          t.w = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
          const int chunk = (kFull & 1) * 4 + g;
          *reinterpret_cast<uint4*>(atom + ((chunk ^ sw) << 4)) = t;
        }
      }
      inv_sum = 1.0f / sum;
    }
    // P (generic-proxy writes) must be visible to the tensor core (async proxy); S reads done.
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait_warp(bar_o, 0);
    tc_fence_after();
    if (warp_live) {
      uint32_t r0[32], r1[32];
      tmem_ld_x32(t_row, r0);
      tmem_ld_x32(t_row + 32, r1);
      tmem_ld_wait();
      if (row < tokens) {
        __nv_bfloat16* o =
            out + (static_cast<long long>(img) * tokens + row) * (heads * 64) + head * 64;
        uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 t;
          t.x = pack_bf16x2(__uint_as_float(r0[8 * g]) * inv_sum, __uint_as_float(r0[8 * g + 1]) * inv_sum);
          t.y = pack_bf16x2(__uint_as_float(r0[8 * g + 2]) * inv_sum, __uint_as_float(r0[8 * g + 3]) * inv_sum);
          t.z = pack_bf16x2(__uint_as_float(r0[8 * g + 4]) * inv_sum, __uint_as_float(r0[8 * g + 5]) * inv_sum);
          t.w = pack_bf16x2(__uint_as_float(r0[8 * g + 6]) * inv_sum, __uint_as_float(r0[8 * g + 7]) * inv_sum);
          o4[g] = t;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 t;
          t.x = pack_bf16x2(__uint_as_float(r1[8 * g]) * inv_sum, __uint_as_float(r1[8 * g + 1]) * inv_sum);
          t.y = pack_bf16x2(__uint_as_float(r1[8 * g + 2]) * inv_sum, __uint_as_float(r1[8 * g + 3]) * inv_sum);
          t.z = pack_bf16x2(__uint_as_float(r1[8 * g + 4]) * inv_sum, __uint_as_float(r1[8 * g + 5]) * inv_sum);
          t.w = pack_bf16x2(__uint_as_float(r1[8 * g + 6]) * inv_sum, __uint_as_float(r1[8 * g + 7]) * inv_sum);
          o4[4 + g] = t;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ persistent bf16 path
// attn_persist_kernel: one CTA per SM loops over (image, kept head) items; KVP = 208 padded keys
// (ViT: 197/198 tokens) or 256 (CCT: 256 tokens).  Per item:
//   * TMA brings the head's Q (two 128-row tiles), K and V ONCE (the per-tile kernel above loads
//     K and V once per query tile), double-buffered so item i+1 lands during item i;
//   * S_t = Q_t K^T for both query tiles goes to TMEM columns [256 t, 256 t + 208);
//   * 8 softmax warps (warp w: TMEM lane quarter w % 4 of query tile w / 4), one score row per
//     thread, two passes over the row (max, then exp2 / sum) with the next TMEM load in flight,
//     P written as packed bf16 pairs back INTO TMEM over S columns the thread has already
//     consumed (P row = columns [0, 104)) -- no shared-memory round trip, no proxy fence.
//     (Splitting each row over two warps -- 16 softmax warps -- was measured SLOWER: the pass
//     is bound by the exp2 unit, 4 lanes / clock / scheduler, not by per-warp latency.)
//   * O_t = P_t V is a tcgen05.mma with the A operand in tensor memory, accumulating into
//     columns [128, 192) of the tile (S there is dead by then); the same warps scale by 1/rowsum
//     and store.
// TMEM: 2 x 256 columns = all 512 (hence one CTA per SM); smem: 2 stages x 84 KB.
// warps 0..7: softmax (two 128-row TMEM regions x four lane quarters), warp 8: control (TMA +
// MMA issue), warps 9..16: output (read O from TMEM, scale by 1/rowsum, store) -- with the
// output on its own warps a softmax warp goes from P_k straight to S_{k+1}, and S_{k+1} is issued
// as soon as O_k has been READ, not stored
constexpr int kPersistThreads = 17 * 32;

long long* g_attn_trace = nullptr;  // devit_debug_set_trace (shared with the GEMM trace buffer)
#ifdef DEVIT_GEMM_TRACE
#define ATTN_TRACE(slot_, idx_)                                                         \
  do {                                                                                  \
    if (trace && blockIdx.x == 0 && lane == 0 && (idx_) < 512)                          \
      trace[(slot_) * 512 + (idx_)] = clock64();                                        \
  } while (0)
#else
#define ATTN_TRACE(slot_, idx_) do { } while (0)
#endif

template <int KVP>
struct AttnPersistCfg {
  static constexpr int kQBytes = 2 * 128 * 128;  // two query tiles
  static constexpr int kKVBytes = KVP * 128;
  static constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
  static constexpr int kOffBar = 2 * kStageBytes;
  static constexpr int kSmemBytes = kOffBar + 128 + 1024 + 1024;  // barriers, 1/rowsum, slack
  static constexpr int kOCols = 128;  // O accumulator: columns [128, 192) of the tile (S there
                                      // is dead once every warp has finished its second pass)
  static_assert(KVP % 16 == 0 && KVP <= 256 && KVP / 2 <= kOCols, "P and O must not overlap");
};

template <int KVP>
__global__ void __launch_bounds__(kPersistThreads, 1)
attn_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    __nv_bfloat16* __restrict__ out, int tokens, int heads, int num_items,
                    float scale_log2e, int dephase, int poll_ns, long long* trace) {
  using Cfg = AttnPersistCfg<KVP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* full_qk = bars + 0;   // [2 stages]  Q and K landed
  uint64_t* full_v = bars + 2;    // [2 stages]  V landed
  uint64_t* empty = bars + 4;     // [2 stages]  every MMA that reads the stage has finished
  uint64_t* s_full = bars + 6;    // [2 tiles]   S_t complete
  uint64_t* p_full = bars + 8;    // [2 tiles]   P_t written (4 warp arrivals)
  uint64_t* o_full = bars + 10;   // [2 tiles]   O_t complete
  uint64_t* o_empty = bars + 12;  // [2 tiles]   O_t read out, S_t columns reusable (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  // 1 / rowsum of the current item, [region][row]: written by the softmax warp before it arrives
  // on p_full, read by the output warp of the same lane quarter after o_full (which follows
  // p_full), overwritten only after the softmax warp has seen the next s_full (which follows
  // o_empty of this item)
  float* inv_s = reinterpret_cast<float*>(smem + Cfg::kOffBar + 128);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&full_qk[i], 1);
        mbar_init(&full_v[i], 1);
        mbar_init(&empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_empty[i], 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int first = blockIdx.x, step = gridDim.x;
  griddep_wait();  // qkv is the previous kernel's output

  if (warp == 8) {
    // ------------------------------------------------------------ control: TMA + MMA issue
    auto load_item = [&](int item, int stage) {
      const int img = item / heads, head = item - img * heads;
      uint8_t* sq = smem + stage * Cfg::kStageBytes;
      uint8_t* sk = sq + Cfg::kQBytes;
      uint8_t* sv = sk + Cfg::kKVBytes;
      if (elect_one()) {
        mbar_expect_tx(&full_qk[stage], Cfg::kQBytes + Cfg::kKVBytes);
        tma_load_3d(sq, &tmQ, &full_qk[stage], head * 64, 0, img);
        tma_load_3d(sq + 16384, &tmQ, &full_qk[stage], head * 64, 128, img);
        tma_load_3d(sk, &tmKV, &full_qk[stage], (heads + head) * 64, 0, img);
        mbar_expect_tx(&full_v[stage], Cfg::kKVBytes);
        tma_load_3d(sv, &tmKV, &full_v[stage], (2 * heads + head) * 64, 0, img);
      }
    };
    const uint32_t idesc_s = make_idesc(kFmtBF16, 128, KVP, 0, 0);
    const uint32_t idesc_o = make_idesc(kFmtBF16, 128, 64, 0, 1);
    const int n_mine = first < num_items ? (num_items - first + step - 1) / step : 0;
    // Event loop.  The two query tiles are independent pipelines (S -> softmax -> PV -> store)
    // that share the exp2 unit; issuing them in lock-step makes both hit it at the same time and
    // both idle together afterwards.  So every step is issued as soon as ITS inputs are ready
    // (non-blocking barrier probes), and tile 1's first S waits for tile 0's first P, which puts
    // the pipelines half a period apart: one tile's exp2 pass runs under the other's MMAs,
    // stores and max pass.
    int k_load = 0;            // next item to request (stage k_load & 1)
    int k_s[2] = {0, 0};       // next item whose S_t is to be issued
    int k_pv[2] = {0, 0};      // next item whose P_t V is to be issued
    auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {
      return __shfl_sync(0xffffffffu, mbar_test_wait(bar, parity) ? 1 : 0, 0) != 0;
    };
    while (k_pv[0] < n_mine || k_pv[1] < n_mine) {
      bool progress = false;
      // ---- loads: two stages; stage of item k is free once both P V of item k-2 retired
      if (k_load < n_mine && (k_load < 2 || ready(&empty[k_load & 1], ((k_load - 2) >> 1) & 1))) {
        load_item(first + k_load * step, k_load & 1);
        ++k_load;
        progress = true;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (k_s[t] == k_pv[t] && k_s[t] < n_mine) {
          // ---- S_t = Q_t K^T (M=128, N=KVP, K=64: 4 UMMAs)
          const int k = k_s[t], stage = k & 1;
          bool ok = k < k_load && ready(&full_qk[stage], (k >> 1) & 1);
          if (ok && k >= 1) ok = ready(&o_empty[t], (k - 1) & 1);       // tile's columns free
          if (ok && dephase && t == 1 && k == 0) ok = ready(&p_full[0], 0);  // phase offset
          if (ok) {
            tc_fence_after();
            const uint32_t sq = smem_u32(smem + stage * Cfg::kStageBytes);
            // region t takes query tile t ^ (k & 1): the tiles of an item are unequal (198 tokens
            // = 128 + 70 rows), alternating them gives both regions' warps the same load
            const uint64_t dq = make_sw128_desc(sq + ((t ^ (k & 1)) * 16384), 1024, 16);
            const uint64_t dk = make_sw128_desc(sq + Cfg::kQBytes, 1024, 16);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_bf16(tmem_base + t * 256, dq + 2 * j, dk + 2 * j, idesc_s, j > 0);
              umma_commit(&s_full[t]);
            }
            ATTN_TRACE(2 + t, k);
            ++k_s[t];
            progress = true;
          }
        } else if (k_pv[t] < k_s[t]) {
          // ---- O_t = P_t V (A from TMEM, V MN-major in smem: 16 keys = 2 KB per UMMA)
          const int k = k_pv[t], stage = k & 1;
          if (ready(&p_full[t], k & 1) && ready(&full_v[stage], (k >> 1) & 1)) {
            ATTN_TRACE(4 + t, k);
            tc_fence_after();
            const uint32_t sv =
                smem_u32(smem + stage * Cfg::kStageBytes) + Cfg::kQBytes + Cfg::kKVBytes;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < KVP / 16; ++j) {
                const uint64_t dv = make_sw128_desc(sv + j * 2048, 1024, 1024);
                umma_bf16_ts(tmem_base + t * 256 + Cfg::kOCols, tmem_base + t * 256 + 8 * j, dv,
                             idesc_o, j > 0);
              }
              umma_commit(&o_full[t]);
              // the stage is free once BOTH tiles' P V of this item retired: commit after the
              // second of the two to be issued (a commit covers every earlier MMA)
              if (k_pv[t ^ 1] > k) umma_commit(&empty[stage]);
            }
            ATTN_TRACE(6 + t, k);
            ++k_pv[t];
            progress = true;
          }
        }
      }
      if (!progress && poll_ns > 0) __nanosleep(poll_ns);  // leave issue slots to the softmax warps
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ softmax warps
    const int quarter = warp & 3;  // TMEM lane quarter
    const int t = warp >> 2;       // TMEM region
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
    constexpr int kFull = KVP / 32;   // 32-column chunks ...
    constexpr int kTail = KVP % 32;   // ... then a 16-column tail (KVP = 208) or nothing (256)
    int k = 0;
    for (int item = first; item < num_items; item += step, ++k) {
      const int q = t ^ (k & 1);   // query tile of this item handled in region t
      const bool warp_live = (q * 128 + quarter * 32) < tokens;
      if (quarter == 0) ATTN_TRACE(8 + 6 * t, k);
      mbar_wait_warp(&s_full[t], k & 1);
      if (quarter == 0) ATTN_TRACE(9 + 6 * t, k);
      tc_fence_after();
      float inv_sum = 0.f;
      if (warp_live) {
        // Both passes are fully unrolled with the TMEM load of chunk c+1 in flight while chunk c
        // is processed (tcgen05.wait::ld waits for everything outstanding, so one load ahead).
        uint32_t r[kFull + (kTail ? 1 : 0)][32];
        // ---- pass 1: row max over the valid keys
        float mx = -INFINITY;
        tmem_ld_x32(t_row, r[0]);
#pragma unroll
        for (int c = 0; c < kFull; ++c) {
          tmem_ld_wait();
          if (c + 1 < kFull) tmem_ld_x32(t_row + (c + 1) * 32, r[c + 1]);
          else if (kTail) tmem_ld_x16(t_row + kFull * 32, r[kFull - (kTail ? 0 : 1)]);
          if ((c + 1) * 32 <= tokens) {
#pragma unroll
            for (int j = 0; j < 32; j += 2)
              mx = fmax3(mx, __uint_as_float(r[c][j]), __uint_as_float(r[c][j + 1]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[c][j]));
          }
        }
        if (kTail) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (kFull * 32 + j < tokens)
              mx = fmaxf(mx, __uint_as_float(r[kFull - (kTail ? 0 : 1)][j]));
        }
        if (quarter == 0) ATTN_TRACE(10 + 6 * t, k);
        // ---- pass 2: p = exp2(s*c - max*c), row sum, P -> TMEM as packed bf16 pairs
        const float moff = mx * scale_log2e;
        float sum = 0.f;
        tmem_ld_x32(t_row, r[0]);
#pragma unroll
        for (int c = 0; c < kFull; ++c) {
          tmem_ld_wait();
          if (c + 1 < kFull) tmem_ld_x32(t_row + (c + 1) * 32, r[c + 1]);
          else if (kTail) tmem_ld_x16(t_row + kFull * 32, r[kFull - (kTail ? 0 : 1)]);
          uint32_t pk[16];
          if ((c + 1) * 32 <= tokens) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float e0 = fast_exp2(fmaf(__uint_as_float(r[c][2 * j]), scale_log2e, -moff));
              const float e1 =
                  fast_exp2(fmaf(__uint_as_float(r[c][2 * j + 1]), scale_log2e, -moff));
              sum += e0 + e1;
              pk[j] = pack_bf16x2(e0, e1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float e0 = fast_exp2(fmaf(__uint_as_float(r[c][2 * j]), scale_log2e, -moff));
              float e1 = fast_exp2(fmaf(__uint_as_float(r[c][2 * j + 1]), scale_log2e, -moff));
              e0 = (c * 32 + 2 * j < tokens) ? e0 : 0.f;
              e1 = (c * 32 + 2 * j + 1 < tokens) ? e1 : 0.f;
              sum += e0 + e1;
              pk[j] = pack_bf16x2(e0, e1);
            }
          }
          // P chunk c -> columns [16c, 16c + 16): S columns this thread has already consumed
          // (chunk c+1, in flight, lies above them)
          tmem_st_x16(t_row + c * 16, pk);
        }
        if (kTail) {
          constexpr int kT = kFull - (kTail ? 0 : 1);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float e0 = fast_exp2(fmaf(__uint_as_float(r[kT][2 * j]), scale_log2e, -moff));
            float e1 = fast_exp2(fmaf(__uint_as_float(r[kT][2 * j + 1]), scale_log2e, -moff));
            e0 = (kFull * 32 + 2 * j < tokens) ? e0 : 0.f;
            e1 = (kFull * 32 + 2 * j + 1 < tokens) ? e1 : 0.f;
            sum += e0 + e1;
            pk[j] = pack_bf16x2(e0, e1);
          }
          tmem_st_x8(t_row + kFull * 16, pk);
        }
        tmem_st_wait();
        inv_sum = 1.0f / sum;
      }
      inv_s[t * 128 + quarter * 32 + lane] = inv_sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (quarter == 0) ATTN_TRACE(11 + 6 * t, k);
    }
  } else {
    // ------------------------------------------------------------ output warps
    const int quarter = warp & 3;    // TMEM lane quarter (a warp reaches lanes 32 (warp % 4) ..)
    const int t = (warp - 9) >> 2;   // TMEM region
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
    int k = 0;
    for (int item = first; item < num_items; item += step, ++k) {
      const int img = item / heads, head = item - img * heads;
      const int q = t ^ (k & 1);
      const int row = q * 128 + quarter * 32 + lane;
      const bool warp_live = (q * 128 + quarter * 32) < tokens;
      mbar_wait_warp(&p_full[t], k & 1);  // orders the softmax warps' 1/rowsum writes
      mbar_wait_warp(&o_full[t], k & 1);
      if (quarter == 0) ATTN_TRACE(12 + 6 * t, k);
      tc_fence_after();
      const float inv_sum = inv_s[t * 128 + quarter * 32 + lane];
      if (warp_live) {
        uint32_t r0[32], r1[32];
        tmem_ld_x32(t_row + Cfg::kOCols, r0);
        tmem_ld_x32(t_row + Cfg::kOCols + 32, r1);
        tmem_ld_wait();
        // O is in registers: hand the region's TMEM columns back BEFORE converting and storing
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[t]);
        if (row < tokens) {
          __nv_bfloat16* o =
              out + (static_cast<long long>(img) * tokens + row) * (heads * 64) + head * 64;
          uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r0[8 * g]) * inv_sum, __uint_as_float(r0[8 * g + 1]) * inv_sum);
            v.y = pack_bf16x2(__uint_as_float(r0[8 * g + 2]) * inv_sum, __uint_as_float(r0[8 * g + 3]) * inv_sum);
            v.z = pack_bf16x2(__uint_as_float(r0[8 * g + 4]) * inv_sum, __uint_as_float(r0[8 * g + 5]) * inv_sum);
            v.w = pack_bf16x2(__uint_as_float(r0[8 * g + 6]) * inv_sum, __uint_as_float(r0[8 * g + 7]) * inv_sum);
            o4[g] = v;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r1[8 * g]) * inv_sum, __uint_as_float(r1[8 * g + 1]) * inv_sum);
            v.y = pack_bf16x2(__uint_as_float(r1[8 * g + 2]) * inv_sum, __uint_as_float(r1[8 * g + 3]) * inv_sum);
            v.z = pack_bf16x2(__uint_as_float(r1[8 * g + 4]) * inv_sum, __uint_as_float(r1[8 * g + 5]) * inv_sum);
            v.w = pack_bf16x2(__uint_as_float(r1[8 * g + 6]) * inv_sum, __uint_as_float(r1[8 * g + 7]) * inv_sum);
            o4[4 + g] = v;
          }
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[t]);
      }
      if (quarter == 0) ATTN_TRACE(13 + 6 * t, k);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------- fp32 parity path
// One CTA per (head, image); K (padded rows, conflict-free) and V live in shared memory as
// fp32; each warp owns query rows round-robin.
constexpr int kF32Threads = 256;

__global__ void __launch_bounds__(kF32Threads)
attn_f32_kernel(const float* __restrict__ qkv, long long in_plane, float* __restrict__ out,
                long long out_plane, int tokens, int heads, float scale) {
  extern __shared__ float sm[];
  float* sK = sm;                       // [tokens][65]
  float* sV = sK + tokens * 65;         // [tokens][64]
  float* sP = sV + tokens * 64;         // [8][256]
  float* sQ = sP + 8 * 256;             // [8][64]
  const int head = blockIdx.x, img = blockIdx.y;
  const int ld = 3 * heads * 64;
  const float* base = qkv + static_cast<long long>(img) * tokens * ld;
  for (int i = threadIdx.x; i < tokens * 64; i += kF32Threads) {
    const int t = i >> 6, d = i & 63;
    const long long ko = static_cast<long long>(t) * ld + (heads + head) * 64 + d;
    const long long vo = static_cast<long long>(t) * ld + (2 * heads + head) * 64 + d;
    sK[t * 65 + d] = base[ko] + base[ko + in_plane];
    sV[t * 64 + d] = base[vo] + base[vo + in_plane];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* myP = sP + warp * 256;
  float* myQ = sQ + warp * 64;
  for (int row = warp; row < tokens; row += 8) {
    const long long qo = static_cast<long long>(row) * ld + head * 64;
    myQ[lane] = base[qo + lane] + base[qo + lane + in_plane];
    myQ[lane + 32] = base[qo + lane + 32] + base[qo + lane + 32 + in_plane];
    __syncwarp();
    float s[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + 32 * i;
      float acc = 0.f;
      if (j < tokens) {
        const float* kr = sK + j * 65;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) acc = fmaf(myQ[d], kr[d], acc);
        acc *= scale;
        mx = fmaxf(mx, acc);
      }
      s[i] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + 32 * i;
      if (j < tokens) {
        const float e = expf(s[i] - mx);
        sum += e;
        myP[j] = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < tokens; ++j) {
      const float pj = myP[j];
      o0 = fmaf(pj, sV[j * 64 + lane], o0);
      o1 = fmaf(pj, sV[j * 64 + lane + 32], o1);
    }
    const float inv = 1.0f / sum;
    o0 *= inv;
    o1 *= inv;
    float* orow = out + (static_cast<long long>(img) * tokens + row) * (heads * 64) + head * 64;
    const float h0 = tf32_hi(o0), h1 = tf32_hi(o1);
    orow[lane] = h0;
    orow[lane + 32] = h1;
    orow[lane + out_plane] = o0 - h0;
    orow[lane + 32 + out_plane] = o1 - h1;
    __syncwarp();
  }
}

template <int KVP>
static int launch_attn_bf16(const void* qkv, void* out, int batch, int tokens, int heads,
                            float scale, cudaStream_t stream) {
  using Cfg = AttnCfg<KVP>;
  static bool attr_done[64] = {};
  int dev = 0;
  DEVIT_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    DEVIT_CUDA_OK(cudaFuncSetAttribute(attn_bf16_kernel<KVP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_done[dev & 63] = true;
  }
  const uint64_t ld = 3ull * heads * 64;
  CUtensorMap tq, tkv;
  int rc = encode_tmap_3d(&tq, qkv, 2, ld, tokens, batch, ld, ld * tokens, 64, 128, 1);
  if (rc) return rc;
  rc = encode_tmap_3d(&tkv, qkv, 2, ld, tokens, batch, ld, ld * tokens, 64, KVP, 1);
  if (rc) return rc;
  dim3 grid((tokens + 127) / 128, heads, batch);
  {
    ProfScope ps(kTagAttention, stream);
    attn_bf16_kernel<KVP><<<grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(
        tq, tkv, static_cast<__nv_bfloat16*>(out), tokens, heads, scale * 1.4426950408889634f);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

template <int KVP>
static int launch_attn_persist(const void* qkv, void* out, int batch, int tokens, int heads,
                               float scale, cudaStream_t stream) {
  using Cfg = AttnPersistCfg<KVP>;
  static bool attr_done[64] = {};
  int dev = 0;
  DEVIT_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    DEVIT_CUDA_OK(cudaFuncSetAttribute(attn_persist_kernel<KVP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_done[dev & 63] = true;
  }
  const uint64_t ld = 3ull * heads * 64;
  CUtensorMap tq, tkv;
  int rc = encode_tmap_3d(&tq, qkv, 2, ld, tokens, batch, ld, ld * tokens, 64, 128, 1);
  if (rc) return rc;
  rc = encode_tmap_3d(&tkv, qkv, 2, ld, tokens, batch, ld, ld * tokens, 64, KVP, 1);
  if (rc) return rc;
  const int items = batch * heads;
  const int grid = items < num_sms() ? items : num_sms();
  static int env_lockstep = kEnvUnread;  // debug: issue both query tiles in lock-step
  static int env_poll = kEnvUnread;      // control warp's back-off between barrier probes (ns)
  {
    ProfScope ps(kTagAttention, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPersistThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    DEVIT_CUDA_OK(cudaLaunchKernelEx(&cfg, attn_persist_kernel<KVP>, tq, tkv,
                                     static_cast<__nv_bfloat16*>(out), tokens, heads, items,
                                     scale * 1.4426950408889634f,
                                     env_int("DEVIT_ATTN_LOCKSTEP", 0, &env_lockstep) ? 0 : 1,
                                     env_int("DEVIT_ATTN_POLL_NS", 40, &env_poll), g_attn_trace));
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

}  // namespace devit

extern "C" int devit_attention(int32_t precision, const void* qkv, int64_t qkv_plane_stride,
                               void* out, int64_t out_plane_stride, int32_t batch,
                               int32_t tokens, int32_t heads, float scale, void* stream_v) {
  using namespace devit;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  int rc = check_device();
  if (rc) return rc;
  DEVIT_REQUIRE(qkv && out, "devit_attention: null pointer");
  DEVIT_REQUIRE(batch > 0 && heads > 0 && tokens > 0 && tokens <= 256,
                "devit_attention: need batch>0, heads>0, 0<tokens<=256 (got %d,%d,%d)", batch,
                heads, tokens);
  if (precision == DEVIT_BF16) {
    const int kvp = (tokens + 15) & ~15;
    static int persist = -1;  // DEVIT_ATTN_PERSIST=0: per-tile kernel only (debug / comparison)
    if (persist < 0) {
      const char* e = getenv("DEVIT_ATTN_PERSIST");
      persist = (e && e[0] == '0') ? 0 : 1;
    }
    if (kvp <= 64) return launch_attn_bf16<64>(qkv, out, batch, tokens, heads, scale, stream);
    if (kvp <= 208 && kvp > 128 && persist)
      return launch_attn_persist<208>(qkv, out, batch, tokens, heads, scale, stream);
    if (kvp > 208 && persist)
      return launch_attn_persist<256>(qkv, out, batch, tokens, heads, scale, stream);
    if (kvp <= 208) return launch_attn_bf16<208>(qkv, out, batch, tokens, heads, scale, stream);
    return launch_attn_bf16<256>(qkv, out, batch, tokens, heads, scale, stream);
  }
  DEVIT_REQUIRE(precision == DEVIT_FP32, "devit_attention: bad precision %d", precision);
  DEVIT_REQUIRE(qkv_plane_stride > 0 && out_plane_stride > 0,
                "devit_attention: DEVIT_FP32 needs plane strides");
  const size_t smem = static_cast<size_t>(tokens) * (65 + 64) * 4 + 8 * 256 * 4 + 8 * 64 * 4;
  static bool attr_done[64] = {};
  int dev = 0;
  DEVIT_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    DEVIT_CUDA_OK(cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       256 * (65 + 64) * 4 + 8 * 256 * 4 + 8 * 64 * 4));
    attr_done[dev & 63] = true;
  }
  dim3 grid(heads, batch);
  {
    ProfScope ps(kTagAttention, stream);
    attn_f32_kernel<<<grid, kF32Threads, smem, stream>>>(
        static_cast<const float*>(qkv), qkv_plane_stride, static_cast<float*>(out),
        out_plane_stride, tokens, heads, scale);
  }
  DEVIT_CUDA_OK(cudaGetLastError());
  count_launch();
  return DEVIT_OK;
}

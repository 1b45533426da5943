#include <cstdlib>

#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace devit {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_sm_budget{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int env_int(const char* name, int dflt, int* cache) {
  if (*cache == kEnvUnread) {
    const char* e = getenv(name);
    *cache = e ? atoi(e) : dflt;
  }
  return *cache;
}

int pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DEVIT_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProfRec {
  cudaEvent_t a, b;
  int tag;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

ProfScope::ProfScope(int tag, cudaStream_t s) : stream(s), on(g_prof_on) {
  if (!on) return;
  ProfRec r;
  r.tag = tag;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
    on = false;
    return;
  }
  cudaEventRecord(r.a, stream);
  end_event = r.b;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(static_cast<cudaEvent_t>(end_event), stream);
}

struct DevInfo {
  int major = -1, minor = -1, sms = 0;
};
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static int device_info(DevInfo* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return set_error(DEVIT_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= 64) return set_error(DEVIT_ERR_DEVICE, "device index %d unsupported", dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev[dev].major < 0) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess)
      return set_error(DEVIT_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    g_dev[dev].major = prop.major;
    g_dev[dev].minor = prop.minor;
    g_dev[dev].sms = prop.multiProcessorCount;
  }
  *out = g_dev[dev];
  return DEVIT_OK;
}

int check_device() {
  DevInfo d;
  int rc = device_info(&d);
  if (rc) return rc;
  if (d.major != 10)
    return set_error(DEVIT_ERR_DEVICE,
                     "devit_b200 needs an sm_100 (B200) device, found sm_%d%d; there is no "
                     "fallback path",
                     d.major, d.minor);
  return DEVIT_OK;
}

int num_sms() {
  DevInfo d;
  if (device_info(&d)) return 1;
  // devit_set_sm_budget (or DEVIT_SM_LIMIT=<n> for experiments): size every persistent grid
  // for n SMs so that several kernel chains share the chip; even, so CTA pairs stay whole
  static int limit_cache = kEnvUnread;
  int limit = g_sm_budget.load(std::memory_order_relaxed);
  if (limit <= 0) limit = env_int("DEVIT_SM_LIMIT", 0, &limit_cache);
  int sms = d.sms > 0 ? d.sms : 1;
  if (limit >= 2 && limit < sms) sms = limit & ~1;
  return sms;
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Tensor maps are pure functions of (base pointer, element size, dims, strides, box, swizzle,
// L2 promotion): every launcher used to re-encode up to seven of them per call through the driver.
// They are cached here, keyed by exactly those inputs ("mutex-guarded
// CUtensorMap caches keyed by pointer + shape"); steady-state launches of a model hit the cache.
// The cache is bounded (cleared when full) and can be switched off with DEVIT_TMAP_CACHE=0.
struct TmapKey {
  uint64_t base;
  uint64_t dims[3];
  uint64_t strides[2];
  uint32_t box[3];
  uint32_t misc;  // elem_bytes | rank << 8 | weight_like << 16 | swizzle128 << 17
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) {
      h ^= w[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;
static std::atomic<long long> g_tmap_hits{0}, g_tmap_misses{0};
constexpr size_t kTmapCacheMax = 8192;

static int encode_uncached(CUtensorMap* map, const void* base, int elem_bytes, int rank,
                           const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                           const cuuint32_t* box, bool weight_like, bool swizzle128);

static int encode(CUtensorMap* map, const void* base, int elem_bytes, int rank,
                  const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, bool weight_like, bool swizzle128 = true) {
  static int cache_on = kEnvUnread;
  if (!env_int("DEVIT_TMAP_CACHE", 1, &cache_on))
    return encode_uncached(map, base, elem_bytes, rank, dims, strides_bytes, box, weight_like,
                           swizzle128);
  TmapKey k;
  std::memset(&k, 0, sizeof(k));
  k.base = reinterpret_cast<uint64_t>(base);
  for (int i = 0; i < rank; ++i) {
    k.dims[i] = dims[i];
    k.box[i] = box[i];
    if (i + 1 < rank) k.strides[i] = strides_bytes[i];
  }
  k.misc = static_cast<uint32_t>(elem_bytes) | (static_cast<uint32_t>(rank) << 8) |
           (weight_like ? 1u << 16 : 0u) | (swizzle128 ? 1u << 17 : 0u);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(k);
    if (it != g_tmaps.end()) {
      *map = it->second;
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return DEVIT_OK;
    }
  }
  int rc = encode_uncached(map, base, elem_bytes, rank, dims, strides_bytes, box, weight_like,
                           swizzle128);
  if (rc) return rc;
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmaps.size() >= kTmapCacheMax) g_tmaps.clear();
  g_tmaps.emplace(k, *map);
  return DEVIT_OK;
}

static int encode_uncached(CUtensorMap* map, const void* base, int elem_bytes, int rank,
                           const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                           const cuuint32_t* box, bool weight_like, bool swizzle128) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return set_error(DEVIT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(DEVIT_ERR_ARG, "TMA base pointer %p is not 16-byte aligned", base);
  for (int i = 0; i < rank - 1; ++i)
    if (strides_bytes[i] % 16 != 0)
      return set_error(DEVIT_ERR_ARG, "TMA stride %llu bytes is not a multiple of 16",
                       (unsigned long long)strides_bytes[i]);
  if (swizzle128 && box[0] * (cuuint32_t)elem_bytes != 128)
    return set_error(DEVIT_ERR_ARG, "TMA inner box must be 128 bytes for SWIZZLE_128B");
  if (!swizzle128 && box[0] * (cuuint32_t)elem_bytes != 64)
    return set_error(DEVIT_ERR_ARG, "TMA inner box must be 64 bytes for SWIZZLE_64B");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapDataType dt =
      elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  weight_like ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                              : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(DEVIT_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu box "
                     "%u,%u)",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                     box[0], box[1]);
  return DEVIT_OK;
}

int encode_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols,
                   uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows,
                   bool weight_like) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(map, base, elem_bytes, 2, dims, strides, box, weight_like);
}

int encode_tmap_2d_sw64(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols,
                          uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(map, base, elem_bytes, 2, dims, strides, box, false, false);
}

int encode_tmap_3d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t d0, uint64_t d1,
                   uint64_t d2, uint64_t stride1, uint64_t stride2, uint32_t box0,
                   uint32_t box1, uint32_t box2) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1 * (uint64_t)elem_bytes, stride2 * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  return encode(map, base, elem_bytes, 3, dims, strides, box, false);
}

}  // namespace devit

extern "C" {

int devit_abi_version(void) { return DEVIT_ABI_VERSION; }
const char* devit_last_error(void) { return devit::g_err; }
int devit_device_check(void) { return devit::check_device(); }
long long devit_launch_count(void) { return devit::g_launches.load(); }
int devit_tmap_cache_stats(long long* hits, long long* misses) {
  if (hits) *hits = devit::g_tmap_hits.load();
  if (misses) *misses = devit::g_tmap_misses.load();
  std::lock_guard<std::mutex> lk(devit::g_tmap_mu);
  return static_cast<int>(devit::g_tmaps.size());
}
int devit_set_sm_budget(int sms) { return devit::g_sm_budget.exchange(sms < 0 ? 0 : sms); }

int devit_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(devit::g_prof_mu);
  for (auto& r : devit::g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  devit::g_prof.clear();
  devit::g_prof_on = on != 0;
  return DEVIT_OK;
}

int devit_profile_collect(double* ms_by_tag, long long* count_by_tag) {
  if (!ms_by_tag || !count_by_tag) return devit::set_error(DEVIT_ERR_ARG, "null output");
  std::lock_guard<std::mutex> lk(devit::g_prof_mu);
  for (int i = 0; i < devit::kNumProfTags; ++i) {
    ms_by_tag[i] = 0.0;
    count_by_tag[i] = 0;
  }
  for (auto& r : devit::g_prof) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess)
      return devit::set_error(DEVIT_ERR_CUDA, "profile event: %s", cudaGetErrorString(e));
    ms_by_tag[r.tag] += ms;
    count_by_tag[r.tag] += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  devit::g_prof.clear();
  return DEVIT_OK;
}

}  // extern "C"

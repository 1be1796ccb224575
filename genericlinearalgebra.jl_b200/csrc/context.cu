// context.cu -- error plumbing and per-thread library state (no global mutable state besides
// lazily cached device properties).
#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace gla {

static thread_local char g_err[512] = "";
thread_local double g_last_ms = 0.0;
thread_local i64 g_last_info = 0;

void set_error(int code, const char* what, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "gla error %d: %s (%s:%d)", code, what, file, line);
}

int check_cuda(cudaError_t e, const char* file, int line) {
  if (e == cudaSuccess) return 0;
  int code = 1000 + (int)e;
  set_error(code, cudaGetErrorString(e), file, line);
  (void)cudaGetLastError();  // clear sticky-less errors
  return code;
}

const char* last_error() { return g_err; }

int sm_count() {
  static std::mutex mu;
  static int cached[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (!have[dev]) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    cached[dev] = n;
    have[dev] = true;
    (void)cudaGetLastError();
  }
  return cached[dev];
}

// Library-owned stream-ordered memory pool (one per device): workspace freed with cudaFreeAsync stays in THIS pool
// (release threshold = max) instead of going back to the driver at every synchronisation, and the process-wide default
// pool -- shared with whatever else lives in the host process (Julia's CUDA.jl, torch) -- is left untouched.
int pool_malloc(void** p, size_t bytes, cudaStream_t st) {
  static std::mutex mu;
  static cudaMemPool_t pools[64];
  static bool have[64];
  int dev = 0;
  GLA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return check_cuda(cudaMallocAsync(p, bytes ? bytes : 1, st), __FILE__, __LINE__);
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
      cudaMemPoolProps props;
      memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      GLA_CUDA(cudaMemPoolCreate(&pools[dev], &props));
      unsigned long long thr = ~0ull;
      GLA_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr));
      have[dev] = true;
    }
    pool = pools[dev];
  }
  return check_cuda(cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, st), __FILE__, __LINE__);
}

namespace {
struct AuxCache {
  std::map<int, AuxCtx> per_dev;
  ~AuxCache() {
    for (auto& kv : per_dev) {
      for (auto& e : kv.second.ev)
        if (e) cudaEventDestroy(e);
      if (kv.second.hi) cudaStreamDestroy(kv.second.hi);
      if (kv.second.hi2) cudaStreamDestroy(kv.second.hi2);
    }
  }
};
}  // namespace

int aux_ctx(AuxCtx** out) {
  static thread_local AuxCache cache;
  int dev = 0;
  GLA_CUDA(cudaGetDevice(&dev));
  AuxCtx& a = cache.per_dev[dev];
  if (!a.hi) {
    int lo = 0, hi = 0;
    GLA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    GLA_CUDA(cudaStreamCreateWithPriority(&a.hi, cudaStreamNonBlocking, hi));
    GLA_CUDA(cudaStreamCreateWithPriority(&a.hi2, cudaStreamNonBlocking, hi));
    for (auto& e : a.ev) GLA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  *out = &a;
  return 0;
}

namespace {
thread_local bool pdl_scope_on = false;
}
bool pdl_enabled() {
  static const bool on = getenv("GLA_NO_PDL") == nullptr;
  return on && pdl_scope_on;
}
long long pdl_max_n() {
  static const long long v = [] { const char* e = getenv("GLA_PDL_MAXN"); return e ? atoll(e) : 2048ll; }();
  return v;
}
PdlScope::PdlScope(bool on) : prev(pdl_scope_on) { pdl_scope_on = on; }
PdlScope::~PdlScope() { pdl_scope_on = prev; }

int ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find({dev, func});
    if (it != done.end() && it->second >= bytes) return 0;
  }
  int rc = check_cuda(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), __FILE__, __LINE__);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(mu);
  int& v = done[{dev, func}];
  if (v < bytes) v = bytes;
  return 0;
}

}  // namespace gla

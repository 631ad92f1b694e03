// api.cu — process-wide plumbing of the C ABI: error string, launch counter, version.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace arcvae {
static thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }

// ---- per-device state ------------------------------------------------------------------------------------------
static constexpr int MAX_DEV = 64;
static std::atomic<int> g_once[MAX_DEV][ONCE_NSLOTS];
static std::atomic<int> g_sms[MAX_DEV];
static std::atomic<int*> g_errflag[MAX_DEV];
static std::mutex g_dev_mutex;
static int cur_dev() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < MAX_DEV) ? dev : 0;
}
bool first_use_on_device(int slot) { return g_once[cur_dev()][slot].exchange(1) == 0; }
int device_sm_count() {
  const int dev = cur_dev();
  int n = g_sms[dev].load();
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    g_sms[dev].store(n);
  }
  return n;
}
int* device_error_flag() {
  const int dev = cur_dev();
  int* p = g_errflag[dev].load();
  if (p == nullptr) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    p = g_errflag[dev].load();
    if (p == nullptr) {
      if (cudaMalloc(&p, 256) != cudaSuccess) return nullptr;
      cudaMemset(p, 0, 256);
      g_errflag[dev].store(p);
    }
  }
  return p;
}

static std::atomic<uint64_t> g_flops[TIME_NCAT];      // in MFLOP to stay integral
void count_flops(int cat, double flop) { g_flops[cat].fetch_add((uint64_t)(flop * 1e-6), std::memory_order_relaxed); }

// ---- per-category event timing -------------------------------------------------------------------------------
static bool g_timing = false;
struct Pair { cudaEvent_t a, b; int cat; };
static std::vector<Pair> g_pairs;      // recorded
static std::vector<Pair> g_free;       // recycled
static int g_open[TIME_NCAT] = {0};
static Pair g_cur[TIME_NCAT];
void timing_begin(int cat, cudaStream_t st) {
  if (!g_timing) return;
  if (cat != TIME_RECURRENCE && g_open[TIME_RECURRENCE] > 0) return;   // attributed to the enclosing recurrence scope
  if (g_open[cat]++ > 0) return;
  Pair p;
  if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
  else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
  p.cat = cat;
  cudaEventRecord(p.a, st);
  g_cur[cat] = p;
}
void timing_end(int cat, cudaStream_t st) {
  if (!g_timing) return;
  if (cat != TIME_RECURRENCE && g_open[TIME_RECURRENCE] > 0) return;
  if (--g_open[cat] > 0) return;
  cudaEventRecord(g_cur[cat].b, st);
  g_pairs.push_back(g_cur[cat]);
}
}  // namespace arcvae

extern "C" int arcvae_timing_enable(int on) {
  arcvae::g_timing = on != 0;
  return 0;
}
// synchronises the device, returns summed milliseconds and launch-scope count per category, then resets
extern "C" int arcvae_timing_read(double* ms, int* counts, int ncat) {
  using namespace arcvae;
  ARCVAE_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < ncat; i++) { ms[i] = 0.0; counts[i] = 0; }
  for (auto& p : g_pairs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, p.a, p.b);
    if (p.cat < ncat) { ms[p.cat] += t; counts[p.cat]++; }
    g_free.push_back(p);
  }
  g_pairs.clear();
  return 0;
}

extern "C" int arcvae_device_error_read(int* flag, void* stream) {
  using namespace arcvae;
  ARCVAE_REQUIRE(flag != nullptr, "flag");
  int* p = device_error_flag();
  ARCVAE_REQUIRE(p != nullptr, "device error flag allocation failed");
  ARCVAE_CUDA(cudaMemcpyAsync(flag, p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  ARCVAE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
extern "C" int arcvae_device_error_clear(void* stream) {
  using namespace arcvae;
  int* p = device_error_flag();
  ARCVAE_REQUIRE(p != nullptr, "device error flag allocation failed");
  ARCVAE_CUDA(cudaMemsetAsync(p, 0, sizeof(int), (cudaStream_t)stream));
  return 0;
}
// test hook: raise the flag the way a timed-out kernel would
extern "C" int arcvae_debug_raise_device_error(void* stream) {
  using namespace arcvae;
  int* p = device_error_flag();
  ARCVAE_REQUIRE(p != nullptr, "device error flag allocation failed");
  ARCVAE_CUDA(cudaMemsetAsync(p, 1, 1, (cudaStream_t)stream));
  return 0;
}

extern "C" double arcvae_flop_count(int category) {
  if (category < 0 || category >= arcvae::TIME_NCAT) return 0.0;
  return (double)arcvae::g_flops[category].load() * 1e6;
}

extern "C" const char* arcvae_last_error(void) { return arcvae::g_error.c_str(); }
extern "C" int arcvae_abi_version(void) { return ARCVAE_ABI_VERSION; }
extern "C" uint64_t arcvae_launch_count(void) { return arcvae::g_launches.load(); }
extern "C" int arcvae_zero(void* ptr, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  ARCVAE_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
  return 0;
}

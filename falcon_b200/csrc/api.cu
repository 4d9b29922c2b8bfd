// Library-wide state: error reporting, launch counter, device check, get_dim.
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace flc {

std::atomic<uint64_t> g_launches{0};

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// ---------------------------------------------------------------- per-kernel timing
struct ProfRecord {
  const char* name;
  cudaEvent_t a, b;
};
static std::atomic<int> g_prof_enabled{0};
static std::mutex g_prof_mutex;
static std::vector<ProfRecord> g_prof_records;
static std::vector<std::pair<std::string, std::pair<double, int>>> g_prof_summary;

ProfScope::ProfScope(const char* name, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_enabled.load(std::memory_order_relaxed)) return;
  ProfRecord r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  slot = static_cast<int>(g_prof_records.size());
  g_prof_records.push_back(r);
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  cudaEventRecord(g_prof_records[slot].b, stream);
}

static void prof_collect() {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (auto& kv : g_prof_summary) {
    acc[kv.first] = kv.second;
    order.push_back(kv.first);
  }
  for (auto& r : g_prof_records) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    if (!acc.count(r.name)) order.push_back(r.name);
    acc[r.name].first += ms;
    acc[r.name].second += 1;
  }
  g_prof_records.clear();
  g_prof_summary.clear();
  for (auto& n : order) g_prof_summary.push_back({n, acc[n]});
}

}  // namespace flc

extern "C" {

void flc_profile_enable(int on) { flc::g_prof_enabled.store(on ? 1 : 0); }

void flc_profile_reset(void) {
  flc::prof_collect();
  std::lock_guard<std::mutex> lock(flc::g_prof_mutex);
  flc::g_prof_summary.clear();
}

int flc_profile_count(void) {
  flc::prof_collect();
  return static_cast<int>(flc::g_prof_summary.size());
}

int flc_profile_get(int i, char* name, int name_bytes, double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lock(flc::g_prof_mutex);
  if (i < 0 || i >= static_cast<int>(flc::g_prof_summary.size()))
    return flc::set_error(FLC_ERR_INVALID, "profile index out of range");
  const auto& e = flc::g_prof_summary[i];
  snprintf(name, name_bytes, "%s", e.first.c_str());
  *total_ms = e.second.first;
  *launches = e.second.second;
  return FLC_OK;
}

const char* flc_last_error(void) { return flc::error_buffer(); }

int flc_version(void) { return 100; }

uint64_t flc_launch_count(void) { return flc::g_launches.load(); }

void flc_reset_launch_count(void) { flc::g_launches.store(0); }

int flc_check_device(int device) {
  cudaDeviceProp prop;
  FLC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return flc::set_error(FLC_ERR_UNSUPPORTED,
                          "falcon_b200 needs an sm_100 (B200) device, found sm_%d%d (%s)",
                          prop.major, prop.minor, prop.name);
  return FLC_OK;
}

// falcon/cluster/spectrum.py:172-199.  The numba signature forces float32
// arithmetic; Python's % on positive floats is fmod.
int flc_get_dim(float min_mz, float max_mz, float bin_size, uint32_t* vec_len,
                float* start_dim, float* end_dim) {
  FLC_REQUIRE(bin_size > 0.f, "bin_size must be positive");
  FLC_REQUIRE(vec_len && start_dim && end_dim, "null output pointer");
  volatile float lo_mod = fmodf(min_mz, bin_size);
  volatile float start = min_mz - lo_mod;
  volatile float hi_plus = max_mz + bin_size;
  volatile float hi_mod = fmodf(max_mz, bin_size);
  volatile float end = hi_plus - hi_mod;
  volatile float span = end - start;
  volatile float q = span / bin_size;
  *vec_len = static_cast<uint32_t>(ceilf(q));
  *start_dim = start;
  *end_dim = end;
  return FLC_OK;
}

}  // extern "C"

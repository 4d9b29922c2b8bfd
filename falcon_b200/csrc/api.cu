// Library-wide state: error reporting, launch counter, device check, get_dim.
#include <cmath>

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

}  // namespace flc

extern "C" {

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

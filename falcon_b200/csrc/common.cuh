// Shared helpers for the falcon_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/falcon_b200.h"

namespace flc {

// ---------------------------------------------------------------- errors
char* error_buffer();  // thread local, 512 bytes (api.cu)
int set_error(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(uint64_t k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

#define FLC_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t err__ = (expr);                                                             \
    if (err__ != cudaSuccess)                                                               \
      return flc::set_error(FLC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                            cudaGetErrorString(err__), __FILE__, __LINE__);                 \
  } while (0)

#define FLC_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    flc::count_launch();                                                                    \
    cudaError_t err__ = cudaGetLastError();                                                 \
    if (err__ != cudaSuccess)                                                               \
      return flc::set_error(FLC_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                            cudaGetErrorString(err__), __FILE__, __LINE__);                 \
  } while (0)

#define FLC_REQUIRE(cond, ...)                                                              \
  do {                                                                                      \
    if (!(cond)) return flc::set_error(FLC_ERR_INVALID, __VA_ARGS__);                       \
  } while (0)

#define FLC_TRY(expr)                                                                       \
  do {                                                                                      \
    int rc__ = (expr);                                                                      \
    if (rc__ != FLC_OK) return rc__;                                                        \
  } while (0)

// ---------------------------------------------------------------- per-kernel timing
// When enabled (flc_profile_enable), a ProfScope brackets a kernel launch with
// CUDA events on the launching stream; flc_profile_get() reports the summed
// device time and launch count per kernel name.
struct ProfScope {
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};

template <typename F>
inline void timed(const char* name, cudaStream_t stream, F&& launch) {
  ProfScope scope(name, stream);
  launch();
}
template <typename F>
inline void timed(const char* name, flc_stream_t stream, F&& launch) {
  ProfScope scope(name, static_cast<cudaStream_t>(stream));
  launch();
}

// ---------------------------------------------------------------- workspace carving
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  bool ok;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0), ok(true) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    if (base == nullptr) {  // sizing pass
      used += bytes;
      return nullptr;
    }
    if (used + bytes > size) {
      ok = false;
      used += bytes;
      return nullptr;
    }
    T* out = reinterpret_cast<T*>(base + used);
    used += bytes;
    return out;
  }
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline cudaStream_t as_stream(flc_stream_t s) { return static_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// MurmurHash3_x86_32 of one little-endian int32 key (len = 4, no tail).
__device__ __forceinline__ uint32_t murmur3_int32(uint32_t key, uint32_t seed) {
  uint32_t k = key * 0xcc9e2d51u;
  k = rotl32(k, 15);
  k *= 0x1b873593u;
  uint32_t h = seed ^ k;
  h = rotl32(h, 13);
  h = h * 5u + 0xe6546b64u;
  h ^= 4u;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint16_t f32_to_bf16_rne(float f) {
  uint32_t u = __float_as_uint(f);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace flc

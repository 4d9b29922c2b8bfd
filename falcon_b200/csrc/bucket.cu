// Stage a5: precursor-mass buckets.  Bucket key, stable ordering by
// (charge, interval, precursor m/z) and bucket offsets, all on the device.
// Sorting is plumbing and uses CUB's radix sort from the CUDA toolkit.
#include <cub/cub.cuh>

#include "common.cuh"

namespace flc {

constexpr double kHydrogenMass = 1.00794;
constexpr double kIsotopeSpacing = 1.0005079;

__global__ void bucket_key_kernel(const double* __restrict__ mz, const int32_t* __restrict__ charge,
                                  int64_t n, int32_t mz_interval, uint32_t* __restrict__ key,
                                  uint64_t* __restrict__ mz_bits, int32_t* __restrict__ idx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = mz[i];
  int32_t z = charge[i];
  const int32_t az = max(abs(z), 1);
  // round-half-even like Python's round(); floor division by the interval width.
  const long long neutral = llrint((m - kHydrogenMass) * static_cast<double>(az) / kIsotopeSpacing);
  long long q = neutral / mz_interval;
  if ((neutral % mz_interval != 0) && ((neutral < 0) != (mz_interval < 0))) --q;
  const uint32_t zc = static_cast<uint32_t>(min(max(z, 0), 255));
  key[i] = (zc << 24) | (static_cast<uint32_t>(q) & 0xFFFFFFu);
  uint64_t b = static_cast<uint64_t>(__double_as_longlong(m));
  b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
  mz_bits[i] = b;
  idx[i] = static_cast<int32_t>(i);
}

template <typename T>
__global__ void gather_kernel(const T* __restrict__ in, const int32_t* __restrict__ order, int64_t n,
                              T* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[order[i]];
}

__global__ void scatter32_kernel(const uint32_t* __restrict__ in, const int32_t* __restrict__ order,
                                 int64_t n, uint32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[order[i]] = in ? in[i] : static_cast<uint32_t>(i);  // in == NULL: inverse permutation
}

__global__ void bucket_head_kernel(const uint32_t* __restrict__ key, int64_t n,
                                   uint8_t* __restrict__ flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0) || (key[i] != key[i - 1]);
}

// bucket_ptr[n_buckets .. max(n_buckets, pad_to)] = n: the end of the last bucket, and (pad_to) empty
// buckets behind it so that later stages can run with a host-side upper bound of the bucket count.
__global__ void bucket_close_kernel(int64_t* bucket_ptr, const int64_t* n_buckets, int64_t n, int64_t pad_to,
                                    int64_t* n_buckets_dev) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nb = *n_buckets;
  if (i == 0 && n_buckets_dev) *n_buckets_dev = nb;
  if (i == 0 || nb + i <= pad_to) bucket_ptr[nb + i] = n;
}

struct BucketSortLayout {
  uint32_t* key_a;
  uint32_t* key_b;
  uint64_t* mz_a;
  uint64_t* mz_b;
  int32_t* idx_a;
  int32_t* idx_b;
  uint8_t* flag;
  int64_t* n_sel;
  void* cub_tmp;
  size_t cub_bytes;
};

static int bucket_sort_layout(Workspace& ws, int64_t n, BucketSortLayout& L) {
  L.key_a = ws.take<uint32_t>(n);
  L.key_b = ws.take<uint32_t>(n);
  L.mz_a = ws.take<uint64_t>(n);
  L.mz_b = ws.take<uint64_t>(n);
  L.idx_a = ws.take<int32_t>(n);
  L.idx_b = ws.take<int32_t>(n);
  L.flag = ws.take<uint8_t>(n);
  L.n_sel = ws.take<int64_t>(1);
  size_t b1 = 0, b2 = 0, b3 = 0;
  const int num = static_cast<int>(n);
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (int32_t*)nullptr, (int32_t*)nullptr, num);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (int32_t*)nullptr, (int32_t*)nullptr, num);
  cub::DeviceSelect::Flagged(nullptr, b3, cub::CountingInputIterator<int64_t>(0), (uint8_t*)nullptr,
                             (int64_t*)nullptr, (int64_t*)nullptr, num);
  L.cub_bytes = b1 > b2 ? b1 : b2;
  if (b3 > L.cub_bytes) L.cub_bytes = b3;
  L.cub_tmp = ws.take<char>(L.cub_bytes);
  return FLC_OK;
}

}  // namespace flc

extern "C" {

size_t flc_bucket_sort_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  flc::Workspace ws(nullptr, 0);
  flc::BucketSortLayout L;
  flc::bucket_sort_layout(ws, n, L);
  return ws.used + 256;
}

int flc_bucket_sort(const double* precursor_mz, const int32_t* charge, int64_t n,
                    int32_t mz_interval, int32_t* order, uint32_t* key_sorted, double* mz_sorted,
                    int64_t* bucket_ptr, int64_t* n_buckets, int64_t pad_buckets, int64_t* n_buckets_dev,
                    void* workspace, size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "n out of range");
  FLC_REQUIRE(mz_interval > 0, "mz_interval must be positive");
  FLC_REQUIRE(n_buckets != nullptr || n_buckets_dev != nullptr, "no output for n_buckets");
  FLC_REQUIRE(pad_buckets >= 0 && pad_buckets <= n, "pad_buckets must be in [0, n]");
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    if (n_buckets) *n_buckets = 0;
    FLC_CUDA(cudaMemsetAsync(bucket_ptr, 0, sizeof(int64_t), stream));
    if (n_buckets_dev) FLC_CUDA(cudaMemsetAsync(n_buckets_dev, 0, sizeof(int64_t), stream));
    return FLC_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  BucketSortLayout L;
  bucket_sort_layout(ws, n, L);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "bucket_sort workspace too small: need %zu", ws.used);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  const int num = static_cast<int>(n);
  timed("bucket_key", stream, [&] { bucket_key_kernel<<<blocks, 256, 0, stream>>>(precursor_mz, charge, n, mz_interval, L.key_a, L.mz_a,
                                                L.idx_a); });
  FLC_LAUNCH_CHECK();
  size_t tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, L.mz_a, L.mz_b, L.idx_a, L.idx_b, num, 0, 64,
                                           stream));
  count_launch(9);
  timed("gather", stream, [&] { gather_kernel<uint32_t><<<blocks, 256, 0, stream>>>(L.key_a, L.idx_b, n, L.key_b); });
  FLC_LAUNCH_CHECK();
  tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, L.key_b, key_sorted, L.idx_b, order, num, 0, 32,
                                           stream));
  count_launch(5);
  timed("gather", stream, [&] { gather_kernel<double><<<blocks, 256, 0, stream>>>(precursor_mz, order, n, mz_sorted); });
  FLC_LAUNCH_CHECK();
  timed("bucket_head", stream, [&] { bucket_head_kernel<<<blocks, 256, 0, stream>>>(key_sorted, n, L.flag); });
  FLC_LAUNCH_CHECK();
  tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceSelect::Flagged(L.cub_tmp, tmp, cub::CountingInputIterator<int64_t>(0), L.flag,
                                      bucket_ptr, L.n_sel, num, stream));
  count_launch(2);
  timed("bucket_close", stream, [&] {
    bucket_close_kernel<<<static_cast<unsigned>((pad_buckets + 1 + 255) / 256), 256, 0, stream>>>(
        bucket_ptr, L.n_sel, n, pad_buckets, n_buckets_dev); });
  FLC_LAUNCH_CHECK();
  if (n_buckets) {  // NULL: no synchronisation, the count stays on the device (n_buckets_dev)
    FLC_CUDA(cudaMemcpyAsync(n_buckets, L.n_sel, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    FLC_CUDA(cudaStreamSynchronize(stream));
  }
  return FLC_OK;
}

int flc_gather(const void* in, const int32_t* order, int64_t n, int elem_bytes, void* out,
               flc_stream_t stream) {
  using namespace flc;
  FLC_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
  if (n <= 0) return FLC_OK;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (elem_bytes == 4)
    timed("gather", stream, [&] { gather_kernel<uint32_t><<<blocks, 256, 0, as_stream(stream)>>>(
        static_cast<const uint32_t*>(in), order, n, static_cast<uint32_t*>(out)); });
  else
    timed("gather", stream, [&] { gather_kernel<uint64_t><<<blocks, 256, 0, as_stream(stream)>>>(
        static_cast<const uint64_t*>(in), order, n, static_cast<uint64_t*>(out)); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

int flc_scatter32(const void* in, const int32_t* order, int64_t n, void* out, flc_stream_t stream) {
  using namespace flc;
  if (n <= 0) return FLC_OK;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  timed("scatter32", stream, [&] { scatter32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const uint32_t*>(in), order, n,
                                                        static_cast<uint32_t*>(out)); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"

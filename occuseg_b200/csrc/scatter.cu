// ResolutionBasedScattering: for every high-resolution point the row of the low-resolution voxel it falls into.
// Replaces ResolutionBasedScatteringCuda (Metadata/ConvolutionRules.h:327-342: cudpp multivalue hash insert of the lr
// points + retrieve of hr / stride) with one radix sort + unique + binary search -- no hash table, no cuckoo rebuilds.
// Row semantics are the reference's: the value of an lr key is its RANK among the sorted unique lr keys
// (CUDA/CUDPPWrapper.hpp:789-829, hash_multivalue.cpp:59-110), i.e. the index into points_lr whenever that list is a
// sample's spatial locations in row order (how sparseconvnet/utils.py:72-132 uses it); -1 (0xFFFFFFFF) = no such voxel.
#include "common.cuh"
#include <cub/cub.cuh>

namespace scn {

__global__ void k_xyz_keys(const int *__restrict__ xyz, long long n, int stride, uint64_t *__restrict__ keys) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // integer division truncates toward zero like ATen's (point_hr_flat / stride); negative coordinates never match
  const int x = xyz[3 * i] / stride, y = xyz[3 * i + 1] / stride, z = xyz[3 * i + 2] / stride;
  const bool ok = x >= 0 && y >= 0 && z >= 0 && x < COORD_LIMIT && y < COORD_LIMIT && z < COORD_LIMIT;
  keys[i] = ok ? make_key(0u, (uint32_t)z, (uint32_t)y, (uint32_t)x) : EMPTY_KEY;
}

__global__ void k_rank_lookup(const uint64_t *__restrict__ queries, long long n, const uint64_t *__restrict__ uniq,
                              const int *__restrict__ n_uniq, int *__restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t q = queries[i];
  int lo = 0, hi = *n_uniq;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&uniq[mid]) < q) lo = mid + 1;
    else hi = mid;
  }
  out[i] = (q != EMPTY_KEY && lo < *n_uniq && __ldg(&uniq[lo]) == q) ? lo : -1;
}

void resolution_scatter(const int *lr_xyz, long long n_lr, const int *hr_xyz, long long n_hr, int stride, int *hr2lr,
                        cudaStream_t s) {
  SCN_CHECK(stride >= 1, "ResolutionBasedScattering: stride must be >= 1");
  SCN_CHECK(n_lr >= 0 && n_hr >= 0 && n_lr < (1ll << 31) && n_hr < (1ll << 31), "ResolutionBasedScattering: bad sizes");
  if (n_hr == 0) return;
  DevBuf<uint64_t> keys, sorted, uniq, q;
  DevBuf<int> n_uniq;
  n_uniq.alloc(1, s);
  SCN_CUDA(cudaMemsetAsync(n_uniq.p, 0, sizeof(int), s));
  q.alloc((size_t)n_hr, s);
  const int B = 256;
  if (n_lr > 0) {
    keys.alloc((size_t)n_lr, s);
    sorted.alloc((size_t)n_lr, s);
    uniq.alloc((size_t)n_lr, s);
    k_xyz_keys<<<(unsigned)((n_lr + B - 1) / B), B, 0, s>>>(lr_xyz, n_lr, 1, keys.p);
    SCN_LAUNCH_CHECK();
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t1, keys.p, sorted.p, (int)n_lr, 0, 48, s);
    cub::DeviceSelect::Unique(nullptr, t2, sorted.p, uniq.p, n_uniq.p, (int)n_lr, s);
    DevBuf<uint8_t> tmp;
    tmp.alloc(t1 > t2 ? t1 : t2, s);
    size_t tb = tmp.n;
    SCN_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, sorted.p, (int)n_lr, 0, 48, s));
    tb = tmp.n;
    SCN_CUDA(cub::DeviceSelect::Unique(tmp.p, tb, sorted.p, uniq.p, n_uniq.p, (int)n_lr, s));
    count_launch(6);
    tmp.release(s);
  }
  k_xyz_keys<<<(unsigned)((n_hr + B - 1) / B), B, 0, s>>>(hr_xyz, n_hr, stride, q.p);
  SCN_LAUNCH_CHECK();
  k_rank_lookup<<<(unsigned)((n_hr + B - 1) / B), B, 0, s>>>(q.p, n_hr, uniq.p, n_uniq.p, hr2lr);
  SCN_LAUNCH_CHECK();
  keys.release(s); sorted.release(s); uniq.release(s); q.release(s); n_uniq.release(s);
}

}  // namespace scn

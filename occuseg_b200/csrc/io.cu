// InputLayer / OutputLayer feature movement.  Replaces InputLayer_fp_/bp_ (CUDA/IOLayers.cu:16-75) and
// their four drivers (CUDA/IOLayers.cpp:17-154).  The rule table is the device-resident CSR built by
// build_input_level(); the reference re-uploads a host table on every call (IOLayers.cu:37-38) and
// resolves duplicates with atomicAdd in the backward direction -- here every output element has one
// owner thread, so results are deterministic.
#include "common.cuh"

namespace scn {

template <int VEC> struct VecT;
template <> struct VecT<1> { typedef float T; };
template <> struct VecT<4> { typedef float4 T; };

__device__ __forceinline__ float vzero(float) { return 0.f; }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vfma(float &acc, float m, float v) { acc += m * v; }
__device__ __forceinline__ void vfma(float4 &acc, float m, float4 v) {
  acc.x += m * v.x; acc.y += m * v.y; acc.z += m * v.z; acc.w += m * v.w;
}
__device__ __forceinline__ float vscale(float m, float v) { return m * v; }
__device__ __forceinline__ float4 vscale(float m, float4 v) { return make_float4(m * v.x, m * v.y, m * v.z, m * v.w); }

// out[row] = sum_{p in row} mult * src[p]   (rows own their points; order = original point order)
template <typename V>
__global__ void k_rows_from_points(const V *__restrict__ src, V *__restrict__ dst, const int *__restrict__ ptr,
                                   const int *__restrict__ pts, long long n_rows, int cv, bool average) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n_rows * cv) return;
  long long row = e / cv;
  int c = (int)(e - row * cv);
  int b = ptr[row], en = ptr[row + 1];
  // (T)1 / nActive then out += multiplier * inp, exactly as IOLayers.cu:24-29
  float mult = (average && en > b) ? 1.0f / (float)(en - b) : 1.0f;
  V acc = vzero(V());
  for (int j = b; j < en; ++j) vfma(acc, mult, src[(long long)pts[j] * cv + c]);
  dst[e] = acc;
}

// dst[p] = mult(row(p)) * src[row(p)]
template <typename V>
__global__ void k_points_from_rows(const V *__restrict__ src, V *__restrict__ dst, const int *__restrict__ ptr,
                                   const int *__restrict__ row_of_point, long long n_points, int cv, bool average) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n_points * cv) return;
  long long p = e / cv;
  int c = (int)(e - p * cv);
  int row = row_of_point[p];
  float mult = 1.0f;
  if (average) {
    int cnt = ptr[row + 1] - ptr[row];
    mult = 1.0f / (float)cnt;
  }
  dst[e] = vscale(mult, src[(long long)row * cv + c]);
}

static void rows_from_points(Meta *m, const float *src, float *dst, int C, bool average, cudaStream_t s) {
  SCN_CHECK(!m->levels.empty(), "InputLayer has not been built on this handle");
  long long n = m->levels[0]->n;
  if (n == 0) return;
  bool v4 = (C % 4 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  int cv = v4 ? C / 4 : C;
  long long total = n * cv;
  int grid = (int)((total + 255) / 256);
  if (v4)
    k_rows_from_points<float4><<<grid, 256, 0, s>>>((const float4 *)src, (float4 *)dst, m->rule_ptr.p, m->rule_pts.p, n, cv, average);
  else
    k_rows_from_points<float><<<grid, 256, 0, s>>>(src, dst, m->rule_ptr.p, m->rule_pts.p, n, cv, average);
  SCN_LAUNCH_CHECK();
}

static void points_from_rows(Meta *m, const float *src, float *dst, int C, bool average, cudaStream_t s) {
  SCN_CHECK(!m->levels.empty(), "InputLayer has not been built on this handle");
  long long P = m->n_points;
  if (P == 0) return;
  bool v4 = (C % 4 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  int cv = v4 ? C / 4 : C;
  long long total = P * cv;
  int grid = (int)((total + 255) / 256);
  if (v4)
    k_points_from_rows<float4><<<grid, 256, 0, s>>>((const float4 *)src, (float4 *)dst, m->rule_ptr.p, m->row_of_point.p, P, cv, average);
  else
    k_points_from_rows<float><<<grid, 256, 0, s>>>(src, dst, m->rule_ptr.p, m->row_of_point.p, P, cv, average);
  SCN_LAUNCH_CHECK();
}

void input_layer_fwd(Meta *m, const float *feats, int C, float *out, cudaStream_t s) {
  rows_from_points(m, feats, out, C, m->mode == 4, s);
}
void input_layer_bwd(Meta *m, const float *d_out, int C, float *d_feats, cudaStream_t s) {
  points_from_rows(m, d_out, d_feats, C, m->mode == 4, s);
}
// OutputLayer: roles swapped, never averaged (IOLayers.cpp:127-129, :150-152)
void output_layer_fwd(Meta *m, const float *in, int C, float *out, cudaStream_t s) {
  points_from_rows(m, in, out, C, false, s);
}
void output_layer_bwd(Meta *m, const float *d_out, int C, float *d_in, cudaStream_t s) {
  rows_from_points(m, d_out, d_in, C, false, s);
}

// Float point cloud -> the InputLayer's integer coordinate list, on the device.  The reference does this on the host:
// the data loader shifts the augmented cloud so that it starts at (10,10,10) plus a random sub-voxel offset, drops points
// outside [0, full_scale) (examples/ScanNet/datasets/scannet.py:133-137,160), appends the sample index as 4th column
// (:210) and InputLayer truncates the float list with .type(torch.LongTensor) (sparseconvnet/ioLayers.py:56).
__global__ void k_float_coords(const float *__restrict__ xyz, long long n, float ox, float oy, float oz, int batch_index,
                               float full_scale, long long *__restrict__ coords, uint8_t *__restrict__ keep) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = xyz[3 * i] - ox, y = xyz[3 * i + 1] - oy, z = xyz[3 * i + 2] - oz;
  const float lo = fminf(x, fminf(y, z)), hi = fmaxf(x, fmaxf(y, z));
  const bool ok = lo >= 0.f && hi < full_scale;            // idxs = (a.min(1) >= 0) * (a.max(1) < full_scale)
  if (keep) keep[i] = ok ? 1 : 0;
  longlong2 a, b;
  a.x = (long long)x; a.y = (long long)y;                  // truncation toward zero, as LongTensor conversion does
  b.x = (long long)z; b.y = batch_index;
  reinterpret_cast<longlong2 *>(coords)[2 * i] = a;
  reinterpret_cast<longlong2 *>(coords)[2 * i + 1] = b;
}

void float_coords(const float *xyz, long long n, const float offset[3], int batch_index, float full_scale, long long *coords,
                  uint8_t *keep, cudaStream_t s) {
  if (n == 0) return;
  k_float_coords<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xyz, n, offset[0], offset[1], offset[2], batch_index, full_scale,
                                                             coords, keep);
  SCN_LAUNCH_CHECK();
}

}  // namespace scn

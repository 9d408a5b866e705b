// Exact-fp32 table convolutions (FMA on CUDA cores, fp32 accumulate): the SCN_FP32 path.
// Used for (a) parity at rel 1e-5 against the reference CPU arithmetic, (b) every layer whose channel
// counts the tensor-core path does not take (the 3->m input layer, channel counts < 32).
// Replaces dConvolution_KMxKN_forwardA_ChunkBased / backward_dI_ChunkBased / backward_dW_RuleBookBased
// and the generic forward2/backward_dW2 kernels (CUDA/Convolution.cu:77-995,1059-1152) as well as the
// Deconvolution twins (CUDA/Deconvolution.cu:9-554).
//
// Design differences from the reference:
//  * output-stationary: one CTA owns a 64-row x 64-channel output tile and walks the taps, so the
//    forward and dgrad passes need NO atomics (the reference atomically adds every output element
//    Cin/16 times, Convolution.cu:1140-1150);
//  * the table lives on the device for the whole batch (no per-call cudaMalloc/H2D/cudaFree,
//    Convolution.cu:1331-1333,1372; no 27 blocking rule uploads, :793-798);
//  * taps with no present neighbour in the tile are skipped.
#include "common.cuh"
#include <atomic>
#include <cstdlib>

namespace scn {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// FLAT   : K index runs over (tap, channel) pairs flattened -- for tiny / odd Cin (the 3->m layer), GATHER only
// !FLAT  : tap loop with tap skipping; AVEC = Cin % 4 == 0 and 16-byte aligned rows -> 128-bit gathers
// BVEC   : Cout % 4 == 0 and pointers 16-byte aligned -> 128-bit weight loads and output stores
// SCATTER: input-stationary; after each tap the tile is written to rows tbl[tap][p] (each output row has
//          exactly one producer, so plain stores suffice)
template <bool FLAT, bool AVEC, bool BVEC, bool SCATTER>
__global__ void __launch_bounds__(NT) k_table_conv(ConvArgs a) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int rows_s[BM];
  __shared__ int orow_s[BM];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Cin = a.c_in, Cout = a.c_out;

  float acc[4][4];
  auto reset = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  };
  auto store = [&](const int *orow) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = orow[ty * 4 + i];
      if (r < 0) continue;
      float *dst = a.out + (long long)r * Cout + n0 + tx * 4;
      if (BVEC) {
        if (n0 + tx * 4 < Cout) *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + tx * 4 + j < Cout) dst[j] = acc[i][j];
      }
    }
  };
  auto mac_tile = [&]() {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  };
  auto load_b = [&](const float *wk, int c0) {   // Bs[kk][*] = wk[(c0+kk)][n0 + *]
    int kk = tid >> 4, nq = (tid & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + kk < Cin) {
      const float *src = wk + (long long)(c0 + kk) * Cout + n0 + nq;
      if (BVEC) {
        if (n0 + nq < Cout) v = __ldg(reinterpret_cast<const float4 *>(src));
      } else {
        if (n0 + nq + 0 < Cout) v.x = __ldg(src + 0);
        if (n0 + nq + 1 < Cout) v.y = __ldg(src + 1);
        if (n0 + nq + 2 < Cout) v.z = __ldg(src + 2);
        if (n0 + nq + 3 < Cout) v.w = __ldg(src + 3);
      }
    }
    *reinterpret_cast<float4 *>(&Bs[kk][nq]) = v;
  };

  reset();

  if (FLAT) {
    // ---- flattened K = V*Cin, no tap skipping; weights [V][Cin][Cout] are already a [V*Cin][Cout] matrix
    const int K = a.V * Cin;
    if (tid < BM) orow_s[tid] = (row0 + tid < a.n_rows) ? row0 + tid : -1;
    for (int k0 = 0; k0 < K; k0 += BK) {
      // A: 64 rows x 16 k-values, 4 per thread; lanes run along rows so the smem stores are conflict-free
      {
        int r = tid & 63, kq = (tid >> 6) * 4;
        int orow = row0 + r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int kf = k0 + kq + j;
          float v = 0.f;
          if (kf < K && orow < a.n_rows) {
            int tap = kf / Cin, c = kf - tap * Cin;
            int trow = a.mirror ? a.V - 1 - tap : tap;
            int src = __ldg(&a.tbl[(long long)trow * a.tbl_stride + orow]);
            if (src >= 0) v = __ldg(&a.in[(long long)src * Cin + c]);
          }
          As[kq + j][r] = v;
        }
      }
      {
        int kk = tid >> 4, nq = (tid & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + kk < K) {
          const float *src = a.weight + (long long)(k0 + kk) * Cout + n0 + nq;
          if (BVEC) {
            if (n0 + nq < Cout) v = __ldg(reinterpret_cast<const float4 *>(src));
          } else {
            if (n0 + nq + 0 < Cout) v.x = __ldg(src + 0);
            if (n0 + nq + 1 < Cout) v.y = __ldg(src + 1);
            if (n0 + nq + 2 < Cout) v.z = __ldg(src + 2);
            if (n0 + nq + 3 < Cout) v.w = __ldg(src + 3);
          }
        }
        *reinterpret_cast<float4 *>(&Bs[kk][nq]) = v;
      }
      __syncthreads();
      mac_tile();
      __syncthreads();
    }
    if (a.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + tx * 4 + j < Cout) acc[i][j] += __ldg(&a.bias[n0 + tx * 4 + j]);
    }
    store(orow_s);
    return;
  }

  // ---- tap loop ---------------------------------------------------------------------------------
  for (int tap = 0; tap < a.V; ++tap) {
    const int trow = a.mirror ? a.V - 1 - tap : tap;
    int mine = -1;
    if (tid < BM) {
      int r = row0 + tid;
      int t = (r < a.n_rows) ? __ldg(&a.tbl[(long long)trow * a.tbl_stride + r]) : -1;
      if (SCATTER) {
        orow_s[tid] = t;                       // destination row of this tap
        rows_s[tid] = (r < a.n_rows) ? r : -1; // source row is the tile row itself
      } else {
        rows_s[tid] = t;
        orow_s[tid] = (r < a.n_rows) ? r : -1;
      }
      mine = t;
    }
    // block-uniform skip of taps that touch nothing in this tile (also acts as the barrier for rows_s)
    if (!__syncthreads_or(mine >= 0)) continue;

    const float *wk = a.weight + (long long)tap * Cin * Cout;
    for (int c0 = 0; c0 < Cin; c0 += BK) {
      {
        int r = tid & 63, kq = (tid >> 6) * 4;
        int src = rows_s[r];
        if (SCATTER && orow_s[r] < 0) src = -1;   // no child at this offset: contribute nothing
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) {
          const float *p = a.in + (long long)src * Cin + c0 + kq;
          if (AVEC) {
            if (c0 + kq < Cin) v = __ldg(reinterpret_cast<const float4 *>(p));
          } else {
            if (c0 + kq + 0 < Cin) v.x = __ldg(p + 0);
            if (c0 + kq + 1 < Cin) v.y = __ldg(p + 1);
            if (c0 + kq + 2 < Cin) v.z = __ldg(p + 2);
            if (c0 + kq + 3 < Cin) v.w = __ldg(p + 3);
          }
        }
        As[kq + 0][r] = v.x;
        As[kq + 1][r] = v.y;
        As[kq + 2][r] = v.z;
        As[kq + 3][r] = v.w;
      }
      load_b(wk, c0);
      __syncthreads();
      mac_tile();
      __syncthreads();
    }
    if (SCATTER) {
      store(orow_s);
      reset();
      __syncthreads();   // orow_s is rewritten by the next tap
    }
  }
  if (!SCATTER) {
    if (a.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + tx * 4 + j < Cout) acc[i][j] += __ldg(&a.bias[n0 + tx * 4 + j]);
    }
    store(orow_s);
  }
}

void conv_simt(const ConvArgs &a, cudaStream_t s) {
  if (a.n_rows == 0) return;
  dim3 grid((a.n_rows + BM - 1) / BM, (a.c_out + BN - 1) / BN);
  bool bvec = (a.c_out % 4 == 0) && ((uintptr_t)a.weight % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
  bool avec = (a.c_in % 4 == 0) && ((uintptr_t)a.in % 16 == 0);
  bool flat = !a.scatter && (a.c_in % BK != 0);
  SCN_CHECK(!(a.scatter && a.bias), "scatter convolution takes no bias");
#define SCN_DISPATCH(F, A, B, S) k_table_conv<F, A, B, S><<<grid, NT, 0, s>>>(a)
  if (a.scatter) {
    if (avec && bvec) SCN_DISPATCH(false, true, true, true);
    else if (avec) SCN_DISPATCH(false, true, false, true);
    else if (bvec) SCN_DISPATCH(false, false, true, true);
    else SCN_DISPATCH(false, false, false, true);
  } else if (flat) {
    if (bvec) SCN_DISPATCH(true, false, true, false);
    else SCN_DISPATCH(true, false, false, false);
  } else {
    if (avec && bvec) SCN_DISPATCH(false, true, true, false);
    else if (avec) SCN_DISPATCH(false, true, false, false);
    else if (bvec) SCN_DISPATCH(false, false, true, false);
    else SCN_DISPATCH(false, false, false, false);
  }
#undef SCN_DISPATCH
  SCN_LAUNCH_CHECK();
}

// -----------------------------------------------------------------------------------------------------
// per-tap transpose  dst[k][co][ci] = src[k][ci][co]
// -----------------------------------------------------------------------------------------------------
__global__ void k_transpose_weight(const float *__restrict__ src, float *__restrict__ dst, int c_in, int c_out) {
  __shared__ float tile[32][33];
  const float *s = src + (long long)blockIdx.z * c_in * c_out;
  float *d = dst + (long long)blockIdx.z * c_in * c_out;
  int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int ci = ci0 + i, co = co0 + threadIdx.x;
    tile[i][threadIdx.x] = (ci < c_in && co < c_out) ? s[(long long)ci * c_out + co] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int co = co0 + i, ci = ci0 + threadIdx.x;
    if (co < c_out && ci < c_in) d[(long long)co * c_in + ci] = tile[threadIdx.x][i];
  }
}

// the same with the bf16 rounding of the tensor-core operand folded in: dst[k][co][ci] = bf16(src[k][ci][co])
__global__ void k_transpose_weight_bf16(const float *__restrict__ src, uint16_t *__restrict__ dst, int c_in, int c_out) {
  __shared__ float tile[32][33];
  const float *s = src + (long long)blockIdx.z * c_in * c_out;
  uint16_t *d = dst + (long long)blockIdx.z * c_in * c_out;
  int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int ci = ci0 + i, co = co0 + threadIdx.x;
    tile[i][threadIdx.x] = (ci < c_in && co < c_out) ? s[(long long)ci * c_out + co] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int co = co0 + i, ci = ci0 + threadIdx.x;
    if (co < c_out && ci < c_in) {
      uint32_t r;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(tile[threadIdx.x][i]));
      d[(long long)co * c_in + ci] = (uint16_t)(r & 0xFFFFu);
    }
  }
}

void transpose_weight_bf16(const float *src, uint16_t *dst, int V, int c_in, int c_out, cudaStream_t s) {
  dim3 grid((c_out + 31) / 32, (c_in + 31) / 32, V);
  k_transpose_weight_bf16<<<grid, dim3(32, 8), 0, s>>>(src, dst, c_in, c_out);
  SCN_LAUNCH_CHECK();
}

void transpose_weight(const float *src, float *dst, int V, int c_in, int c_out, cudaStream_t s) {
  dim3 grid((c_out + 31) / 32, (c_in + 31) / 32, V);
  k_transpose_weight<<<grid, dim3(32, 8), 0, s>>>(src, dst, c_in, c_out);
  SCN_LAUNCH_CHECK();
}

static std::atomic<int> g_deterministic{[] {
  const char *e = getenv("SCN_DETERMINISTIC");
  return e ? (atoi(e) != 0 ? 1 : 0) : 0;
}()};
bool deterministic() { return g_deterministic.load(std::memory_order_relaxed) != 0; }
int set_deterministic(int on) { return g_deterministic.exchange(on ? 1 : 0); }

// Fixed summation tree: 16 part-lanes per element each add every 16th partial in ascending order, then the 16 lane sums are
// added in lane order -- the same order on every run, without a 600-step dependent chain per element (the column statistics of a
// convolution are 128 elements x 592 partials).
constexpr int SP_LANES = 16;
template <typename T>
__global__ void __launch_bounds__(32 * SP_LANES) k_sum_partials(const T *__restrict__ partial, int parts, long long n,
                                                               T *__restrict__ out) {
  __shared__ T lane_sum[SP_LANES][33];
  const long long e = blockIdx.x * 32ll + threadIdx.x;
  T acc = T(0);
  if (e < n)
    for (int p = threadIdx.y; p < parts; p += SP_LANES) acc += partial[(long long)p * n + e];
  lane_sum[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && e < n) {
    T t = lane_sum[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < SP_LANES; ++y) t += lane_sum[y][threadIdx.x];
    out[e] = t;
  }
}
// many elements (weight gradients): one thread per element walks the partials in order; the threads supply the parallelism
template <typename T>
__global__ void k_sum_partials_wide(const T *__restrict__ partial, int parts, long long n, T *__restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    T acc = partial[e];
    for (int p = 1; p < parts; ++p) acc += partial[(long long)p * n + e];
    out[e] = acc;
  }
}
template <typename T> static void sum_partials_t(const T *partial, int parts, long long n, T *out, cudaStream_t s) {
  if (n == 0 || parts == 0) return;
  if (n >= 16384) {        // (the choice depends on sizes only, so a given layer always takes the same tree)
    long long g = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    k_sum_partials_wide<T><<<(int)(g > cap ? cap : g), 256, 0, s>>>(partial, parts, n, out);
  } else {
    k_sum_partials<T><<<(unsigned)((n + 31) / 32), dim3(32, SP_LANES), 0, s>>>(partial, parts, n, out);
  }
  SCN_LAUNCH_CHECK();
}
void sum_partials(const float *partial, int parts, long long n, float *out, cudaStream_t s) { sum_partials_t(partial, parts, n, out, s); }
void sum_partials(const double *partial, int parts, long long n, double *out, cudaStream_t s) { sum_partials_t(partial, parts, n, out, s); }

// y += x
__global__ void k_axpy(const float *__restrict__ x, float *__restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
void axpy(const float *x, float *y, long long n, cudaStream_t s) {
  if (n == 0) return;
  long long g = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  k_axpy<<<(int)(g > cap ? cap : g), 256, 0, s>>>(x, y, n);
  SCN_LAUNCH_CHECK();
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread
__global__ void k_cast_bf16(const float *__restrict__ src, uint16_t *__restrict__ dst, long long n8, long long n) {
  auto cvt2 = [](float lo, float hi) -> uint32_t {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4 *>(src) + 2 * i + 1);
    uint4 o;
    o.x = cvt2(a.x, a.y); o.y = cvt2(a.z, a.w); o.z = cvt2(b.x, b.y); o.w = cvt2(b.z, b.w);
    reinterpret_cast<uint4 *>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n8 * 8) {       // tail
    const long long i = n8 * 8 + threadIdx.x;
    uint32_t r = cvt2(src[i], 0.f);
    dst[i] = (uint16_t)(r & 0xFFFFu);
  }
}

// rows of `cols` floats `ld` floats apart -> dense bf16 [rows, cols]  (cols % 8 == 0)
__global__ void k_cast_bf16_rows(const float *__restrict__ src, long long ld, int cols8, long long n8,
                                 uint16_t *__restrict__ dst) {
  auto cvt2 = [](float lo, float hi) -> uint32_t {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols8;
    const int c = (int)(i - r * cols8);
    const float4 *p = reinterpret_cast<const float4 *>(src + r * ld) + 2 * c;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    uint4 o;
    o.x = cvt2(a.x, a.y); o.y = cvt2(a.z, a.w); o.z = cvt2(b.x, b.y); o.w = cvt2(b.z, b.w);
    reinterpret_cast<uint4 *>(dst)[i] = o;
  }
}

void cast_bf16_rows(const float *src, long long ld, long long rows, int cols, uint16_t *dst, cudaStream_t s) {
  if (ld == cols || ld == 0) { cast_bf16(src, dst, rows * cols, s); return; }
  if (rows == 0) return;
  SCN_CHECK(cols % 8 == 0 && ld % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0,
            "cast_bf16_rows: strided rows need cols % 8 == 0 and 16-byte aligned rows");
  const long long n8 = rows * (cols / 8);
  long long g = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  k_cast_bf16_rows<<<(int)g, 256, 0, s>>>(src, ld, cols / 8, n8, dst);
  SCN_LAUNCH_CHECK();
}

void cast_bf16(const float *src, uint16_t *dst, long long n, cudaStream_t s) {
  if (n == 0) return;
  SCN_CHECK((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0, "cast_bf16: unaligned buffers");
  const long long n8 = n / 8;
  long long g = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  k_cast_bf16<<<(int)g, 256, 0, s>>>(src, dst, n8, n);
  SCN_LAUNCH_CHECK();
}

// -----------------------------------------------------------------------------------------------------
// weight gradient  dW[k] = sum_r A[ia(k,r)]^T B[ib(k,r)]   (split over row chunks, fp32 atomics to merge)
// -----------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NT) k_wgrad(WgradArgs a, int rows_per_cta, int tiles_b) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int ia_s[BK], ib_s[BK];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tap = blockIdx.y;
  const int ca0 = (blockIdx.z / tiles_b) * BM, cb0 = (blockIdx.z % tiles_b) * BN;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(r_begin + rows_per_cta, a.n_rows);
  const int *trow = a.tbl + (long long)tap * a.tbl_stride;
  const int Ca = a.c_a, Cb = a.c_b;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  bool any_work = false;

  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    int t = -1;
    if (tid < BK) {
      int r = r0 + tid;
      t = (r < r_end) ? __ldg(&trow[r]) : -1;
      int other = (t >= 0) ? r : -1;
      ia_s[tid] = a.table_on_a ? t : other;
      ib_s[tid] = a.table_on_a ? other : t;
    }
    if (!__syncthreads_or(t >= 0)) continue;   // 16-row slab with no rule at this tap
    any_work = true;
    {
      int kk = tid >> 4, cq = (tid & 15) * 4;
      int ra = ia_s[kk], rb = ib_s[kk];
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (ra >= 0) {
        const float *pa = a.a + (long long)ra * Ca + ca0 + cq;
        const float *pb = a.b + (long long)rb * Cb + cb0 + cq;
        if (VEC) {
          if (ca0 + cq < Ca) va = __ldg(reinterpret_cast<const float4 *>(pa));
          if (cb0 + cq < Cb) vb = __ldg(reinterpret_cast<const float4 *>(pb));
        } else {
          if (ca0 + cq + 0 < Ca) va.x = __ldg(pa + 0);
          if (ca0 + cq + 1 < Ca) va.y = __ldg(pa + 1);
          if (ca0 + cq + 2 < Ca) va.z = __ldg(pa + 2);
          if (ca0 + cq + 3 < Ca) va.w = __ldg(pa + 3);
          if (cb0 + cq + 0 < Cb) vb.x = __ldg(pb + 0);
          if (cb0 + cq + 1 < Cb) vb.y = __ldg(pb + 1);
          if (cb0 + cq + 2 < Cb) vb.z = __ldg(pb + 2);
          if (cb0 + cq + 3 < Cb) vb.w = __ldg(pb + 3);
        }
      }
      *reinterpret_cast<float4 *>(&As[kk][cq]) = va;
      *reinterpret_cast<float4 *>(&Bs[kk][cq]) = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (!any_work) return;
  float *dw = a.dw + blockIdx.x * a.part_stride + (long long)tap * Ca * Cb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ca = ca0 + ty * 4 + i;
    if (ca >= Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cb = cb0 + tx * 4 + j;
      if (cb < Cb) atomicAdd(&dw[(long long)ca * Cb + cb], acc[i][j]);
    }
  }
}

void wgrad_simt(const WgradArgs &a_in, cudaStream_t s) {
  WgradArgs a = a_in;
  const size_t n_dw = (size_t)a.V * a.c_a * a.c_b;
  SCN_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * n_dw, s));
  if (a.n_rows == 0) return;
  int tiles_a = (a.c_a + BM - 1) / BM, tiles_b = (a.c_b + BN - 1) / BN;
  // enough row chunks to fill the machine a few times over, but at least 256 rows each so the final
  // atomics stay a small fraction of the work
  long long target = (long long)sm_count() * 8 / ((long long)a.V * tiles_a * tiles_b) + 1;
  int rows_per_cta = (int)((a.n_rows + target - 1) / target);
  rows_per_cta = ((rows_per_cta < 256 ? 256 : rows_per_cta) + BK - 1) / BK * BK;
  dim3 grid((a.n_rows + rows_per_cta - 1) / rows_per_cta, a.V, tiles_a * tiles_b);
  bool vec = (a.c_a % 4 == 0) && (a.c_b % 4 == 0) && ((uintptr_t)a.a % 16 == 0) && ((uintptr_t)a.b % 16 == 0);
  DevBuf<float> part;
  if (deterministic() && grid.x > 1) {      // one zeroed slice per row chunk: a single atomic per address, summed in chunk order below
    part.alloc(n_dw * grid.x, s);
    SCN_CUDA(cudaMemsetAsync(part.p, 0, sizeof(float) * n_dw * grid.x, s));
    a.dw = part.p;
    a.part_stride = (long long)n_dw;
  }
  if (vec) k_wgrad<true><<<grid, NT, 0, s>>>(a, rows_per_cta, tiles_b);
  else k_wgrad<false><<<grid, NT, 0, s>>>(a, rows_per_cta, tiles_b);
  SCN_LAUNCH_CHECK();
  if (part.p) sum_partials(part.p, (int)grid.x, (long long)n_dw, a_in.dw, s);
  part.release(s);
}

// -----------------------------------------------------------------------------------------------------
// bias gradient (column sums).  Replaces Convolution_bp_bias_ (Convolution.cu:58-75).  Unused by the UNet
// (bias=False everywhere) but part of the entry points' contract.
// -----------------------------------------------------------------------------------------------------
__global__ void k_bias_grad(const float *__restrict__ d_out, float *__restrict__ d_bias, long long n, int C) {
  int c = blockIdx.x * 32 + threadIdx.x;
  __shared__ double part[8][33];
  double acc = 0.0;
  if (c < C)
    for (long long r = threadIdx.y; r < n; r += 8) acc += (double)d_out[r * C + c];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    d_bias[c] = (float)t;
  }
}

void bias_grad(const float *d_out, float *d_bias, long long n_rows, int C, cudaStream_t s) {
  k_bias_grad<<<(C + 31) / 32, dim3(32, 8), 0, s>>>(d_out, d_bias, n_rows, C);
  SCN_LAUNCH_CHECK();
}

}  // namespace scn

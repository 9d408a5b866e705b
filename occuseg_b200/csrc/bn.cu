// BatchNorm + (leaky) ReLU, forward and backward.  Replaces BatchNormalization_f_train / f_test / _b
// (CUDA/BatchNormalization.cu:14-199, dispatch :201-238; drivers CUDA/BatchNormalization.cpp:21-71).
//
// The reference walks all N rows with at most 16 CTAs (grid = min(16, C/NTX)) and accumulates the
// statistics as fp32 running sums.  Here the reduction is spread over 4 CTAs per SM, partial sums are
// short fp32 chains merged in fp64, and the normalise pass is a 128-bit streaming kernel.
// HBM-bound: forward 3*N*C*4 bytes (read x twice, write y), backward 5*N*C*4 (x and d_out twice, d_in once).
#include "common.cuh"

namespace scn {

constexpr int RED_THREADS = 256;
constexpr int ROWS_PER_FLUSH = 64;      // rows a thread folds in fp32 before adding to its fp64 partial

// column sums of u(x) and v(x):  MODE 0: (x, x*x)          (forward statistics)
//                                MODE 1: (d', (x-mean)*d')  with d' = d * (y>0 ? 1 : leak)   (backward)
// Thread t owns column group t % cv (VEC columns) and row lane t / cv, so every lane of the block is busy for any
// C (the block uses cv * floor(256/cv) threads) and a warp reads whole contiguous rows.  Four independent rows
// are in flight per thread.  fp32 chains of <= 64 rows are merged in fp64 (registers -> shared -> one fp64
// atomic per column per CTA).
// The backward passes do not read the forward output: the ReLU mask y > 0 is recomputed from x with the very
// expression the forward pass used (fmaf(w, x, b), w = invstd*gamma, b = beta - mean*w), which saves one of the three
// [N, C] reads of each backward kernel.
template <int VEC, int MODE>
__global__ void __launch_bounds__(RED_THREADS) k_bn_reduce(const float *__restrict__ x, const float *__restrict__ invstd,
                                                          const float *__restrict__ d, const float *__restrict__ mean,
                                                          const float *__restrict__ gamma, const float *__restrict__ beta,
                                                          long long n, int C, float leak, double *__restrict__ acc,
                                                          long long part_stride) {
  const int cv = C / VEC;
  const int row_lanes = blockDim.x / cv;
  const int cg = threadIdx.x % cv, rl = threadIdx.x / cv;
  extern __shared__ double red[];          // [2][row_lanes][C]
  float m[VEC], wv[VEC], bv[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    m[j] = wv[j] = bv[j] = 0.f;
    if (MODE == 1) {
      const int c = cg * VEC + j;
      m[j] = __ldg(&mean[c]);
      wv[j] = __ldg(&invstd[c]) * (gamma ? __ldg(&gamma[c]) : 1.f);
      bv[j] = -m[j] * wv[j] + (beta ? __ldg(&beta[c]) : 0.f);
    }
  }
  double t0[VEC], t1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) t0[j] = t1[j] = 0.0;
  const long long rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  const long long r_begin = blockIdx.x * rows_per_cta;
  const long long r_end = r_begin + rows_per_cta < n ? r_begin + rows_per_cta : n;
  auto fold = [&](long long r, float *s0, float *s1) {
    float xv[VEC], dv[VEC];
    if (VEC == 4) {
      float4 t = __ldg(reinterpret_cast<const float4 *>(x + r * C) + cg);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      if (MODE == 1) {
        float4 w = __ldg(reinterpret_cast<const float4 *>(d + r * C) + cg);
        dv[0] = w.x; dv[1] = w.y; dv[2] = w.z; dv[3] = w.w;
      }
    } else {
      xv[0] = __ldg(x + r * C + cg);
      if (MODE == 1) dv[0] = __ldg(d + r * C + cg);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (MODE == 0) {
        s0[j] += xv[j];
        s1[j] = fmaf(xv[j], xv[j], s1[j]);
      } else {
        float dd = fmaf(wv[j], xv[j], bv[j]) > 0.f ? dv[j] : dv[j] * leak;
        s0[j] += dd;
        s1[j] = fmaf(xv[j] - m[j], dd, s1[j]);
      }
    }
  };
  constexpr int U = MODE == 0 ? 8 : 4;      // independent rows in flight per thread (8 x 16-byte loads either way)
  for (long long rb = r_begin + rl; rb < r_end; rb += (long long)row_lanes * ROWS_PER_FLUSH) {
    float s0[U][VEC], s1[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) s0[u][j] = s1[u][j] = 0.f;
    long long r = rb;
    const long long stop = rb + (long long)row_lanes * ROWS_PER_FLUSH < r_end ? rb + (long long)row_lanes * ROWS_PER_FLUSH : r_end;
    for (; r + (long long)(U - 1) * row_lanes < stop; r += (long long)U * row_lanes) {
#pragma unroll
      for (int u = 0; u < U; ++u) fold(r + (long long)u * row_lanes, s0[u], s1[u]);
    }
    for (; r < stop; r += row_lanes) fold(r, s0[0], s1[0]);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) { a0 += s0[u][j]; a1 += s1[u][j]; }
      t0[j] += (double)a0;
      t1[j] += (double)a1;
    }
  }
  if (threadIdx.x < cv * row_lanes) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      red[(0 * row_lanes + rl) * C + cg * VEC + j] = t0[j];
      red[(1 * row_lanes + rl) * C + cg * VEC + j] = t1[j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += blockDim.x) {
    const int which = e / C, c = e - which * C;
    double t = 0.0;
    for (int i = 0; i < row_lanes; ++i) t += red[(which * row_lanes + i) * C + c];
    atomicAdd(&acc[blockIdx.x * part_stride + which * C + c], t);
  }
}

struct RedCfg { int threads, grid; size_t smem; };
static RedCfg reduce_cfg(long long n, int C, int vec) {
  const int cv = C / vec;
  RedCfg c;
  const int row_lanes = cv >= RED_THREADS ? 1 : RED_THREADS / cv;
  c.threads = cv >= RED_THREADS ? RED_THREADS : cv * row_lanes;
  c.smem = sizeof(double) * 2 * (size_t)row_lanes * C;
  long long g = (n + (long long)row_lanes * 16 - 1) / ((long long)row_lanes * 16);
  long long cap = (long long)sm_count() * 4;
  c.grid = (int)(g < 1 ? 1 : g > cap ? cap : g);
  return c;
}

// forward finalize: mean / invstd / running statistics.  Formulas follow BatchNormalization.cu:38-51
// (biased variance for normalisation, unbiased for the running estimate), evaluated in fp64.
__global__ void k_bn_finalize_fwd(const double *__restrict__ acc, long long n, int C, float eps, float momentum,
                                  bool train, float *__restrict__ save_mean, float *__restrict__ save_invstd,
                                  float *__restrict__ running_mean, float *__restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (train) {
    double mean = acc[c] / (double)n;
    double var_sum = acc[C + c] - mean * mean * (double)n;
    if (var_sum < 0.0) var_sum = 0.0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var_sum / (double)n + (double)eps));
    running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * (float)mean;
    running_var[c] = momentum * running_var[c] + (1.f - momentum) * (float)(var_sum / (double)(n - 1));
  } else {
    save_mean[c] = running_mean[c];
    save_invstd[c] = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) k_bn_apply_fwd(const float *__restrict__ x, float *__restrict__ y,
                                                      uint16_t *__restrict__ y16, const float *__restrict__ save_mean,
                                                      const float *__restrict__ save_invstd,
                                                      const float *__restrict__ gamma, const float *__restrict__ beta,
                                                      long long n, int C, float leak) {
  extern __shared__ float wb[];   // [2][C]: w = invstd*gamma, b = beta - mean*w  (BatchNormalization.cu:56-60)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float w = save_invstd[c] * (gamma ? gamma[c] : 1.f);
    wb[c] = w;
    wb[C + c] = -save_mean[c] * w + (beta ? beta[c] : 0.f);
  }
  __syncthreads();
  const int cv = C / VEC;
  const long long total = n * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC == 4) {
    constexpr int U = 2;           // two elements in flight per thread (see k_bn_apply_bwd)
    for (long long e0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; e0 < total; e0 += stride * U) {
      float4 t[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (e0 + u * stride < total) t[u] = __ldg(reinterpret_cast<const float4 *>(x) + e0 + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = e0 + u * stride;
        if (e >= total) continue;
        const int cg = (int)(e % cv);
        const float4 w = *reinterpret_cast<const float4 *>(&wb[cg * 4]);
        const float4 b = *reinterpret_cast<const float4 *>(&wb[C + cg * 4]);
        float4 o;
        o.x = fmaf(w.x, t[u].x, b.x); o.y = fmaf(w.y, t[u].y, b.y); o.z = fmaf(w.z, t[u].z, b.z); o.w = fmaf(w.w, t[u].w, b.w);
        o.x = o.x > 0.f ? o.x : o.x * leak; o.y = o.y > 0.f ? o.y : o.y * leak;
        o.z = o.z > 0.f ? o.z : o.z * leak; o.w = o.w > 0.f ? o.w : o.w * leak;
        if (y) reinterpret_cast<float4 *>(y)[e] = o;
        if (y16) {      // bf16 copy for the tensor-core convolution that consumes y (saves that layer's cast pass)
          uint2 h;
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(o.y), "f"(o.x));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(o.w), "f"(o.z));
          reinterpret_cast<uint2 *>(y16)[e] = h;
        }
      }
    }
  } else {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
      const int cg = (int)(e % cv);
      float o = fmaf(wb[cg], __ldg(x + e), wb[C + cg]);
      y[e] = o > 0.f ? o : o * leak;
    }
  }
}

// backward finalize: d_gamma = dotp*invstd, d_beta = sum d'; coef[0][c] = mean(d'), coef[1][c] = dotp*invstd^2/N
// (BatchNormalization.cu:160-170)
// raw_mean != NULL: acc[1] holds sum d'*x (not centred); dotp = acc[1] - mean * acc[0]
__global__ void k_bn_finalize_bwd(const double *__restrict__ acc, long long n, int C,
                                  const float *__restrict__ save_invstd, float *__restrict__ d_gamma,
                                  float *__restrict__ d_beta, float *__restrict__ coef, const float *__restrict__ raw_mean) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double gsum = acc[c], dotp = acc[C + c], is = (double)save_invstd[c];
  if (raw_mean) dotp -= (double)raw_mean[c] * gsum;
  if (d_gamma) d_gamma[c] = (float)(dotp * is);
  if (d_beta) d_beta[c] = (float)gsum;
  coef[c] = (float)(gsum / (double)n);
  coef[C + c] = (float)(dotp * is * is / (double)n);
}

template <int VEC>
__global__ void __launch_bounds__(256) k_bn_apply_bwd(const float *__restrict__ x, const float *__restrict__ beta,
                                                      const float *__restrict__ d, const float *__restrict__ add,
                                                      float *__restrict__ dx, uint16_t *__restrict__ dx16,
                                                      const float *__restrict__ save_mean,
                                                      const float *__restrict__ save_invstd,
                                                      const float *__restrict__ gamma, const float *__restrict__ coef,
                                                      long long n, int C, float leak, bool premasked, long long ld_add) {
  extern __shared__ float sm[];   // [5][C]: mean, gradMean, k, w = invstd*gamma, b = beta - mean*w
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float w = save_invstd[c] * (gamma ? gamma[c] : 1.f);
    sm[c] = save_mean[c];
    sm[C + c] = coef[c];
    sm[2 * C + c] = coef[C + c];
    sm[3 * C + c] = w;
    sm[4 * C + c] = -save_mean[c] * w + (beta ? beta[c] : 0.f);
  }
  __syncthreads();
  const int cv = C / VEC;
  const long long total = n * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // U elements per thread and trip, every load issued before the first use: the pass is latency-bound on its 2-3 loads per
  // element (ncu: long-scoreboard stalls, 0.79 of the copy peak with one element in flight per thread)
  constexpr int U = VEC == 4 ? 2 : 1;
  for (long long e0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; e0 < total; e0 += stride * U) {
    float xv[U][VEC], dv[U][VEC], av[U][VEC];
    int c0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * stride;
      c0[u] = 0;
      if (e >= total) continue;
      c0[u] = (int)(e % cv) * VEC;
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(x) + e);
        const float4 w = __ldg(reinterpret_cast<const float4 *>(d) + e);
        xv[u][0] = t.x; xv[u][1] = t.y; xv[u][2] = t.z; xv[u][3] = t.w;
        dv[u][0] = w.x; dv[u][1] = w.y; dv[u][2] = w.z; dv[u][3] = w.w;
      } else {
        xv[u][0] = __ldg(x + e); dv[u][0] = __ldg(d + e);
      }
      if (add) {          // gradient arriving through the residual shortcut of the same input: accumulated here
        const long long ea = ld_add == C ? e * VEC : (e / cv) * ld_add + c0[u];     // rows of `add` may be ld_add floats apart
        if (VEC == 4) {
          const float4 a4 = __ldg(reinterpret_cast<const float4 *>(add + ea));
          av[u][0] = a4.x; av[u][1] = a4.y; av[u][2] = a4.z; av[u][3] = a4.w;
        } else {
          av[u][0] = __ldg(add + ea);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * stride;
      if (e >= total) continue;
      float ov[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int c = c0[u] + j;
        const float dd = (premasked || fmaf(sm[3 * C + c], xv[u][j], sm[4 * C + c]) > 0.f) ? dv[u][j] : dv[u][j] * leak;
        ov[j] = (dd - sm[C + c] - (xv[u][j] - sm[c]) * sm[2 * C + c]) * sm[3 * C + c];
        if (add) ov[j] += av[u][j];
      }
      if (VEC == 4) reinterpret_cast<float4 *>(dx)[e] = make_float4(ov[0], ov[1], ov[2], ov[3]);
      else dx[e] = ov[0];
      if (VEC == 4 && dx16) {      // bf16 copy for the backward products of the convolution that receives dx as its d_out
        uint2 h;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(ov[1]), "f"(ov[0]));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(ov[3]), "f"(ov[2]));
        reinterpret_cast<uint2 *>(dx16)[e] = h;
      }
    }
  }
}

__global__ void k_bn_eval_coeffs(const float *__restrict__ running_mean, const float *__restrict__ running_var,
                                 const float *__restrict__ gamma, const float *__restrict__ beta, int C, float eps,
                                 float *__restrict__ scale, float *__restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // the same two roundings as k_bn_finalize_fwd (fp64 rsqrt -> float) and k_bn_apply_fwd's prologue
  const float invstd = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  const float w = invstd * (gamma ? gamma[c] : 1.f);
  scale[c] = w;
  shift[c] = -running_mean[c] * w + (beta ? beta[c] : 0.f);
}

void bn_eval_coeffs(const float *running_mean, const float *running_var, const float *gamma, const float *beta, int C,
                    float eps, float *scale, float *shift, cudaStream_t s) {
  k_bn_eval_coeffs<<<(C + 127) / 128, 128, 0, s>>>(running_mean, running_var, gamma, beta, C, eps, scale, shift);
  SCN_LAUNCH_CHECK();
}

static int stream_grid(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  long long cap = (long long)sm_count() * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

void bn_fwd(const float *in, float *out, uint16_t *out_bf16, const double *stats_in, float *save_mean, float *save_invstd,
            float *running_mean,
            float *running_var, const float *gamma, const float *beta, long long n, int C, float eps, float momentum,
            bool train, float leakiness, cudaStream_t s) {
  SCN_CHECK(C > 0 && C <= 4096, "BatchNorm: channel count out of range");
  if (n == 0) return;
  bool v4 = (C % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
  DevBuf<double> acc;
  acc.alloc(2 * (size_t)C, s);
  const double *acc_use = acc.p;
  if (train) SCN_CHECK(n > 1, "BatchNorm (train): needs at least two active rows");   // unbiased running variance divides by n-1
  if (train && stats_in) {
    acc_use = stats_in;          // column sums already accumulated by the kernel that produced `in`
  } else if (train) {
    SCN_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * 2 * C, s));
    const bool r4 = v4 && C / 4 <= RED_THREADS;
    SCN_CHECK(r4 || C <= RED_THREADS, "BatchNorm: more than 256 channels need 16-byte aligned rows");
    const RedCfg rc = reduce_cfg(n, C, r4 ? 4 : 1);
    DevBuf<double> part;              // deterministic mode: one zeroed slice per CTA, summed in CTA order
    double *dst = acc.p;
    long long ps = 0;
    if (deterministic() && rc.grid > 1) {
      part.alloc((size_t)rc.grid * 2 * C, s);
      SCN_CUDA(cudaMemsetAsync(part.p, 0, sizeof(double) * part.n, s));
      dst = part.p;
      ps = 2ll * C;
    }
    if (r4) k_bn_reduce<4, 0><<<rc.grid, rc.threads, rc.smem, s>>>(in, nullptr, nullptr, nullptr, nullptr, nullptr, n, C, 0.f, dst, ps);
    else k_bn_reduce<1, 0><<<rc.grid, rc.threads, rc.smem, s>>>(in, nullptr, nullptr, nullptr, nullptr, nullptr, n, C, 0.f, dst, ps);
    SCN_LAUNCH_CHECK();
    if (part.p) sum_partials(part.p, rc.grid, 2ll * C, acc.p, s);
    part.release(s);
  }
  k_bn_finalize_fwd<<<(C + 127) / 128, 128, 0, s>>>(acc_use, n, C, eps, momentum, train, save_mean, save_invstd,
                                                    running_mean, running_var);
  SCN_LAUNCH_CHECK();
  size_t smem = sizeof(float) * 2 * C;
  SCN_CHECK(!out_bf16 || (v4 && (uintptr_t)out_bf16 % 8 == 0), "BatchNorm: the bf16 copy needs C % 4 == 0 and aligned buffers");
  SCN_CHECK(out || (out_bf16 && v4), "BatchNorm: no output buffer (a bf16-only output needs C % 4 == 0)");
  if (v4) k_bn_apply_fwd<4><<<stream_grid(n * (C / 4), 256), 256, smem, s>>>(in, out, out_bf16, save_mean, save_invstd, gamma, beta, n, C, leakiness);
  else k_bn_apply_fwd<1><<<stream_grid(n * C, 256), 256, smem, s>>>(in, out, nullptr, save_mean, save_invstd, gamma, beta, n, C, leakiness);
  SCN_LAUNCH_CHECK();
  acc.release(s);
}

void bn_bwd(const float *in, const float *out, const float *d_out, const float *save_mean, const float *save_invstd,
            const float *gamma, const float *beta, const float *d_in_add, float *d_in, float *d_gamma, float *d_beta, long long n, int C, float leakiness,
            cudaStream_t s) {
  SCN_CHECK(C > 0 && C <= 4096, "BatchNorm: channel count out of range");
  if (n == 0) return;
  bool v4 = (C % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && ((uintptr_t)d_in % 16 == 0) &&
            ((uintptr_t)d_in_add % 16 == 0);
  DevBuf<double> acc;
  DevBuf<float> coef;
  acc.alloc(2 * (size_t)C, s);
  coef.alloc(2 * (size_t)C, s);
  SCN_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * 2 * C, s));
  const bool r4 = v4 && C / 4 <= RED_THREADS;
  SCN_CHECK(r4 || C <= RED_THREADS, "BatchNorm: more than 256 channels need 16-byte aligned rows");
  const RedCfg rc = reduce_cfg(n, C, r4 ? 4 : 1);
  DevBuf<double> part;                // deterministic mode: see bn_fwd
  double *dst = acc.p;
  long long ps = 0;
  if (deterministic() && rc.grid > 1) {
    part.alloc((size_t)rc.grid * 2 * C, s);
    SCN_CUDA(cudaMemsetAsync(part.p, 0, sizeof(double) * part.n, s));
    dst = part.p;
    ps = 2ll * C;
  }
  if (r4) k_bn_reduce<4, 1><<<rc.grid, rc.threads, rc.smem, s>>>(in, save_invstd, d_out, save_mean, gamma, beta, n, C, leakiness, dst, ps);
  else k_bn_reduce<1, 1><<<rc.grid, rc.threads, rc.smem, s>>>(in, save_invstd, d_out, save_mean, gamma, beta, n, C, leakiness, dst, ps);
  SCN_LAUNCH_CHECK();
  if (part.p) sum_partials(part.p, rc.grid, 2ll * C, acc.p, s);
  part.release(s);
  k_bn_finalize_bwd<<<(C + 127) / 128, 128, 0, s>>>(acc.p, n, C, save_invstd, d_gamma, d_beta, coef.p, nullptr);
  SCN_LAUNCH_CHECK();
  size_t smem = sizeof(float) * 5 * C;
  if (v4) k_bn_apply_bwd<4><<<stream_grid(n * (C / 4), 256), 256, smem, s>>>(in, beta, d_out, d_in_add, d_in, nullptr, save_mean, save_invstd, gamma, coef.p, n, C, leakiness, false, C);
  else k_bn_apply_bwd<1><<<stream_grid(n * C, 256), 256, smem, s>>>(in, beta, d_out, d_in_add, d_in, nullptr, save_mean, save_invstd, gamma, coef.p, n, C, leakiness, false, C);
  SCN_LAUNCH_CHECK();
  acc.release(s);
  coef.release(s);
}

__global__ void k_bn_mask_coeffs(const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                                 const float *__restrict__ gamma, const float *__restrict__ beta, int C, float *__restrict__ coef) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float w = save_invstd[c] * (gamma ? gamma[c] : 1.f);      // the very expressions of k_bn_apply_fwd's prologue
  coef[c] = w;
  coef[C + c] = -save_mean[c] * w + (beta ? beta[c] : 0.f);
}

void bn_mask_coeffs(const float *save_mean, const float *save_invstd, const float *gamma, const float *beta, int C, float *coef,
                    cudaStream_t s) {
  k_bn_mask_coeffs<<<(C + 127) / 128, 128, 0, s>>>(save_mean, save_invstd, gamma, beta, C, coef);
  SCN_LAUNCH_CHECK();
}

void bn_bwd_apply(const float *in, const float *d_masked, const double *acc, const float *save_mean, const float *save_invstd,
                  const float *gamma, const float *d_in_add, long long ld_add, float *d_in, uint16_t *d_in_bf16, float *d_gamma,
                  float *d_beta, long long n, int C, cudaStream_t s) {
  if (!d_in_add || ld_add == 0) ld_add = C;
  SCN_CHECK(ld_add >= C && (ld_add == C || ld_add % 4 == 0), "BatchNorm: bad row stride of d_in_add");
  SCN_CHECK(C > 0 && C <= 4096, "BatchNorm: channel count out of range");
  if (n == 0) return;
  const bool v4 = (C % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)d_masked % 16 == 0) && ((uintptr_t)d_in % 16 == 0) &&
                  ((uintptr_t)d_in_add % 16 == 0);
  SCN_CHECK(!d_in_bf16 || (v4 && (uintptr_t)d_in_bf16 % 8 == 0), "BatchNorm: the bf16 gradient copy needs 16-byte aligned rows");
  DevBuf<float> coef;
  coef.alloc(2 * (size_t)C, s);
  k_bn_finalize_bwd<<<(C + 127) / 128, 128, 0, s>>>(acc, n, C, save_invstd, d_gamma, d_beta, coef.p, save_mean);
  SCN_LAUNCH_CHECK();
  const size_t smem = sizeof(float) * 5 * C;
  if (v4) k_bn_apply_bwd<4><<<stream_grid(n * (C / 4), 256), 256, smem, s>>>(in, nullptr, d_masked, d_in_add, d_in, d_in_bf16, save_mean, save_invstd, gamma, coef.p, n, C, 0.f, true, ld_add);
  else k_bn_apply_bwd<1><<<stream_grid(n * C, 256), 256, smem, s>>>(in, nullptr, d_masked, d_in_add, d_in, d_in_bf16, save_mean, save_invstd, gamma, coef.p, n, C, 0.f, true, ld_add);
  SCN_LAUNCH_CHECK();
  coef.release(s);
}

}  // namespace scn

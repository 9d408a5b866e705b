// Table convolutions with a tiny input-channel count (the 3 -> m layer in front of the UNet,
// examples/ScanNet/model.py:663): exact fp32 FMA, one thread per output row.
//
// With Cin = 3 the whole weight tensor is 27*3*Cout floats (20 KB for Cout = 64): it sits in shared memory and
// is read with warp-broadcast 128-bit loads, the 27 neighbour rows cost 12 bytes each (the [N,3] input stays in
// L2), and every thread keeps its 64 output channels in registers -- so the kernel is bound by the FMA pipe
// (81*Cout FMAs per row), not by the 256-byte output row it writes once.  The generic 64x64 tile kernel
// (conv_simt.cu) spends the same FMAs but re-stages A and B tiles through shared memory 6 times per tile and
// leaves 13/16 of each K step empty; the reference does the same work with atomics (Convolution.cu:1059-1152).
// The weight gradient of that layer is the same contraction transposed: dW[k][ci][co] = sum_o x[nbr_k(o)][ci] *
// g[o][co]; a CTA stages 32 rows of neighbour inputs (zeros where absent) and g in shared memory and every thread
// owns (one output channel) x (a quarter of the 27 taps), i.e. 7*Cin accumulators, merged with fp32 atomics.
#include "common.cuh"

namespace scn {

constexpr int SC_ROWS = 128;      // rows (threads) per CTA, forward
constexpr int SC_TILE = 64;       // output channels per CTA

template <int CIN>
__global__ void __launch_bounds__(SC_ROWS) k_conv_small_cin(ConvArgs a) {
  extern __shared__ __align__(16) float sm_small[];
  float *w_s = sm_small;                                   // [V][CIN][SC_TILE]
  float *o_s = sm_small + a.V * CIN * SC_TILE;             // [4 warps][32][SC_TILE + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.y * SC_TILE;
  const int nc = min(SC_TILE, a.c_out - n0);
  for (int e = tid; e < a.V * CIN * SC_TILE; e += SC_ROWS) {
    const int c = e % SC_TILE, kc = e / SC_TILE;
    w_s[e] = c < nc ? __ldg(&a.weight[(long long)kc * a.c_out + n0 + c]) : 0.f;
  }
  __syncthreads();
  const int row = blockIdx.x * SC_ROWS + tid;
  float acc[SC_TILE];
#pragma unroll
  for (int c = 0; c < SC_TILE; ++c) acc[c] = 0.f;
  // Taps in batches of TB: all table entries of a batch are loaded first, then all neighbour inputs, then the FMAs -- TB and
  // then TB*CIN independent loads in flight per thread instead of a dependent pair per tap (the kernel was bound by that chain,
  // not by its 81*Cout FMAs per row).
  constexpr int TB = 9;
  for (int k0 = 0; k0 < a.V; k0 += TB) {
    int t[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const int k = k0 + j;
      const int trow = a.mirror ? a.V - 1 - k : k;
      t[j] = (k < a.V && row < a.n_rows) ? __ldg(&a.tbl[(long long)trow * a.tbl_stride + row]) : -1;
    }
    float x[TB][CIN];
#pragma unroll
    for (int j = 0; j < TB; ++j)
#pragma unroll
      for (int i = 0; i < CIN; ++i) x[j][i] = t[j] >= 0 ? __ldg(&a.in[(long long)t[j] * CIN + i]) : 0.f;
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      if (k0 + j >= a.V || !__any_sync(0xffffffffu, t[j] >= 0)) continue;
      const float4 *wk = reinterpret_cast<const float4 *>(w_s + (k0 + j) * CIN * SC_TILE);
#pragma unroll
      for (int i = 0; i < CIN; ++i) {
#pragma unroll
        for (int c4 = 0; c4 < SC_TILE / 4; ++c4) {
          const float4 w = wk[i * (SC_TILE / 4) + c4];
          acc[c4 * 4 + 0] = fmaf(x[j][i], w.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(x[j][i], w.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(x[j][i], w.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(x[j][i], w.w, acc[c4 * 4 + 3]);
        }
      }
    }
  }
  // transpose through shared memory so that a warp writes whole 256-byte rows
  float *mine = o_s + warp * 32 * (SC_TILE + 1);
#pragma unroll
  for (int c = 0; c < SC_TILE; ++c) mine[lane * (SC_TILE + 1) + c] = acc[c] + (a.bias && c < nc ? __ldg(&a.bias[n0 + c]) : 0.f);
  __syncwarp();
  const int row_w = blockIdx.x * SC_ROWS + warp * 32;
  for (int r = 0; r < 32; ++r) {
    if (row_w + r >= a.n_rows) break;
    float *dst = a.out + (long long)(row_w + r) * a.c_out + n0;
    for (int c = lane; c < nc; c += 32) dst[c] = mine[r * (SC_TILE + 1) + c];
  }
}

bool conv_small_supported(const ConvArgs &a) {
  return !a.scatter && a.c_in >= 1 && a.c_in <= 4 && a.V <= 32 && a.weight != nullptr;
}

void conv_small(const ConvArgs &a, cudaStream_t s) {
  if (a.n_rows == 0) return;
  dim3 grid((a.n_rows + SC_ROWS - 1) / SC_ROWS, (a.c_out + SC_TILE - 1) / SC_TILE);
  const size_t smem = sizeof(float) * ((size_t)a.V * a.c_in * SC_TILE + 4 * 32 * (SC_TILE + 1));
#define SCN_SMALL(CIN)                                                                                          \
  do {                                                                                                          \
    static SmemAttrCache smem_attr;                                                                             \
    smem_attr.ensure(k_conv_small_cin<CIN>, 100 * 1024);                                                        \
    k_conv_small_cin<CIN><<<grid, SC_ROWS, smem, s>>>(a);                                                       \
  } while (0)
  switch (a.c_in) {
    case 1: SCN_SMALL(1); break;
    case 2: SCN_SMALL(2); break;
    case 3: SCN_SMALL(3); break;
    default: SCN_SMALL(4); break;
  }
#undef SCN_SMALL
  SCN_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int SW_THREADS = 256;   // 64 channels x 4 tap groups
constexpr int SW_ROWS = 32;       // rows staged per step
constexpr int SW_TAPS = 8;        // taps per group (4 groups cover V <= 32)

template <int CIN>
__global__ void __launch_bounds__(SW_THREADS) k_wgrad_small_cin(WgradArgs a, int rows_per_cta) {
  __shared__ __align__(16) float x_s[SW_ROWS][4 * SW_TAPS * CIN + 4];   // [row][tap group][tap in group][ci] (+4: fewer bank conflicts)
  __shared__ float g_s[SW_ROWS][SC_TILE];
  const int tid = threadIdx.x;
  const int co = tid & 63, q = tid >> 6;
  const int n0 = blockIdx.y * SC_TILE;
  const bool live = n0 + co < a.c_b;
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  const long long r_end = r_begin + rows_per_cta < a.n_rows ? r_begin + rows_per_cta : a.n_rows;
  float acc[SW_TAPS * CIN];
#pragma unroll
  for (int i = 0; i < SW_TAPS * CIN; ++i) acc[i] = 0.f;
  // Software pipeline through registers: while the FMAs of step i run from shared memory, the neighbour inputs and gradients of
  // step i+1 and the table entries of step i+2 are in flight (the table -> input dependency is two steps deep), so the
  // latency of the staging loads no longer alternates with the arithmetic.
  constexpr int XS = SW_ROWS * 32 / SW_THREADS;          // (row, tap) slots per thread and step
  constexpr int GS = SW_ROWS * SC_TILE / SW_THREADS;     // gradient elements per thread and step
  int idx[XS];
  float xr[XS][CIN], gr[GS];
  auto load_idx = [&](long long r0) {
#pragma unroll
    for (int j = 0; j < XS; ++j) {
      const int e = tid + j * SW_THREADS;
      const int r = e & (SW_ROWS - 1), k = e / SW_ROWS;        // consecutive threads -> consecutive rows of one tap
      const long long row = r0 + r;
      idx[j] = (k < a.V && row < r_end) ? __ldg(&a.tbl[(long long)k * a.tbl_stride + row]) : -1;
    }
  };
  auto load_rows = [&](long long r0) {
#pragma unroll
    for (int j = 0; j < XS; ++j)
#pragma unroll
      for (int i = 0; i < CIN; ++i) xr[j][i] = idx[j] >= 0 ? __ldg(&a.a[(long long)idx[j] * CIN + i]) : 0.f;
#pragma unroll
    for (int j = 0; j < GS; ++j) {
      const int e = tid + j * SW_THREADS;
      const int r = e / SC_TILE, c = e % SC_TILE;
      const long long row = r0 + r;
      gr[j] = (row < r_end && n0 + c < a.c_b) ? __ldg(&a.b[row * a.c_b + n0 + c]) : 0.f;
    }
  };
  load_idx(r_begin);
  load_rows(r_begin);
  load_idx(r_begin + SW_ROWS);
  for (long long r0 = r_begin; r0 < r_end; r0 += SW_ROWS) {
#pragma unroll
    for (int j = 0; j < XS; ++j) {
      const int e = tid + j * SW_THREADS;
      float *dst = &x_s[e & (SW_ROWS - 1)][(e / SW_ROWS) * CIN];
#pragma unroll
      for (int i = 0; i < CIN; ++i) dst[i] = xr[j][i];
    }
#pragma unroll
    for (int j = 0; j < GS; ++j) {
      const int e = tid + j * SW_THREADS;
      g_s[e / SC_TILE][e % SC_TILE] = gr[j];
    }
    __syncthreads();
    load_rows(r0 + SW_ROWS);                   // rows >= r_end read as absent / zero
    load_idx(r0 + 2 * SW_ROWS);
#pragma unroll 4
    for (int r = 0; r < SW_ROWS; ++r) {
      const float gv = g_s[r][co];
      const float4 *xr4 = reinterpret_cast<const float4 *>(&x_s[r][q * SW_TAPS * CIN]);
#pragma unroll
      for (int i = 0; i < SW_TAPS * CIN / 4; ++i) {
        const float4 xv = xr4[i];
        acc[4 * i + 0] = fmaf(xv.x, gv, acc[4 * i + 0]);
        acc[4 * i + 1] = fmaf(xv.y, gv, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(xv.z, gv, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(xv.w, gv, acc[4 * i + 3]);
      }
    }
    __syncthreads();
  }
  if (!live) return;
#pragma unroll
  for (int kk = 0; kk < SW_TAPS; ++kk) {
    const int k = q * SW_TAPS + kk;
    if (k >= a.V) break;
#pragma unroll
    for (int i = 0; i < CIN; ++i)
      atomicAdd(&a.dw[blockIdx.x * a.part_stride + ((long long)k * CIN + i) * a.c_b + n0 + co], acc[kk * CIN + i]);
  }
}

bool wgrad_small_supported(const WgradArgs &a) {
  return a.table_on_a && a.c_a >= 1 && a.c_a <= 4 && a.V <= 32;
}

void wgrad_small(const WgradArgs &a_in, cudaStream_t s) {
  WgradArgs a = a_in;
  const size_t n_dw = (size_t)a.V * a.c_a * a.c_b;
  SCN_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * n_dw, s));
  if (a.n_rows == 0) return;
  const int tiles = (a.c_b + SC_TILE - 1) / SC_TILE;
  long long ctas = (long long)sm_count() * 4 / tiles;
  if (ctas < 1) ctas = 1;
  long long rows = (a.n_rows + ctas - 1) / ctas;
  rows = (rows + SW_ROWS - 1) / SW_ROWS * SW_ROWS;
  if (rows < 8 * SW_ROWS) rows = 8 * SW_ROWS;
  dim3 grid((unsigned)((a.n_rows + rows - 1) / rows), tiles);
  DevBuf<float> part;
  if (deterministic() && grid.x > 1) {      // see wgrad_simt
    part.alloc(n_dw * grid.x, s);
    SCN_CUDA(cudaMemsetAsync(part.p, 0, sizeof(float) * n_dw * grid.x, s));
    a.dw = part.p;
    a.part_stride = (long long)n_dw;
  }
  switch (a.c_a) {
    case 1: k_wgrad_small_cin<1><<<grid, SW_THREADS, 0, s>>>(a, (int)rows); break;
    case 2: k_wgrad_small_cin<2><<<grid, SW_THREADS, 0, s>>>(a, (int)rows); break;
    case 3: k_wgrad_small_cin<3><<<grid, SW_THREADS, 0, s>>>(a, (int)rows); break;
    default: k_wgrad_small_cin<4><<<grid, SW_THREADS, 0, s>>>(a, (int)rows); break;
  }
  SCN_LAUNCH_CHECK();
  if (part.p) sum_partials(part.p, (int)grid.x, (long long)n_dw, a_in.dw, s);
  part.release(s);
}

}  // namespace scn

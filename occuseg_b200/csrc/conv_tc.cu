// tcgen05 (5th-generation tensor core) table convolutions for sm_100a: the SCN_TF32 path.
//
//   out[o, :] = sum_k  in[tbl[k][o], :] * W[k]          (gather - GEMM, output stationary)
//
// One CTA owns a tile of 128 output rows x TN output channels.  The fp32 accumulator lives in TMEM
// (128 lanes x TN columns) for the whole walk over the taps, so the output is written exactly once and
// no atomics are needed.  For every (tap, 32-channel K chunk) the four producer warps gather the 128
// neighbour rows (128 bytes each) with 16-byte cp.async straight into the 128B-swizzled K-major layout
// that tcgen05.mma reads, absent neighbours are zero-filled by cp.async's src-size operand, and the
// matching [TN x 32] slice of the pre-transposed weights goes into the same stage.  A single elected
// thread issues tcgen05.mma.kind::tf32 (M=128, N=TN, K=8, fp32 accumulate) and releases the stage with
// tcgen05.commit; producers run up to STAGES-1 chunks ahead.  Taps with no neighbour inside the tile
// are skipped by both sides.  fp32 features are consumed as TF32 (10-bit mantissa) -> rel 2e-2 budget.
//
// Replaces the reference's scalar shared-memory FMA kernels (CUDA/Convolution.cu:1059-1152 forward,
// :447-534 dgrad) for channel counts that are multiples of 32.
#include "common.cuh"
#include <cstdlib>

namespace scn {

namespace tc {

constexpr int TM = 128;               // rows per tile == UMMA M
constexpr int KCH = 32;               // tf32 elements per 128-byte swizzle row
constexpr int A_STAGE = TM * 128;     // bytes
constexpr int NPROD = 128;            // producer threads (warps 0-3), also the epilogue threads
constexpr int NTHREADS = 160;         // + warp 4: TMEM allocator and MMA issuer

struct Params {
  const float *in;
  const float *w_nk;      // [V][c_out][c_in]
  const float *bias;
  float *out;
  const int *tbl;
  int tbl_stride, n_rows, V, c_in, c_out;
  int mirror;
  int TN, stages, lag, b_stage, tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1, layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major), 16 B
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NTHREADS) k_conv_tc(Params p) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [A stages][B stages][s_idx V*128 ints][barriers]
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + p.stages * A_STAGE;
  int *s_idx = reinterpret_cast<int *>(smem + p.stages * (A_STAGE + p.b_stage));
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_idx + p.V * TM);
  const uint32_t full_bar = smem_u32(bars);                 // [stages]
  const uint32_t empty_bar = full_bar + 8 * p.stages;       // [stages]
  const uint32_t accum_bar = empty_bar + 8 * p.stages;      // accumulator complete
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);
  uint32_t *s_mask = s_tmem + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * p.TN;
  const int KC = p.c_in / KCH;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, NPROD);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    *s_mask = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  } else {
    // neighbour rows of this tile for every tap, and the set of taps that touch the tile at all
    uint32_t mine = 0;
    const int r = row0 + tid;
    for (int k = 0; k < p.V; ++k) {
      int t = (r < p.n_rows) ? __ldg(&p.tbl[(long long)k * p.tbl_stride + r]) : -1;
      s_idx[k * TM + tid] = t;
      mine |= (t >= 0 ? 1u : 0u) << k;
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(s_mask, mine);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t tapmask = *s_mask;     // bit k: table row k has at least one neighbour in this tile
  const int n_taps = __popc(tapmask);
  const int n_items = n_taps * KC;

  if (warp < 4) {
    // =========================== producers ===========================
    uint32_t remaining = tapmask;
    int trow = -1, kc = KC;            // current table row and K chunk
    for (int it = 0; it < n_items + p.lag; ++it) {
      if (it < n_items) {
        if (kc == KC) {                // next active tap
          trow = __ffs(remaining) - 1;
          remaining &= remaining - 1;
          kc = 0;
        }
        const int s = it % p.stages;
        mbar_wait(empty_bar + 8 * s, ((it / p.stages) & 1) ^ 1);
        // table row `trow` pairs with weight tap (mirror ? V-1-trow : trow)
        const int wtap = p.mirror ? p.V - 1 - trow : trow;
        const uint32_t a_st = a_base + s * A_STAGE;
        const int *idx = s_idx + trow * TM;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = warp * 32 + i * 4 + (lane >> 3);
          const int c = lane & 7;
          const int src = idx[r];
          const float *g = p.in + (long long)(src < 0 ? 0 : src) * p.c_in + kc * KCH + c * 4;
          const uint32_t dst = a_st + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
          cp_async16(dst, g, src < 0 ? 0u : 16u);
        }
        const uint32_t b_st = b_base + s * p.b_stage;
        const float *wsrc = p.w_nk + ((long long)wtap * p.c_out + n0) * p.c_in + kc * KCH;
        for (int q = tid; q < p.TN * 8; q += NPROD) {
          const int n = q >> 3, c = q & 7;
          const uint32_t dst = b_st + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4);
          cp_async16(dst, wsrc + (long long)n * p.c_in + c * 4, 16u);
        }
        ++kc;
      }
      cp_async_commit();
      if (it >= p.lag) {
        // the group committed `lag` iterations ago has landed: publish it to the tensor core (async proxy)
        if (p.lag == 2) cp_async_wait<2>();
        else cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(full_bar + 8 * ((it - p.lag) % p.stages));
      }
    }
    // =========================== epilogue ===========================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int r = row0 + tid;
    float *orow = p.out + (long long)r * p.c_out + n0;
    for (int c0 = 0; c0 < p.TN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);   // warp w owns TMEM lanes 32w..32w+31
      if (r < p.n_rows) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.bias) {
            o.x += __ldg(&p.bias[n0 + c0 + j]);
            o.y += __ldg(&p.bias[n0 + c0 + j + 1]);
            o.z += __ldg(&p.bias[n0 + c0 + j + 2]);
            o.w += __ldg(&p.bias[n0 + c0 + j + 3]);
          }
          *reinterpret_cast<float4 *>(orow + c0 + j) = o;
        }
      }
    }
  } else if (lane == 0) {
    // =========================== MMA issuer (one thread) ===========================
    const uint32_t idesc = make_idesc(p.TN);
    for (int it = 0; it < n_items; ++it) {
      const int s = it % p.stages;
      mbar_wait(full_bar + 8 * s, (it / p.stages) & 1);
      tc_fence_after();
      const uint64_t ad = make_desc(a_base + s * A_STAGE);
      const uint64_t bd = make_desc(b_base + s * p.b_stage);
#pragma unroll
      for (int k = 0; k < KCH / 8; ++k)   // K = 8 per tf32 MMA = 32 bytes inside the swizzled row
        mma_tf32(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it | k) ? 1u : 0u);
      mma_commit(empty_bar + 8 * s);      // stage reusable once these MMAs have read it
    }
    mma_commit(accum_bar);                // accumulator complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols));
  }
}

static int pick_tn(int c_out) {
  // largest multiple of 32 that divides c_out and fits one UMMA (N <= 256)
  for (int tn = 256; tn >= 32; tn -= 32)
    if (c_out % tn == 0) return tn;
  return 0;
}

}  // namespace tc

bool conv_tc_supported(const ConvArgs &a) {
  return !a.scatter && a.c_in >= 32 && a.c_in % 32 == 0 && a.c_out >= 32 && a.c_out % 32 == 0 && a.V <= 32 &&
         tc::pick_tn(a.c_out) > 0 && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}

void conv_tc(const ConvArgs &a, cudaStream_t s) {
  using namespace tc;
  SCN_CHECK(a.weight_nk != nullptr, "conv_tc needs the [V][Cout][Cin] weight layout");
  if (a.n_rows == 0) return;
  Params p;
  p.in = a.in; p.w_nk = a.weight_nk; p.bias = a.bias; p.out = a.out; p.tbl = a.tbl;
  p.tbl_stride = a.tbl_stride; p.n_rows = a.n_rows; p.V = a.V; p.c_in = a.c_in; p.c_out = a.c_out;
  p.mirror = a.mirror ? 1 : 0;
  p.TN = pick_tn(a.c_out);
  p.b_stage = p.TN * 128;
  int stage = A_STAGE + p.b_stage;
  p.stages = 100 * 1024 / stage;
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 2) p.stages = 2;
  p.lag = p.stages >= 3 ? 2 : 1;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.TN) p.tmem_cols <<= 1;
  size_t smem = 1024 + (size_t)p.stages * stage + (size_t)a.V * TM * sizeof(int) + 8 * (2 * p.stages + 1) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    SCN_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((a.n_rows + TM - 1) / TM, a.c_out / p.TN);
  k_conv_tc<<<grid, NTHREADS, smem, s>>>(p);
  SCN_LAUNCH_CHECK();
}

// =====================================================================================================
// weight gradient on tensor cores
//
//   D_k[Cg x Cs] = sum_r  G[tbl[k][r], :]^T  S[r, :]            dW[k] = D_k  (or D_k^T for Deconvolution)
//
// G is the gathered operand (input rows for Sub/Convolution, fine d_out rows for Deconvolution), S the
// stationary one (rows r themselves).  The GEMM's K dimension is the voxel row, so both operands sit in
// shared memory MN-major: a gathered 128-byte row piece (32 channels) is one line of the SWIZZLE_128B_BASE32B
// atom (the only MN-major layout tcgen05 accepts for tf32), 4 consecutive rows form one 512-byte K group.  M = 128 stacks four 32-channel "slots"; slots
// enumerate (tap, channel group) pairs, so for Cg = 64 one MMA covers two taps.  Each CTA keeps up to
// 512/N accumulators in TMEM, walks its share of the rows in 32-row chunks (chunks in which none of an
// accumulator's taps has a rule are skipped using a per-chunk tap mask built with the rulebook) and
// finally adds its partial sums into dW with fp32 atomics.
// Replaces dConvolution_KMxKN_backward_dW_RuleBookBased (CUDA/Convolution.cu:695-753, 27 launches with
// 27 blocking rule uploads, :789-807).
// =====================================================================================================
namespace wg {
using namespace tc;

constexpr int KR = 32;            // voxel rows (GEMM K) per pipeline item
constexpr int SUB = KR * 128;     // bytes of one [KR rows x 32 channels] sub-tile

struct Params {
  const float *G, *S;
  float *dw;
  const int *tbl;
  const uint32_t *cmask;
  int tbl_stride, n_rows, V, Cg, Cs, transpose_out;
  int N, acc_per_cta, n_acc_total, rows_per_cta;
  int stages, lag, stage_bytes, tmem_cols;
  int dbg;   // SCN_WG_DBG bit0: skip atomics, bit1: skip MMAs, bit2: skip gathered loads, bit3: skip stationary loads
};

// MN-major tf32 operands have exactly one legal shared-memory layout: SWIZZLE_128B_BASE32B (layout type 1).
// Its atom is 4 K-rows x 128 bytes (32 channels); inside a row the 32-byte unit index is XORed with the row
// index (address bits [5,7) ^= bits [7,9)).  32-channel atoms are SUB bytes apart (LBO), 4-row K groups
// 512 bytes apart (SBO); one K=8 MMA therefore reads two K groups = 1024 bytes of every atom.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(SUB >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// byte offset of 16-byte piece cc (0..7) of K-row `row` inside a [KR x 128 B] sub-tile
__device__ __forceinline__ uint32_t mn_offset(int row, int cc) {
  return (uint32_t)(row * 128 + ((((cc >> 1) ^ (row & 3)) << 5) | ((cc & 1) << 4)));
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}

__global__ void __launch_bounds__(NTHREADS) k_wgrad_tc(Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * p.stage_bytes);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t accum_bar = empty_bar + 8 * p.stages;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);
  int *s_idx = reinterpret_cast<int *>(s_tmem + 2);  // [2][taps of this CTA][KR] gathered-row indices, double buffered

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int spt = p.Cg >> 5;                         // 32-channel slots per tap
  const int n_slots = p.V * spt;
  const int acc0 = blockIdx.y * p.acc_per_cta;       // first accumulator of this CTA
  const int n_acc = min(p.acc_per_cta, p.n_acc_total - acc0);
  const int n0 = blockIdx.z * p.N;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(r_begin + p.rows_per_cta, p.n_rows);
  const int n_chunks = (r_end - r_begin + KR - 1) / KR;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, NPROD);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // taps touched by accumulator j (bit per tap); identical in every thread
  auto acc_tapmask = [&](int ja) -> uint32_t {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int slot = 4 * ja + q;
      if (slot < n_slots) m |= 1u << (slot / spt);
    }
    return m;
  };

  if (warp < 4) {
    // =========================== producers ===========================
    int it = 0;                 // items issued
    int pending = 0;            // items issued whose arrival is still owed (<= lag)
    auto publish_oldest = [&](int owed_index) { mbar_arrive(full_bar + 8 * (owed_index % p.stages)); };
    const int tap_lo = (4 * acc0) / spt;
    const int tap_hi = min(p.V - 1, (4 * (acc0 + n_acc) - 1) / spt);
    const int n_taps_cta = tap_hi - tap_lo + 1;
    for (int c = 0; c < n_chunks; ++c) {
      const int r0 = r_begin + c * KR;
      const uint32_t cm = __ldg(&p.cmask[r0 >> 5]);
      // stage this chunk's table entries once (they are reused by every slot of a tap and by 8 lanes each)
      int *idx = s_idx + (c & 1) * (p.V * KR);
      for (int e = tid; e < n_taps_cta * KR; e += NPROD) {
        const int t = e / KR, row = e - t * KR;
        idx[e] = (r0 + row < r_end) ? __ldg(&p.tbl[(long long)(tap_lo + t) * p.tbl_stride + r0 + row]) : -1;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
      for (int jl = 0; jl < n_acc; ++jl) {
        if (!(cm & acc_tapmask(acc0 + jl))) continue;
        const int s = it % p.stages;
        mbar_wait(empty_bar + 8 * s, ((it / p.stages) & 1) ^ 1);
        const uint32_t st = base + s * p.stage_bytes;
        // gathered operand: 4 slots x 32 rows x 8 sixteen-byte pieces
        if (!(p.dbg & 4))
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int q = i >> 1;
          const int row = (tid >> 3) + 16 * (i & 1);
          const int cc = tid & 7;
          const int slot = 4 * (acc0 + jl) + q;
          int src = -1;
          int ch = 0;
          if (slot < n_slots) {
            const int tap = slot / spt;
            ch = (slot - tap * spt) << 5;
            src = idx[(tap - tap_lo) * KR + row];
          }
          const float *g = p.G + (long long)(src < 0 ? 0 : src) * p.Cg + ch + cc * 4;
          const uint32_t dst = st + q * SUB + mn_offset(row, cc);
          cp_async16(dst, g, src < 0 ? 0u : 16u);
        }
        // stationary operand: N/32 atoms x 32 rows x 8 pieces
        if (!(p.dbg & 8))
        for (int q = tid; q < p.N * 8; q += NPROD) {
          const int atom = q >> 8, rem = q & 255, row = rem >> 3, cc = rem & 7;
          const bool live = r0 + row < r_end;
          const float *g = p.S + (long long)(live ? r0 + row : 0) * p.Cs + n0 + atom * 32 + cc * 4;
          const uint32_t dst = st + (4 + atom) * SUB + mn_offset(row, cc);
          cp_async16(dst, g, live ? 16u : 0u);
        }
        cp_async_commit();
        ++it;
        ++pending;
        if (pending > p.lag) {
          if (p.lag == 3) cp_async_wait<3>();
          else if (p.lag == 2) cp_async_wait<2>();
          else cp_async_wait<1>();
          fence_proxy_async();
          publish_oldest(it - pending);
          --pending;
        }
      }
    }
    // drain
    cp_async_wait<0>();
    fence_proxy_async();
    while (pending > 0) {
      publish_oldest(it - pending);
      --pending;
    }
    // =========================== epilogue ===========================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // which accumulators received at least one MMA (same scan the MMA thread did)
    uint32_t started = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t cm = __ldg(&p.cmask[(r_begin + c * KR) >> 5]);
      for (int jl = 0; jl < n_acc; ++jl)
        if (cm & acc_tapmask(acc0 + jl)) started |= 1u << jl;
    }
    for (int jl = 0; jl < n_acc; ++jl) {
      if (!(started & (1u << jl))) continue;
      const int slot = 4 * (acc0 + jl) + warp;         // TMEM lane = warp*32 + lane -> slot `warp` of the MMA's M
      if (slot >= n_slots) continue;                   // warp-uniform
      const int tap = slot / spt;
      const int cg = ((slot - tap * spt) << 5) + lane;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + jl * p.N + c0, v);
        if (p.dbg & 1) continue;
        if (!p.transpose_out) {
          float *dst = p.dw + ((long long)tap * p.Cg + cg) * p.Cs + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
        } else {
          float *dst = p.dw + ((long long)tap * p.Cs + n0 + c0) * p.Cg + cg;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + (long long)j * p.Cg, v[j]);
        }
      }
    }
  } else if (lane == 0) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_mn(p.N);
    uint32_t started = 0;
    int it = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t cm = __ldg(&p.cmask[(r_begin + c * KR) >> 5]);
      for (int jl = 0; jl < n_acc; ++jl) {
        if (!(cm & acc_tapmask(acc0 + jl))) continue;
        const int s = it % p.stages;
        mbar_wait(full_bar + 8 * s, (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t st = base + s * p.stage_bytes;
        const uint64_t ad = make_desc_mn(st);
        const uint64_t bd = make_desc_mn(st + 4 * SUB);
        if (!(p.dbg & 2))
#pragma unroll
        for (int k = 0; k < KR / 8; ++k)     // K = 8 rows per tf32 MMA = one 1024-byte K group
          mma_tf32(tmem + jl * p.N, ad + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), idesc,
                   ((started >> jl) & 1u) | (k ? 1u : 0u));
        started |= 1u << jl;
        mma_commit(empty_bar + 8 * s);
        ++it;
      }
    }
    mma_commit(accum_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols));
  }
}

}  // namespace wg

bool wgrad_tc_supported(const WgradArgs &a) {
  const int cg = a.table_on_a ? a.c_a : a.c_b, cs = a.table_on_a ? a.c_b : a.c_a;
  return a.chunk_mask != nullptr && cg >= 32 && cg % 32 == 0 && cs >= 32 && cs % 32 == 0 && a.V <= 32 &&
         tc::pick_tn(cs) > 0 && ((uintptr_t)a.a % 16 == 0) && ((uintptr_t)a.b % 16 == 0);
}

void wgrad_tc(const WgradArgs &a, cudaStream_t s) {
  using wg::KR;
  using wg::SUB;
  SCN_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.V * a.c_a * a.c_b, s));
  if (a.n_rows == 0) return;
  wg::Params p;
  p.G = a.table_on_a ? a.a : a.b;
  p.S = a.table_on_a ? a.b : a.a;
  p.Cg = a.table_on_a ? a.c_a : a.c_b;
  p.Cs = a.table_on_a ? a.c_b : a.c_a;
  p.transpose_out = a.table_on_a ? 0 : 1;
  p.dw = a.dw; p.tbl = a.tbl; p.cmask = a.chunk_mask; p.tbl_stride = a.tbl_stride; p.n_rows = a.n_rows; p.V = a.V;
  p.N = tc::pick_tn(p.Cs);
  const int n_slots = a.V * (p.Cg / 32);
  p.n_acc_total = (n_slots + 3) / 4;
  int max_acc = 512 / p.N;
  int groups = (p.n_acc_total + max_acc - 1) / max_acc;
  p.acc_per_cta = (p.n_acc_total + groups - 1) / groups;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.acc_per_cta * p.N) p.tmem_cols <<= 1;
  const int n_tiles_n = p.Cs / p.N;
  int row_splits = (2 * sm_count()) / (groups * n_tiles_n);
  if (row_splits < 1) row_splits = 1;
  int rows = (a.n_rows + row_splits - 1) / row_splits;
  if (rows < 512) rows = 512;
  p.rows_per_cta = (rows + KR - 1) / KR * KR;
  row_splits = (a.n_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  p.stage_bytes = (4 + p.N / 32) * SUB;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.lag = p.stages >= 5 ? 3 : (p.stages >= 3 ? 2 : 1);
  p.stages = (int)((200 * 1024 - 2 * a.V * KR * sizeof(int)) / p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  p.lag = p.stages >= 5 ? 3 : (p.stages >= 3 ? 2 : 1);
  size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 8 * (2 * p.stages + 1) + 16 + 2 * a.V * KR * sizeof(int);
  static size_t configured = 0;
  if (smem > configured) {
    SCN_CUDA(cudaFuncSetAttribute(wg::k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  {
    const char *e = getenv("SCN_WG_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  dim3 grid(row_splits, groups, n_tiles_n);
  wg::k_wgrad_tc<<<grid, tc::NTHREADS, smem, s>>>(p);
  SCN_LAUNCH_CHECK();
}

}  // namespace scn

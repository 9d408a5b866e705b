// TMA-fed tcgen05 table convolutions for sm_100a (SCN_TF32 path, default).
//
// Same math and tiling as conv_tc.cu, but the operands are moved by the Tensor Memory Accelerator instead
// of by per-thread cp.async:
//   * gathered rows:  cp.async.bulk.tensor.2d ... tile::gather4 -- ONE instruction fetches four arbitrary
//     rows (128 bytes each) of the feature matrix into four consecutive swizzled shared-memory rows; an
//     absent neighbour is requested as the out-of-bounds row index `rows`, which TMA zero-fills;
//   * weights / stationary rows: ordinary 2-D tile loads.
// A producer warp issues the copies (each lane one gather4), completion is counted in bytes on an
// mbarrier (expect_tx), the MMA thread consumes the stage and frees it with tcgen05.commit.  No LSU
// traffic, no per-thread fences, no 128-way barrier arrivals: the hand-off that dominated the cp.async
// version (profiles/r01_notes.md) is one arrive + one wait per stage.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace scn {
namespace tma {

// ------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (EncodeFn)p;
  }();
  return fn;
}

// row-major fp32 matrix [rows, cols]; box = box_cols x box_rows elements
static CUtensorMap make_map(const float *base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                            CUtensorMapSwizzle sw) {
  EncodeFn fn = encode_fn();
  SCN_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return m;
}

// ------------------------------------------------------------------------------------------ device helpers
constexpr int TM = 128;
constexpr int KCH = 32;
constexpr int A_STAGE = TM * 128;
constexpr int NTHREADS = 192;     // warp 0 producer, warp 1 MMA + TMEM owner, warps 2-5 epilogue (TMEM quarter = warp & 3)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"((uint64_t)map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
//  K-major  SWIZZLE_128B        : 128-byte rows, 8-row groups 1024 B apart (SBO)
//  MN-major SWIZZLE_128B_BASE32B: 32-channel atoms `lbo` bytes apart, 4-row K groups 512 B apart -- the only
//                                 MN-major layout tcgen05 accepts for tf32
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn32(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// instruction descriptor (InstrDescriptor): kind::tf32, fp32 accumulate, M=128, N=n; majors: 0 = K, 1 = MN
__device__ __forceinline__ uint32_t idesc_tf32(int n, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
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
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}

// =====================================================================================================
// forward / dgrad:  out[o,:] = sum_k in[tbl[k][o],:] * W[k]
// =====================================================================================================
struct ConvParams {
  const float *bias;
  float *out;
  const int *tbl;
  int tbl_stride, n_rows, in_rows, V, c_in, c_out, mirror;
  int TN, stages, b_stage, tmem_cols;
};

__global__ void __launch_bounds__(NTHREADS) k_conv_tma(const __grid_constant__ CUtensorMap map_x,
                                                       const __grid_constant__ CUtensorMap map_w, ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + p.stages * A_STAGE;
  int *s_idx = reinterpret_cast<int *>(smem + p.stages * (A_STAGE + p.b_stage));
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_idx + p.V * TM);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t accum_bar = empty_bar + 8 * p.stages;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);
  uint32_t *s_mask = s_tmem + 1;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler
  const int row0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * p.TN;
  const int KC = p.c_in / KCH;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    *s_mask = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
  } else if (warp >= 2) {
    // neighbour rows of this tile for every tap (absent -> the out-of-bounds row TMA zero-fills) + tap mask
    const int e = tid - 64;
    const int r = row0 + e;
    uint32_t mine = 0;
    for (int k = 0; k < p.V; ++k) {
      int t = (r < p.n_rows) ? __ldg(&p.tbl[(long long)k * p.tbl_stride + r]) : -1;
      s_idx[k * TM + e] = t < 0 ? p.in_rows : t;
      mine |= (t >= 0 ? 1u : 0u) << k;
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(s_mask, mine);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t tapmask = *s_mask;
  const int n_items = __popc(tapmask) * KC;

  if (warp == 0) {
    // =========================== weight producer + transaction accounting ===========================
    if (elect_one()) {
      uint32_t remaining = tapmask;
      int it = 0;
      const uint32_t stage_bytes = (uint32_t)(A_STAGE + p.b_stage);
      while (remaining) {
        const int trow = __ffs(remaining) - 1;
        remaining &= remaining - 1;
        const int wtap = p.mirror ? p.V - 1 - trow : trow;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % p.stages;
          mbar_wait(empty_bar + 8 * s, ((it / p.stages) & 1) ^ 1);
          mbar_expect_tx(full_bar + 8 * s, stage_bytes);
          tma_tile_2d(b_base + s * p.b_stage, &map_w, kc * KCH, wtap * p.c_out + n0, full_bar + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(p.TN, 0, 0);
      for (int it = 0; it < n_items; ++it) {
        const int s = it % p.stages;
        mbar_wait(full_bar + 8 * s, (it / p.stages) & 1);
        tc_fence_after();
        const uint64_t ad = desc_k128(a_base + s * A_STAGE);
        const uint64_t bd = desc_k128(b_base + s * p.b_stage);
#pragma unroll
        for (int k = 0; k < KCH / 8; ++k)
          mma_tf32(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it | k) ? 1u : 0u);
        mma_commit(empty_bar + 8 * s);
      }
      mma_commit(accum_bar);
    }
  } else {
    // =========================== gather producers (warps 2-5), then epilogue ===========================
    // warp w issues the 8 gather4 copies of rows 32*(w-2) .. 32*(w-2)+31 of every stage from ONE elected
    // thread with warp-uniform operands (indices are read from shared memory at uniform addresses), so each
    // copy is a single UTMALDG instead of a per-lane vote loop.
    const int pw = warp - 2;
    if (elect_one()) {
      uint32_t remaining = tapmask;
      int it = 0;
      while (remaining) {
        const int trow = __ffs(remaining) - 1;
        remaining &= remaining - 1;
        const int4 *rows4 = reinterpret_cast<const int4 *>(s_idx + trow * TM + pw * 32);
        int4 r[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) r[g] = rows4[g];
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % p.stages;
          mbar_wait(empty_bar + 8 * s, ((it / p.stages) & 1) ^ 1);
          const uint32_t dst = a_base + s * A_STAGE + pw * 4096;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            tma_gather4(dst + g * 512, &map_x, kc * KCH, r[g].x, r[g].y, r[g].z, r[g].w, full_bar + 8 * s);
        }
      }
    }
    __syncwarp();
    const int quarter = warp & 3;                     // TMEM lanes 32*quarter .. +31 belong to this warp
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int r = row0 + quarter * 32 + lane;
    float *orow = p.out + (long long)r * p.c_out + n0;
    for (int c0 = 0; c0 < p.TN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

static int pick_tn(int c_out) {
  for (int tn = 256; tn >= 32; tn -= 32)
    if (c_out % tn == 0) return tn;
  return 0;
}

// =====================================================================================================
// weight gradient:  D_k[Cg x Cs] = sum_r G[tbl[k][r],:]^T S[r,:]     (see conv_tc.cu for the slot scheme)
// =====================================================================================================
constexpr int KR = 32;            // voxel rows (GEMM K) per pipeline item
constexpr int SUB = KR * 128;     // bytes of one [KR rows x 32 channels] sub-tile
constexpr int G_STAGE = 4 * SUB;  // four 32-channel slots = UMMA M of 128

struct WgParams {
  float *dw;
  const int *tbl;
  const uint32_t *cmask;
  int tbl_stride, n_rows, g_rows, V, Cg, Cs, transpose_out;
  int N, acc_per_cta, n_acc_total, rows_per_cta;
  int g_stages, s_stages, s_stage, tmem_cols;
  int dbg;   // SCN_WG_DBG: bit1 skip MMAs, bit2 skip gathers, bit3 skip stationary loads (timing experiments only)
};

__global__ void __launch_bounds__(NTHREADS) k_wgrad_tma(const __grid_constant__ CUtensorMap map_g,
                                                        const __grid_constant__ CUtensorMap map_s, WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const uint32_t g_base = base;
  const uint32_t s_base = base + p.g_stages * G_STAGE;
  int *s_idx = reinterpret_cast<int *>(smem + p.g_stages * G_STAGE + p.s_stages * p.s_stage);   // [4 warps][2][V][KR]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_idx + 4 * 2 * p.V * KR);
  const uint32_t g_full = smem_u32(bars);
  const uint32_t g_empty = g_full + 8 * p.g_stages;
  const uint32_t s_full = g_empty + 8 * p.g_stages;
  const uint32_t s_empty = s_full + 8 * p.s_stages;
  const uint32_t accum_bar = s_empty + 8 * p.s_stages;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.g_stages + 2 * p.s_stages + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int spt = p.Cg >> 5;
  const int n_slots = p.V * spt;
  const int acc0 = blockIdx.y * p.acc_per_cta;
  const int n_acc = min(p.acc_per_cta, p.n_acc_total - acc0);
  const int n0 = blockIdx.z * p.N;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(r_begin + p.rows_per_cta, p.n_rows);
  const int n_chunks = (r_end - r_begin + KR - 1) / KR;
#define CM(i) ((p.dbg & 64) ? 0xFFFFFFFFu : __ldg(&p.cmask[i]))

  if (tid == 0) {
    for (int s = 0; s < p.g_stages; ++s) {
      mbar_init(g_full + 8 * s, 1);
      mbar_init(g_empty + 8 * s, 1);
    }
    for (int s = 0; s < p.s_stages; ++s) {
      mbar_init(s_full + 8 * s, 1);
      mbar_init(s_empty + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // taps touched by accumulator ja (identical in every thread)
  auto acc_tapmask = [&](int ja) -> uint32_t {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int slot = 4 * ja + q;
      if (slot < n_slots) m |= 1u << (slot / spt);
    }
    return m;
  };
  uint32_t cta_taps = 0;
  for (int jl = 0; jl < n_acc; ++jl) cta_taps |= acc_tapmask(acc0 + jl);

  if (warp == 0) {
    // =========================== control producer: stationary rows + transaction accounting =========
    if (elect_one()) {
      int it = 0, sc = 0;
      uint32_t cm_next = CM(r_begin >> 5);
      for (int c = 0; c < n_chunks; ++c) {
        const int r0 = r_begin + c * KR;
        const uint32_t cm = cm_next;
        if (c + 1 < n_chunks) cm_next = CM((r0 + KR) >> 5);
        if (!(cm & cta_taps)) continue;
        const int cs = sc % p.s_stages;
        mbar_wait(s_empty + 8 * cs, ((sc / p.s_stages) & 1) ^ 1);
        mbar_expect_tx(s_full + 8 * cs, (p.dbg & 8) ? 0u : (uint32_t)p.s_stage);
        if (!(p.dbg & 8))
        for (int a = 0; a < p.N / 32; ++a)
          tma_tile_2d(s_base + cs * p.s_stage + a * SUB, &map_s, n0 + a * 32, r0, s_full + 8 * cs);
        ++sc;
        for (int jl = 0; jl < n_acc; ++jl) {
          if (!(cm & acc_tapmask(acc0 + jl))) continue;
          const int gs = it % p.g_stages;
          mbar_wait(g_empty + 8 * gs, ((it / p.g_stages) & 1) ^ 1);
          mbar_expect_tx(g_full + 8 * gs, (p.dbg & 4) ? 0u : (uint32_t)G_STAGE);
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(p.N, 1, 1);
      uint32_t started = 0;
      int it = 0, sc = 0;
      uint32_t cm_next = CM(r_begin >> 5);
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t cm = cm_next;
        if (c + 1 < n_chunks) cm_next = CM((r_begin + (c + 1) * KR) >> 5);
        if (!(cm & cta_taps)) continue;
        const int cs = sc % p.s_stages;
        mbar_wait(s_full + 8 * cs, (sc / p.s_stages) & 1);
        const uint64_t bd = desc_mn32(s_base + cs * p.s_stage, SUB);
        for (int jl = 0; jl < n_acc; ++jl) {
          if (!(cm & acc_tapmask(acc0 + jl))) continue;
          const int gs = it % p.g_stages;
          mbar_wait(g_full + 8 * gs, (it / p.g_stages) & 1);
          tc_fence_after();
          const uint64_t ad = desc_mn32(g_base + gs * G_STAGE, SUB);
          if (!(p.dbg & 2))
#pragma unroll
          for (int k = 0; k < KR / 8; ++k)    // K = 8 rows per MMA = two 512-byte K groups of every atom
            mma_tf32(tmem + jl * p.N, ad + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), idesc,
                     ((started >> jl) & 1u) | (k ? 1u : 0u));
          started |= 1u << jl;
          mma_commit(g_empty + 8 * gs);
          ++it;
        }
        mma_commit(s_empty + 8 * cs);
        ++sc;
      }
      mma_commit(accum_bar);
    }
  } else {
    // =========================== gather producers (warps 2-5), then epilogue ===========================
    // warp pw fills slot pw (32 channels of one tap) of every stage: 8 gather4 copies issued by one elected
    // thread with uniform operands.  Each warp keeps its own double-buffered copy of the chunk's table
    // entries, staged one chunk ahead with 4-byte cp.async, so no cross-warp synchronisation is needed.
    const int pw = warp - 2;
    const int tap_lo = (4 * acc0) / spt;
    const int tap_hi = min(p.V - 1, (4 * (acc0 + n_acc) - 1) / spt);
    const int n_taps_cta = tap_hi - tap_lo + 1;
    int *my_idx = s_idx + pw * (2 * p.V * KR);
    auto stage_idx = [&](int c) {
      int *dst = my_idx + (c & 1) * (p.V * KR);
      const int r0 = r_begin + c * KR;
      if (!(p.dbg & 16))
      for (int t = 0; t < n_taps_cta; ++t) {
        const int *src = p.tbl + (long long)(tap_lo + t) * p.tbl_stride + r0 + lane;   // table is padded to 128 rows
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + t * KR + lane)), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_idx(0);
    const bool leader = elect_one();
    uint32_t cm_next = CM(r_begin >> 5);
    int it = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t cm = cm_next;
      if (c + 1 < n_chunks) {
        cm_next = CM((r_begin + (c + 1) * KR) >> 5);
        stage_idx(c + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      const int *idx = my_idx + (c & 1) * (p.V * KR);
      for (int jl = 0; jl < n_acc; ++jl) {
        if (!(cm & acc_tapmask(acc0 + jl))) continue;
        const int gs = it % p.g_stages;
        ++it;
        if (!leader || (p.dbg & 4)) continue;
        mbar_wait(g_empty + 8 * gs, (((it - 1) / p.g_stages) & 1) ^ 1);
        const int slot = 4 * (acc0 + jl) + pw;
        const uint32_t dst = g_base + gs * G_STAGE + pw * SUB;
        if (slot < n_slots) {
          const int tap = slot / spt;
          const int ch = (slot - tap * spt) << 5;
          const int4 *rows4 = reinterpret_cast<const int4 *>(idx + (tap - tap_lo) * KR);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            int4 r = rows4[g];
            r.x = r.x < 0 ? p.g_rows : r.x;
            r.y = r.y < 0 ? p.g_rows : r.y;
            r.z = r.z < 0 ? p.g_rows : r.z;
            r.w = r.w < 0 ? p.g_rows : r.w;
            tma_gather4(dst + g * 512, &map_g, ch, r.x, r.y, r.z, r.w, g_full + 8 * gs);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g)   // slot past the last tap: all rows out of bounds -> zeros
            tma_gather4(dst + g * 512, &map_g, 0, p.g_rows, p.g_rows, p.g_rows, p.g_rows, g_full + 8 * gs);
        }
      }
      // the other lanes must not start staging chunk c+2 into this buffer while the leader still reads it
      __syncwarp();
    }
    // ---- epilogue: TMEM -> fp32 atomics into dW
    const int quarter = warp & 3;
    uint32_t started = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t cm = CM((r_begin + c * KR) >> 5);
      for (int jl = 0; jl < n_acc; ++jl)
        if (cm & acc_tapmask(acc0 + jl)) started |= 1u << jl;
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (!(p.dbg & 32))
    for (int jl = 0; jl < n_acc; ++jl) {
      if (!(started & (1u << jl))) continue;
      const int slot = 4 * (acc0 + jl) + quarter;
      if (slot >= n_slots) continue;
      const int tap = slot / spt;
      const int cg = ((slot - tap * spt) << 5) + lane;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + jl * p.N + c0, v);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

}  // namespace tma

// ------------------------------------------------------------------------------------------ host launchers
bool conv_tma_supported(const ConvArgs &a) {
  return !a.scatter && a.c_in >= 32 && a.c_in % 32 == 0 && a.c_out >= 32 && a.c_out % 32 == 0 && a.V <= 32 &&
         tma::pick_tn(a.c_out) > 0 && a.in_rows > 0 && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}

void conv_tma(const ConvArgs &a, cudaStream_t s) {
  using namespace tma;
  SCN_CHECK(a.weight_nk != nullptr, "conv_tma needs the [V][Cout][Cin] weight layout");
  if (a.n_rows == 0) return;
  ConvParams p;
  p.bias = a.bias; p.out = a.out; p.tbl = a.tbl; p.tbl_stride = a.tbl_stride; p.n_rows = a.n_rows;
  p.in_rows = a.in_rows; p.V = a.V; p.c_in = a.c_in; p.c_out = a.c_out; p.mirror = a.mirror ? 1 : 0;
  p.TN = pick_tn(a.c_out);
  p.b_stage = p.TN * 128;
  const int stage = A_STAGE + p.b_stage;
  const int fixed = 1024 + a.V * TM * (int)sizeof(int) + 256;
  p.stages = (112 * 1024 - fixed) / stage;         // two CTAs per SM
  if (p.stages > 6) p.stages = 6;
  if (p.stages < 2) p.stages = 2;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.TN) p.tmem_cols <<= 1;
  const size_t smem = (size_t)fixed + (size_t)p.stages * stage;
  CUtensorMap mx = make_map(a.in, (uint64_t)a.c_in, (uint64_t)a.in_rows, KCH, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap mw = make_map(a.weight_nk, (uint64_t)a.c_in, (uint64_t)a.V * a.c_out, KCH, (uint32_t)p.TN,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  static size_t configured = 0;
  if (smem > configured) {
    SCN_CUDA(cudaFuncSetAttribute(k_conv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((a.n_rows + TM - 1) / TM, a.c_out / p.TN);
  k_conv_tma<<<grid, NTHREADS, smem, s>>>(mx, mw, p);
  SCN_LAUNCH_CHECK();
}

bool wgrad_tma_supported(const WgradArgs &a) {
  const int cg = a.table_on_a ? a.c_a : a.c_b, cs = a.table_on_a ? a.c_b : a.c_a;
  return a.chunk_mask != nullptr && a.g_rows > 0 && cg >= 32 && cg % 32 == 0 && cs >= 32 && cs % 32 == 0 && a.V <= 32 &&
         tma::pick_tn(cs) > 0 && ((uintptr_t)a.a % 16 == 0) && ((uintptr_t)a.b % 16 == 0);
}

void wgrad_tma(const WgradArgs &a, cudaStream_t s) {
  using namespace tma;
  SCN_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.V * a.c_a * a.c_b, s));
  if (a.n_rows == 0) return;
  WgParams p;
  const float *G = a.table_on_a ? a.a : a.b;
  const float *S = a.table_on_a ? a.b : a.a;
  p.Cg = a.table_on_a ? a.c_a : a.c_b;
  p.Cs = a.table_on_a ? a.c_b : a.c_a;
  p.transpose_out = a.table_on_a ? 0 : 1;
  p.dw = a.dw; p.tbl = a.tbl; p.cmask = a.chunk_mask; p.tbl_stride = a.tbl_stride; p.n_rows = a.n_rows;
  p.g_rows = a.g_rows; p.V = a.V;
  p.N = pick_tn(p.Cs);
  const int n_slots = a.V * (p.Cg / 32);
  p.n_acc_total = (n_slots + 3) / 4;
  const int max_acc = 512 / p.N;
  const int groups = (p.n_acc_total + max_acc - 1) / max_acc;
  p.acc_per_cta = (p.n_acc_total + groups - 1) / groups;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.acc_per_cta * p.N) p.tmem_cols <<= 1;
  const int n_tiles_n = p.Cs / p.N;
  int row_splits = sm_count() / (groups * n_tiles_n);
  if (row_splits < 1) row_splits = 1;
  int rows = (a.n_rows + row_splits - 1) / row_splits;
  if (rows < 512) rows = 512;
  p.rows_per_cta = (rows + KR - 1) / KR * KR;
  row_splits = (a.n_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  p.s_stage = (p.N / 32) * SUB;
  p.s_stages = p.s_stage <= 16384 ? 3 : 2;
  const int fixed = 1024 + 8 * (2 * 8 + 2 * 3 + 1) + 64 + 4 * 2 * a.V * KR * (int)sizeof(int);
  p.g_stages = (200 * 1024 - fixed - p.s_stages * p.s_stage) / G_STAGE;
  if (p.g_stages > 8) p.g_stages = 8;
  SCN_CHECK(p.g_stages >= 2, "wgrad_tma: shared memory budget");
  const size_t smem = (size_t)fixed + (size_t)p.g_stages * G_STAGE + (size_t)p.s_stages * p.s_stage;
  CUtensorMap mg = make_map(G, (uint64_t)p.Cg, (uint64_t)a.g_rows, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  CUtensorMap ms = make_map(S, (uint64_t)p.Cs, (uint64_t)a.n_rows, 32, KR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  static size_t configured = 0;
  if (smem > configured) {
    SCN_CUDA(cudaFuncSetAttribute(k_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  {
    const char *e = getenv("SCN_WG_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  dim3 grid(row_splits, groups, n_tiles_n);
  k_wgrad_tma<<<grid, NTHREADS, smem, s>>>(mg, ms, p);
  SCN_LAUNCH_CHECK();
}

}  // namespace scn

// TMA-fed tcgen05 table convolutions for sm_100a (SCN_TF32 path, default).
//
// Same math and tiling as conv_tc.cu, but the operands are moved by the Tensor Memory Accelerator instead
// of by per-thread cp.async:
//   * gathered rows:  cp.async.bulk.tensor.2d ... tile::gather4 -- ONE instruction fetches four arbitrary
//     rows (128 bytes each) of the feature matrix into four consecutive swizzled shared-memory rows; an
//     absent neighbour is requested as the out-of-bounds row index `rows`, which TMA zero-fills;
//   * weights / stationary rows: ordinary 2-D tile loads.
// Completion is counted in bytes on an mbarrier (expect_tx); the MMA thread consumes the stage and frees it
// with tcgen05.commit.
//
// What the micro-benchmarks (tools/ubench_pipe.cu, profiles/r01_ubench.md) say about this machine, and what the
// kernels do about it:
//   * one thread issues a gather4 every ~28 clk (58 with a dependent shared-memory load in between); the TMA
//     engine itself keeps up with at least 8 issuing threads per SM (56 B/clk/SM).  So the copies of ONE pipeline
//     item are issued by ONE elected thread with the indices already in registers, and consecutive items go to
//     different producer warps (item-interleaved producers), which also post the item's byte count themselves;
//   * a row index past the end of the tensor is zero-filled, but costs ~12-14 clk per row against 3.8 clk for a
//     real row.  Neither kernel asks for such rows any more: the forward/dgrad kernel masks absent neighbours
//     with tcgen05.mma's disable-output-lane operand (one bit per accumulator row) and does not fetch them at
//     all; the weight-gradient kernel walks per-tap compacted rule lists, so every fetched row is a real rule;
//   * a pipeline item costs ~300 clk of fixed hand-off per producer thread, so items are kept at >= 16 KB.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstdio>

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
constexpr int NTHREADS = 160;     // warp 0: TMEM owner + MMA issuer; warps 1-4: table set-up, TMA producers, epilogue (TMEM quarter = warp & 3)

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
// same, with the disable-output-lane vector: bit i of m[j] set => accumulator row 32*j+i is NOT updated
__device__ __forceinline__ void mma_tf32_masked(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accum, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// zero 32 consecutive columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}

// =====================================================================================================
// forward / dgrad:  out[o,:] = sum_k in[tbl[k][o],:] * W[k]          (output stationary, 128 rows x TN per CTA)
// =====================================================================================================
struct ConvParams {
  const float *in;
  const float *bias;
  float *out;
  const int *tbl;
  int tbl_stride, n_rows, in_rows, V, c_in, c_out, mirror;
  int TN, stages, nprod, b_stage, tmem_cols, prefetch;
  long long *trace;     // SCN_TRACE=1: per-CTA clock64 breakdown of sampled tiles (debug only)
};
#define TRACE_ON (p.trace != nullptr && (blockIdx.x & 127) == 5 && blockIdx.y == 0)
#define TRACE_PUT(i, v) do { if (TRACE_ON) p.trace[(blockIdx.x >> 7) * 16 + (i)] = (v); } while (0)

__global__ void __launch_bounds__(NTHREADS) k_conv_tma(const __grid_constant__ CUtensorMap map_x,
                                                       const __grid_constant__ CUtensorMap map_w, ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + p.stages * A_STAGE;
  int *s_idx = reinterpret_cast<int *>(smem + p.stages * (A_STAGE + p.b_stage));     // [V][TM] row to fetch
  uint32_t *s_pm = reinterpret_cast<uint32_t *>(s_idx + p.V * TM);                    // [V][4] present-row bits
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_pm + p.V * 4);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t accum_bar = empty_bar + 8 * p.stages;
  const uint32_t zero_bar = accum_bar + 8;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 2);
  uint32_t *s_mask = s_tmem + 1;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler
  const int row0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * p.TN;
  const int KC = p.c_in / KCH;
  const long long t_entry = clock64();

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    mbar_init(zero_bar, 4);
    *s_mask = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
  } else {
    // Rows to fetch for every tap.  An absent neighbour is never used by the MMA (its accumulator row is
    // masked), so it is not fetched either: inside a 4-row gather group that has at least one present row the
    // absent ones repeat a present row of the group (a cache hit), and groups with no present row are skipped.
    const int e = tid - 32;
    const int r = row0 + e;
    const int pw = warp - 1;
    const int gsh = lane & ~3;
    uint32_t mine = 0;
    // all V table entries of this row first (independent loads, one memory latency), then the warp votes
    int tv[32];
#pragma unroll
    for (int k = 0; k < 32; ++k)
      tv[k] = (k < p.V && r < p.n_rows) ? __ldg(&p.tbl[(long long)k * p.tbl_stride + r]) : -1;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < p.V) {
        const int t = tv[k];
        const uint32_t pm = __ballot_sync(0xffffffffu, t >= 0);
        const uint32_t gm = (pm >> gsh) & 0xFu;
        const int src = gsh + (gm ? __ffs(gm) - 1 : 0);
        const int trep = __shfl_sync(0xffffffffu, t, src);
        s_idx[k * TM + e] = t >= 0 ? t : (gm ? trep : 0);
        if (lane == 0) s_pm[k * 4 + pw] = pm;
        mine |= (pm ? 1u : 0u) << k;
        if (t >= 0 && p.prefetch) {
          const char *row = reinterpret_cast<const char *>(p.in + (long long)t * p.c_in);
          for (int b = 0; b < p.c_in * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
        }
      }
    }
    if (lane == 0 && mine) atomicOr(s_mask, mine);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t tapmask = *s_mask;
  if (tid == 0) { TRACE_PUT(0, t_entry); TRACE_PUT(1, clock64() - t_entry); TRACE_PUT(8, (long long)__popc(tapmask) * KC); }

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      long long w_full = 0;
      const uint32_t idesc = idesc_tf32(p.TN, 0, 0);
      mbar_wait(zero_bar, 0);          // the accumulator starts as zeros (every MMA accumulates; rows can be masked from the first tap on)
      tc_fence_after();
      uint32_t remaining = tapmask;
      int s = 0;
      uint32_t ph = 0;
      while (remaining) {
        const int trow = __ffs(remaining) - 1;
        remaining &= remaining - 1;
        const uint4 pm = *reinterpret_cast<const uint4 *>(s_pm + trow * 4);
        for (int kc = 0; kc < KC; ++kc) {
          const long long tw = clock64();
          mbar_wait(full_bar + 8 * s, ph);
          w_full += clock64() - tw;
          tc_fence_after();
          const uint64_t ad = desc_k128(a_base + s * A_STAGE);
          const uint64_t bd = desc_k128(b_base + s * p.b_stage);
#pragma unroll
          for (int k = 0; k < KCH / 8; ++k)
            mma_tf32_masked(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u, ~pm.x, ~pm.y, ~pm.z, ~pm.w);
          mma_commit(empty_bar + 8 * s);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
      mma_commit(accum_bar);
      TRACE_PUT(2, w_full); TRACE_PUT(3, clock64() - t_entry);
    }
  } else {
    const int quarter = warp & 3;                     // TMEM lanes 32*quarter .. +31 belong to this warp
    const uint32_t tq = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int c0 = 0; c0 < p.TN; c0 += 32) tmem_zero32(tq + c0);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(zero_bar);
    // =========================== TMA producers, item-interleaved ===========================
    // Producer pw owns items pw, pw+nprod, ... (stages % nprod == 0, so a stage always belongs to the same
    // producer): it waits for the stage, posts the byte count and issues the weight tile and all gather4 copies
    // of the item from one thread with the row indices already in registers.
    const int pw = warp - 1;
    if (pw < p.nprod && elect_one()) {
      uint32_t remaining = tapmask;
      int s = -1, turn = -1;
      long long w_empty = 0, t_issue = 0;
      uint32_t ph = 1;                 // parity to wait for on the empty barrier (first pass: already free)
      while (remaining) {
        const int trow = __ffs(remaining) - 1;
        remaining &= remaining - 1;
        const int wtap = p.mirror ? p.V - 1 - trow : trow;
        for (int kc = 0; kc < KC; ++kc) {
          if (++s == p.stages) { s = 0; ph ^= 1u; }
          if (++turn == p.nprod) turn = 0;
          if (turn != pw) continue;
          const uint4 pm = *reinterpret_cast<const uint4 *>(s_pm + trow * 4);
          const uint32_t pmq[4] = {pm.x, pm.y, pm.z, pm.w};
          uint32_t nz[4];
          int groups = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            nz[q] = (pmq[q] | (pmq[q] >> 1) | (pmq[q] >> 2) | (pmq[q] >> 3)) & 0x11111111u;
            groups += __popc(nz[q]);
          }
          const long long tw = clock64();
          mbar_wait(empty_bar + 8 * s, ph);
          const long long ti = clock64();
          w_empty += ti - tw;
          mbar_expect_tx(full_bar + 8 * s, (uint32_t)(p.b_stage + groups * 512));
          tma_tile_2d(b_base + s * p.b_stage, &map_w, kc * KCH, wtap * p.c_out + n0, full_bar + 8 * s);
          const uint32_t dst = a_base + s * A_STAGE;
          const int4 *rows4 = reinterpret_cast<const int4 *>(s_idx + trow * TM);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (!nz[q]) continue;
            int4 r[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) r[g] = rows4[q * 8 + g];
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if ((nz[q] >> (4 * g)) & 1u)
                tma_gather4(dst + (q * 8 + g) * 512, &map_x, kc * KCH, r[g].x, r[g].y, r[g].z, r[g].w, full_bar + 8 * s);
          }
          t_issue += clock64() - ti;
        }
      }
      if (pw == 0) { TRACE_PUT(4, w_empty); TRACE_PUT(5, clock64() - t_entry); TRACE_PUT(9, t_issue); }
    }
    __syncwarp();
    // =========================== epilogue ===========================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (tid == 32) TRACE_PUT(6, clock64() - t_entry);
    const int r = row0 + quarter * 32 + lane;
    float *orow = p.out + (long long)r * p.c_out + n0;
    for (int c0 = 0; c0 < p.TN; c0 += 32) {
      float v[32];
      tmem_ld32(tq + c0, v);
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
  if (tid == 0) TRACE_PUT(7, clock64() - t_entry);
  if (warp == 0) {
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
// weight gradient over compacted rule lists:
//   D_k[Cg x Cs] = sum over the rules (g, s) of tap k of  G[g,:]^T S[s,:]
// The rules of tap k occupy items item_off[k] .. item_off[k+1]-1 of the lists gi/si (32 rules per item, ordered
// by s, the last item of a tap padded with an out-of-range index that TMA zero-fills).  CTA (tap k, row range rr,
// column tile) takes the items of tap k whose s falls in rows [rr*range, (rr+1)*range) -- blk_item holds the item
// index of every 512-row boundary -- keeps D in TMEM (M = 128 channels of G per accumulator tile, N channels of S,
// K = 32 rules per item) and adds it to dW with vector reductions at the end.  The tap index is the FASTEST grid
// dimension: the 27 CTAs of one row range run together, so every G and S row is fetched from HBM about once and
// re-read from L2 by the other taps (a flat split of the rule lists reads every row from HBM once per tap).
// Both operands are MN-major (the contraction runs over rules = rows), which tcgen05 accepts for tf32 only in
// the SWIZZLE_128B_BASE32B layout: 32-channel atoms of [32 rules][128 B], 4-rule groups 512 B apart.
// =====================================================================================================
constexpr int KR = 32;            // rules (GEMM K) per pipeline item
constexpr int SUB = KR * 128;     // bytes of one [KR rules x 32 channels] atom
constexpr int WG_THREADS = 288;   // warp 0: TMEM + MMA; warps 1-4: producers; warps 5-8: epilogue (TMEM quarter = warp & 3)

struct WgParams {
  float *dw;
  const int *gi, *si, *blk_item;
  int V, Cg, Cs, transpose_out, n_blk, blk_per_range;
  int N, m_tiles, stages, nprod, stage_bytes, tmem_cols;
};

__global__ void __launch_bounds__(WG_THREADS) k_wgrad_tma(const __grid_constant__ CUtensorMap map_g,
                                                          const __grid_constant__ CUtensorMap map_s, WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  int *s_pidx = reinterpret_cast<int *>(smem + p.stages * p.stage_bytes);     // [4 producers][64]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_pidx + 4 * 64);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t accum_bar = empty_bar + 8 * p.stages;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g_atoms = p.Cg >> 5, s_atoms = p.N >> 5;
  const int tap = blockIdx.x;
  const int n0 = blockIdx.z * p.N;
  const int b0 = blockIdx.y * p.blk_per_range;
  const int b1 = min(b0 + p.blk_per_range, p.n_blk);
  const int ib = __ldg(&p.blk_item[tap * (p.n_blk + 1) + b0]);
  const int ie = __ldg(&p.blk_item[tap * (p.n_blk + 1) + b1]);
  if (ib >= ie) return;           // uniform: nothing of this tap in this row range

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) tmem_alloc(s_tmem, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(p.N, 1, 1);
      int s = 0;
      uint32_t ph = 0;
      for (int item = ib; item < ie; ++item) {
        mbar_wait(full_bar + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = base + s * p.stage_bytes;
        const uint64_t bd = desc_mn32(st + g_atoms * SUB, SUB);
        for (int mt = 0; mt < p.m_tiles; ++mt) {
          const uint64_t ad = desc_mn32(st + mt * 4 * SUB, SUB);
#pragma unroll
          for (int kk = 0; kk < KR / 8; ++kk)    // K = 8 rules per MMA = two 512-byte K groups of every atom
            mma_tf32(tmem + mt * p.N, ad + (uint64_t)(kk * 64), bd + (uint64_t)(kk * 64), idesc,
                     (item > ib || kk) ? 1u : 0u);
        }
        mma_commit(empty_bar + 8 * s);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      mma_commit(accum_bar);
    }
  } else if (warp <= 4) {
    // =========================== TMA producers, item-interleaved ===========================
    // The whole warp prefetches the next item's 2 x 32 indices (two coalesced 128-byte loads) one item ahead;
    // lane 0 waits for the stage, posts the byte count and issues the item's copies.
    const int pw = warp - 1;
    if (pw < p.nprod) {
      int *my = s_pidx + pw * 64;
      int it = ib + pw;
      int s = pw;                       // stage of item `it` (stages % nprod == 0: a producer cycles over its own stages)
      uint32_t ph = 1;
      int g_next = 0, s_next = 0;
      if (it < ie) {
        g_next = __ldg(&p.gi[(long long)it * KR + lane]);
        s_next = __ldg(&p.si[(long long)it * KR + lane]);
      }
      for (; it < ie; it += p.nprod) {
        my[lane] = g_next;
        my[32 + lane] = s_next;
        if (it + p.nprod < ie) {
          g_next = __ldg(&p.gi[(long long)(it + p.nprod) * KR + lane]);
          s_next = __ldg(&p.si[(long long)(it + p.nprod) * KR + lane]);
        }
        __syncwarp();
        if (lane == 0) {
          int4 rg[8], rs[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            rg[g] = reinterpret_cast<const int4 *>(my)[g];
            rs[g] = reinterpret_cast<const int4 *>(my + 32)[g];
          }
          mbar_wait(empty_bar + 8 * s, ph);
          mbar_expect_tx(full_bar + 8 * s, (uint32_t)p.stage_bytes);
          const uint32_t dst = base + s * p.stage_bytes;
          for (int a = 0; a < g_atoms; ++a) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              tma_gather4(dst + a * SUB + g * 512, &map_g, a * 32, rg[g].x, rg[g].y, rg[g].z, rg[g].w, full_bar + 8 * s);
          }
          for (int a = 0; a < s_atoms; ++a) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              tma_gather4(dst + (g_atoms + a) * SUB + g * 512, &map_s, n0 + a * 32, rs[g].x, rs[g].y, rs[g].z, rs[g].w,
                          full_bar + 8 * s);
          }
        }
        s += p.nprod;
        if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue: TMEM -> fp32 reductions into dW ===========================
    const int quarter = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t acc = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int mt = 0; mt < p.m_tiles; ++mt) {
      const int cg = mt * 128 + quarter * 32 + lane;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(acc + mt * p.N + c0, v);
        if (cg < p.Cg) {
          if (!p.transpose_out) {
            float *dst = p.dw + ((long long)tap * p.Cg + cg) * p.Cs + n0 + c0;
#pragma unroll
            for (int q = 0; q < 32; q += 4) red_add_v4(dst + q, v[q], v[q + 1], v[q + 2], v[q + 3]);
          } else {
            float *dst = p.dw + ((long long)tap * p.Cs + n0 + c0) * p.Cg + cg;
#pragma unroll
            for (int q = 0; q < 32; ++q) atomicAdd(dst + (long long)q * p.Cg, v[q]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

}  // namespace tma

// ------------------------------------------------------------------------------------------ host launchers
static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

bool conv_tma_supported(const ConvArgs &a) {
  return !a.scatter && a.c_in >= 32 && a.c_in % 32 == 0 && a.c_out >= 32 && a.c_out % 32 == 0 && a.V <= 32 &&
         tma::pick_tn(a.c_out) > 0 && a.in_rows > 0 && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}

void conv_tma(const ConvArgs &a, cudaStream_t s) {
  using namespace tma;
  SCN_CHECK(a.weight_nk != nullptr, "conv_tma needs the [V][Cout][Cin] weight layout");
  if (a.n_rows == 0) return;
  ConvParams p;
  static const int prefetch = env_int("SCN_CONV_PREFETCH", 1);
  p.prefetch = prefetch;
  p.in = a.in; p.bias = a.bias; p.out = a.out; p.tbl = a.tbl; p.tbl_stride = a.tbl_stride; p.n_rows = a.n_rows;
  p.in_rows = a.in_rows; p.V = a.V; p.c_in = a.c_in; p.c_out = a.c_out; p.mirror = a.mirror ? 1 : 0;
  p.TN = pick_tn(a.c_out);
  p.b_stage = p.TN * 128;
  const int stage = A_STAGE + p.b_stage;
  const int fixed = 1024 + a.V * TM * (int)sizeof(int) + a.V * 16 + 8 * (2 * 8 + 2) + 64;
  // two CTAs per SM (one's epilogue and table set-up overlap the other's main loop) when that leaves >= 3 stages
  static const int force_ctas = env_int("SCN_CONV_CTAS", 0);
  int budget = 113 * 1024;
  if (force_ctas == 1 || (force_ctas == 0 && (budget - fixed) / stage < 3)) budget = 225 * 1024;
  int st = (budget - fixed) / stage;
  st = st >= 8 ? 8 : st >= 6 ? 6 : st >= 4 ? 4 : st >= 3 ? 3 : 2;
  p.stages = st;
  p.nprod = st % 4 == 0 ? 4 : st % 3 == 0 ? 3 : 2;
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
  static const int trace = env_int("SCN_TRACE", 0);
  p.trace = nullptr;
  const int n_tr = (int)(grid.x >> 7);
  if (trace && n_tr > 0) {
    SCN_CUDA(cudaMalloc((void **)&p.trace, sizeof(long long) * 16 * (n_tr + 1)));
    SCN_CUDA(cudaMemset(p.trace, 0, sizeof(long long) * 16 * (n_tr + 1)));
  }
  k_conv_tma<<<grid, NTHREADS, smem, s>>>(mx, mw, p);
  SCN_LAUNCH_CHECK();
  if (p.trace) {
    std::vector<long long> h(16 * (n_tr + 1));
    SCN_CUDA(cudaStreamSynchronize(s));
    SCN_CUDA(cudaMemcpy(h.data(), p.trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    double m[16] = {0};
    for (int i = 0; i < n_tr; ++i) for (int j = 1; j < 16; ++j) m[j] += (double)h[i * 16 + j] / n_tr;
    fprintf(stderr, "[conv trace] rows %d C %d->%d TN %d stages %d nprod %d | tiles sampled %d: setup %.0f, items %.1f, "
            "mma wait-full %.0f, mma done %.0f | prod0 wait-empty %.0f, issue %.0f, done %.0f | accum %.0f, end %.0f clk\n",
            a.n_rows, a.c_in, a.c_out, p.TN, p.stages, p.nprod, n_tr, m[1], m[8], m[2], m[3], m[4], m[9], m[5], m[6], m[7]);
  }
}

bool wgrad_tma_supported(const WgradArgs &a) {
  const int cg = a.table_on_a ? a.c_a : a.c_b, cs = a.table_on_a ? a.c_b : a.c_a;
  return a.gi != nullptr && a.g_rows > 0 && a.s_rows > 0 && cg >= 32 && cg % 32 == 0 && cs >= 32 && cs % 32 == 0 &&
         a.V <= 32 && ((uintptr_t)a.a % 16 == 0) && ((uintptr_t)a.b % 16 == 0);
}

void wgrad_tma(const WgradArgs &a, cudaStream_t s) {
  using namespace tma;
  SCN_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.V * a.c_a * a.c_b, s));
  if (a.n_rows == 0 || a.n_blk == 0) return;
  WgParams p;
  const float *G = a.table_on_a ? a.a : a.b;
  const float *S = a.table_on_a ? a.b : a.a;
  p.Cg = a.table_on_a ? a.c_a : a.c_b;
  p.Cs = a.table_on_a ? a.c_b : a.c_a;
  p.transpose_out = a.table_on_a ? 0 : 1;
  p.dw = a.dw; p.gi = a.gi; p.si = a.si; p.blk_item = a.blk_item; p.V = a.V; p.n_blk = a.n_blk;
  p.m_tiles = (p.Cg + 127) / 128;
  p.N = 0;
  for (int n = 256; n >= 32; n -= 32)
    if (p.Cs % n == 0 && p.m_tiles * n <= 512) { p.N = n; break; }
  SCN_CHECK(p.N > 0, "wgrad_tma: no N tile");
  p.tmem_cols = 32;
  while (p.tmem_cols < p.m_tiles * p.N) p.tmem_cols <<= 1;
  // every stage holds the item's G atoms, then its S atoms; an accumulator tile always spans 4 G atoms, so the
  // last tile of a Cg that is not a multiple of 128 reads on into the S atoms (finite data; rows >= Cg are ignored)
  p.stage_bytes = (p.Cg / 32 + p.N / 32) * SUB;
  const int fixed = 1024 + 4 * 64 * 4 + 8 * (2 * 8 + 1) + 64;
  const int tail = p.m_tiles * 128 > p.Cg ? 4 * SUB : 0;
  int budget = 113 * 1024;
  if (512 / p.tmem_cols < 2 || (budget - fixed - tail) / p.stage_bytes < 3) budget = 225 * 1024;
  int st = (budget - fixed - tail) / p.stage_bytes;
  SCN_CHECK(st >= 2, "wgrad_tma: shared memory budget");
  st = st >= 8 ? 8 : st >= 6 ? 6 : st >= 4 ? 4 : st >= 3 ? 3 : 2;
  p.stages = st;
  p.nprod = st % 4 == 0 ? 4 : st % 3 == 0 ? 3 : 2;
  const size_t smem = (size_t)fixed + (size_t)p.stages * p.stage_bytes + tail;
  // row range per CTA: about 8 MB of G + S rows, so the ranges in flight (SMs x CTAs/SM / V of them) fit in L2
  static const int range_kb = env_int("SCN_WG_RANGE_KB", 8192);
  long long rows = (long long)range_kb * 1024 / ((long long)(p.Cg + p.Cs) * 4);
  int bpr = (int)(rows / a.blk_rows);
  if (bpr < 1) bpr = 1;
  p.blk_per_range = bpr;
  const int ranges = (a.n_blk + bpr - 1) / bpr;
  CUtensorMap mg = make_map(G, (uint64_t)p.Cg, (uint64_t)a.g_rows, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  CUtensorMap ms = make_map(S, (uint64_t)p.Cs, (uint64_t)a.s_rows, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  static size_t configured = 0;
  if (smem > configured) {
    SCN_CUDA(cudaFuncSetAttribute(k_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(a.V, ranges, p.Cs / p.N);
  k_wgrad_tma<<<grid, WG_THREADS, smem, s>>>(mg, ms, p);
  SCN_LAUNCH_CHECK();
}

}  // namespace scn

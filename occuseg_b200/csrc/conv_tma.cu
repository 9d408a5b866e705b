// tcgen05 + TMA table convolutions for sm_100a: the SCN_BF16 (default) and SCN_TF32 paths.
//
//   k_conv_tma  : forward / dgrad / strided / deconvolution as an output-stationary gather-GEMM
//   k_wgrad_tma : weight gradients over per-tap compacted rule lists
//
// Operands are moved by the Tensor Memory Accelerator:
//   * gathered rows:  cp.async.bulk.tensor.2d ... tile::gather4 -- ONE instruction fetches four arbitrary rows
//     (128 bytes each = 64 bf16 or 32 tf32 channels) of a feature matrix into four consecutive swizzled
//     shared-memory rows;
//   * weights: ordinary 2-D tile loads.
// Completion is counted in bytes on an mbarrier (expect_tx); one elected thread issues tcgen05.mma (fp32
// accumulators in TMEM) and frees the stage with tcgen05.commit.  Replaces the reference's scalar shared-memory FMA
// kernels with global atomics (CUDA/Convolution.cu:447-534,695-753,1059-1152; CUDA/Deconvolution.cu:9-554).
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
// -DSCN_TRACE_BUILD compiles clock64 accounting into k_conv_tma (SCN_TRACE=1 prints it); off in normal builds
#ifdef SCN_TRACE_BUILD
#define TRC(x) x
#else
#define TRC(x)
#endif

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

// row-major fp32 or bf16 matrix [rows, cols]; box = box_cols x box_rows elements
static CUtensorMap make_map(const void *base, bool bf16, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                            CUtensorMapSwizzle sw) {
  EncodeFn fn = encode_fn();
  SCN_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base),
                  dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return m;
}

// ------------------------------------------------------------------------------------------ device helpers
constexpr int TM = 128;
constexpr int A_STAGE = TM * 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits pass a suspend-time hint to mbarrier.try_wait: the thread is parked by the hardware until the phase completes (or the
// hint, 10 ms, runs out) instead of re-issuing the instruction.  Without it try_wait returns after a few cycles, and the spinning
// waiters -- 110 M try_wait + branch pairs per level-0 launch (ncu source page, profiles/r01_ncu_l0_stalls.txt) -- took about
// 30 % of the issue slots away from the producers' single-thread instruction chains.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
}
// long waits (epilogue warps): one lane waits, the rest of the warp sleeps at the warp barrier
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
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
// MN-major SWIZZLE_128B (16-bit types): 64-channel blocks `lbo` bytes apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mn128(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (InstrDescriptor): fp32 accumulate, M=128, N=n; operand format 2 = tf32 (kind::tf32),
// 1 = bf16 (kind::f16); majors: 0 = K, 1 = MN
__device__ __forceinline__ uint32_t idesc_make(int n, uint32_t a_mn, uint32_t b_mn, bool bf16) {
  const uint32_t fmt = bf16 ? 1u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_masked(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accum, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
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
// forward / dgrad:  out[o,:] = sum_k in[tbl[k][o],:] * W[k]          (output stationary)
//
// Persistent CTAs (one per SM).  A tile group = MT consecutive 128-row tiles x TN output channels; its MT fp32
// accumulators live in TMEM for the whole walk over the taps (double buffered: the epilogue of one group runs
// under the main loop of the next).  A pipeline item = (group, tap, 32-channel K chunk): ONE [TN x 32] weight
// tile shared by the MT row tiles, plus the rows each of them needs.  There is no set-up phase: the producer warp
// that owns an item reads the tap's table row for the group straight from global memory (one coalesced 16-byte
// load per lane = 128 rows, prefetched one item ahead), votes the present-row bits, and one lane issues the
// copies.  The bits travel to the MMA thread in shared memory next to the stage and become tcgen05.mma's
// disable-output-lane mask; taps with no row in the group are passed as empty items.
// =====================================================================================================
// warp 0: TMEM + MMA issuer; warps 1 .. nprod*ni: producers (item owner = (warp-1)/ni, share of the item's copies =
// (warp-1)%ni); the last EPI_WARPS warps: epilogue, EPI_WARPS/4 per TMEM quarter (quarter = warp & 3), alternating 16-column chunks
// EPI_WARPS = 4 (one per TMEM quarter).  8 (two per quarter, alternating chunks) is supported by the code below but measured
// the same step time within noise, and 800 threads leave only 72 registers per thread
constexpr int EPI_WARPS = 4;
constexpr int EPI_SPLIT = EPI_WARPS / 4;
constexpr int MAX_PROD_WARPS = 16;     // 4 owner slots x 4 shares (5 slots = 800 threads: 72 registers, spills, measured 25 % slower)
constexpr int CONV_MAX_THREADS = 32 + MAX_PROD_WARPS * 32 + 32 * EPI_WARPS;
constexpr int MAX_MT = 2;

struct ConvParams {
  const float *bias;
  const float *residual;       // optional [n_rows, c_out] fp32 added to the result (residual shortcut fused into the epilogue)
  double *stats;               // optional [2][c_out]: column sums and sums of squares of the result (for the BatchNorm that follows)
  long long stats_stride;      // deterministic mode: epilogue warp q of CTA x adds into stats + (4x + q) * stats_stride (zeroed scratch)
  // optional fused inference BatchNorm + (leaky) ReLU of the layer that follows: out = leaky(scale[c] * acc + shift[c]),
  // plus an optional bf16 copy of that result for the next tensor-core convolution
  const float *ep_scale, *ep_shift;
  float ep_leak;
  uint16_t *out_bf16;
  // optional fused BatchNorm BACKWARD of the layer that produced this product's input operand (dgrad products): the result
  // row r is the gradient w.r.t. that BatchNorm's output; bnb_x = the BatchNorm's input [rows, c_out], bnb_mean / bnb_invstd its
  // saved statistics, bnb_gamma / bnb_beta its affine parameters (w = invstd*gamma, b = beta - mean*w as its forward evaluated them).  The epilogue applies the activation mask recomputed from x, writes the masked
  // gradient d', and accumulates the column sums of d' and d'*x into `stats` instead of (sum, sum of squares).
  const float *bnb_x, *bnb_mean, *bnb_invstd, *bnb_gamma, *bnb_beta;     // gamma / beta may be NULL (no affine)
  float bnb_leak;
  float *out;
  const int *tbl;
  int tbl_stride, n_rows, V, c_in, c_out, mirror;
  int TN, MT, stages, nprod, ni, b_stage, stage_bytes, tmem_cols, n_groups;
  int bf16, kelems;            // operand type; elements per 128-byte K chunk (32 tf32 / 64 bf16)
  const void *in;              // base of the gathered matrix (L2 prefetch of upcoming rows)
  // one-tap-per-row form (see ConvArgs): weight tap of a tile group from item_off, result rows scattered through out_rows
  const int *out_rows, *item_off;
  int rows_per_item, n_taps, out_limit;
  // optional [tiles] bit t set = table row t has at least one row in that 128-row tile: taps with no row in a tile group
  // never enter the pipeline (NULL: all V taps are walked)
  const uint32_t *tile_mask;
  int n_tiles;
  int dbg;                     // SCN_CONV_DBG (timing experiments only, results are WRONG): 1 = no weight-tile copies, 2 = MMAs of half the width, 4 = no MMAs, 8 = no epilogue stores, 16 = no row copies
  int prefetch;                // pull the next item's feature rows towards L2 (prefetch.global.L2) while the current one is issued
  unsigned long long *trace;   // SCN_TRACE=1: clock64 totals over all CTAs (debug only): see conv_tma()
};

__global__ void __launch_bounds__(CONV_MAX_THREADS, 1) k_conv_tma(const __grid_constant__ CUtensorMap map_x,
                                                              const __grid_constant__ CUtensorMap map_w, ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  uint4 *s_masks = reinterpret_cast<uint4 *>(smem + p.stages * p.stage_bytes);        // [stages][MAX_MT] present-row bits
  int4 *s_rows = reinterpret_cast<int4 *>(s_masks + 8 * MAX_MT);                       // [producer warps][MT*32/ni] gather groups each warp issues
  uint32_t *s_words = reinterpret_cast<uint32_t *>(s_rows + p.nprod * p.MT * 32);      // [producer warps][8] present-row words + destination slots of the item being prepared
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_words + MAX_PROD_WARPS * 8);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t accf_bar = empty_bar + 8 * p.stages;       // [2] accumulator buffer complete
  const uint32_t acce_bar = accf_bar + 16;                  // [2] accumulator buffer drained and zeroed
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 4);
  float4 *s_tile = reinterpret_cast<float4 *>(s_tmem + 4);                              // [EPI_WARPS][32 rows][4] transpose tiles (2 KB each)
  float *s_stat = reinterpret_cast<float *>(s_tile + EPI_WARPS * 128);                  // [EPI_WARPS][2][256] column statistics of this CTA

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler
  const int n0 = blockIdx.y * p.TN;
  const int KC = p.c_in / p.kelems;
  const int G = gridDim.x;
  const int acc_cols = p.MT * p.TN;
  // table rows (taps) present in tile group tg; every role derives the group's item list from it
  auto group_mask = [&](int tg_) -> uint32_t {
    if (tg_ >= p.n_groups) return 0u;
    if (!p.tile_mask) return p.V >= 32 ? 0xFFFFFFFFu : ((1u << p.V) - 1u);
    uint32_t mk = 0u;
    for (int m = 0; m < p.MT; ++m) {
      const int tile = tg_ * p.MT + m;
      if (tile < p.n_tiles) mk |= __ldg(&p.tile_mask[tile]);
    }
    return mk;
  };

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, p.ni);       // every share of an item arrives: no producer warp can fall a phase behind
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar + 8 * b, 1);
      mbar_init(acce_bar + 8 * b, EPI_WARPS);
    }
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
      const uint32_t idesc = idesc_make((p.dbg & 2) ? p.TN / 2 : p.TN, 0, 0, p.bf16 != 0);
      int s = 0, gi = 0;
      uint32_t ph = 0;
      TRC(long long w_full = 0; long long w_acc = 0; long long n_tiles_mma = 0; long long t_mma = 0; long long t_commit = 0; const long long t_begin = clock64();)
      for (int tg = blockIdx.x; tg < p.n_groups; tg += G, ++gi) {
        const int buf = gi & 1;
        TRC(const long long ta = clock64();)
        mbar_wait(acce_bar + 8 * buf, (uint32_t)(gi >> 1) & 1u);      // zeroed by the epilogue warps (or at start)
        TRC(w_acc += clock64() - ta;)
        tc_fence_after();
        const uint32_t acc = tmem + buf * acc_cols;
        const int n_items = __popc(group_mask(tg)) * KC;
        for (int j = 0; j < n_items; ++j) {
          TRC(const long long tw = clock64();)
          mbar_wait(full_bar + 8 * s, ph);
          TRC(w_full += clock64() - tw;)
          tc_fence_after();
          const uint32_t st = base + s * p.stage_bytes;
          const uint64_t bd = desc_k128(st + p.MT * A_STAGE);
          TRC(const long long tm0 = clock64();)
          for (int m = 0; m < p.MT; ++m) {
            const uint4 pm = s_masks[s * MAX_MT + m];
            if ((pm.x | pm.y | pm.z | pm.w) == 0u || (p.dbg & 4)) continue;
            TRC(++n_tiles_mma;)
            const uint64_t ad = desc_k128(st + m * A_STAGE);
            if (p.bf16) {
#pragma unroll
              for (int k = 0; k < 4; ++k)        // 4 x (K = 16 bf16 = 32 bytes of every row)
                mma_bf16_masked(acc + m * p.TN, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u, ~pm.x, ~pm.y, ~pm.z,
                                ~pm.w);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)        // 4 x (K = 8 tf32 = 32 bytes of every row)
                mma_tf32_masked(acc + m * p.TN, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u, ~pm.x, ~pm.y, ~pm.z,
                                ~pm.w);
            }
          }
          TRC(const long long tc0 = clock64();)
          mma_commit(empty_bar + 8 * s);
          TRC(const long long tc1 = clock64(); t_mma += tc1 - tm0; t_commit += tc1 - tc0;)
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        mma_commit(accf_bar + 8 * buf);
      }
      TRC(if (p.trace) {
        atomicAdd(p.trace + 0, (unsigned long long)(clock64() - t_begin));
        atomicAdd(p.trace + 1, (unsigned long long)w_full);
        atomicAdd(p.trace + 2, (unsigned long long)w_acc);
        atomicAdd(p.trace + 3, (unsigned long long)n_tiles_mma);
        atomicAdd(p.trace + 40, (unsigned long long)t_mma);
        atomicAdd(p.trace + 43, (unsigned long long)t_commit);
      })
    }
  } else if (warp <= p.nprod * p.ni) {
    // =========================== TMA producers, item-interleaved ===========================
    // Owner slot pw takes items pw, pw+nprod, ... (nprod <= stages, so a parity wait can never be two phases off);
    // the ni warps of a slot split the item's copies between them because one thread cannot issue them fast
    // enough (tools/ubench_pipe.cu); share 0 also posts the byte count, the masks and the weight tile.
    const int pw = (warp - 1) / p.ni, part = (warp - 1) % p.ni;
    {
      // Every share is self-contained and lies inside ONE row tile (ni = 4 shares, MT <= 2 tiles: a share owns MT consecutive
      // 32-row chunks): it reads that tile's table row, publishes the present-row words of its chunks, posts its own byte count
      // (arrive.expect_tx) and issues its copies; share 0 adds the weight tile.
      // The loop is written for instruction count: an ncu source-level profile of the level-0 launch (round 2) showed the
      // producers executing ~600 instructions per share and item -- 76 % of all instructions of the kernel, on single-thread
      // dependency chains -- and the slots, never blocked on `empty`, setting the pace.  Hence: no division / modulo (the item
      // iterator is incremental: remaining-taps mask + K-chunk counter), the group's gather list is COMPACTED by the lanes that
      // own a present group (the elected lane walks a dense list instead of testing 8 x chunks predicates and reading every
      // slot), and shares no longer read the other tile or build its masks.
      const int m = (part * p.MT) >> 2;              // the row tile of this share
      const int lane_lo = ((part * p.MT) & 3) * 8;   // lanes lane_lo .. lane_lo + 8*MT - 1 hold this share's gather groups
      const bool own = lane >= lane_lo && lane < lane_lo + 8 * p.MT;
      int4 *my_rows = s_rows + (warp - 1) * 8 * p.MT;                         // [8*MT] compacted gather groups of the item
      uint32_t *my_words = s_words + (warp - 1) * 8;                          // [4] chunk words, then [16] destination slots (bytes)
      uint8_t *my_dst = reinterpret_cast<uint8_t *>(my_words + 4);
      auto load_tbl = [&](int tg_, int trow_) -> int4 {
        const int row = (tg_ * p.MT + m) * TM + 4 * lane;
        return row < p.tbl_stride ? __ldg(reinterpret_cast<const int4 *>(p.tbl + (long long)trow_ * p.tbl_stride + row))
                                  : make_int4(-1, -1, -1, -1);
      };
      // ONE iterator, running two items of this slot ahead of the item being issued: tile group, the taps of the group still
      // to come (lowest set bit = the item's table row), K chunk.  The tap mask of the following group is fetched when a group
      // is entered, so that stepping into it does not wait for global memory.
      int tgF = blockIdx.x, kcF = 0;
      uint32_t gmF = group_mask(tgF), gkN = group_mask(tgF + G);
      auto skip_empty = [&]() {
        while (!gmF && tgF < p.n_groups) {
          tgF += G;
          gmF = gkN;
          gkN = group_mask(tgF + G);
        }
      };
      auto step = [&](int n) {           // n items on
        while (tgF < p.n_groups) {       // whole groups first
          const int avail = __popc(gmF) * KC - kcF;      // items from the current one to the end of its group
          if (n < avail) break;
          n -= avail;
          kcF = 0;
          tgF += G;
          gmF = gkN;
          gkN = group_mask(tgF + G);
        }
        kcF += n;
        while (kcF >= KC) {              // at most n taps
          kcF -= KC;
          gmF &= gmF - 1u;
        }
      };
      skip_empty();
      step(pw);
      int tg = tgF, kc = kcF, trow = __ffs((int)gmF) - 1;
      step(p.nprod);
      int tg1 = tgF, kc1 = kcF, trow1 = __ffs((int)gmF) - 1;
      int s = pw;
      uint32_t ph = 1;                   // parity to wait for on the empty barrier (first pass: already free)
      TRC(long long w_empty = 0; long long t_issue = 0; long long n_items = 0; long long n_copies = 0; long long t_pre = 0;
          long long t_b = 0; long long t_ph[3]; t_ph[0] = t_ph[1] = t_ph[2] = 0; const long long t_begin = clock64();)
      int4 cur = make_int4(-1, -1, -1, -1), nxt = cur, far = cur;
      if (tg < p.n_groups) cur = load_tbl(tg, trow);
      if (tg1 < p.n_groups) nxt = load_tbl(tg1, trow1);
      while (tg < p.n_groups) {
        TRC(const long long t_top = clock64();)
        step(p.nprod);
        const int tg2 = tgF, kc2 = kcF, trow2 = __ffs((int)gmF) - 1;
        if (tg2 < p.n_groups) far = load_tbl(tg2, trow2);
        TRC(const long long t_p1 = clock64(); t_ph[0] += t_p1 - t_top; const long long t_p2 = t_p1;)
        // present-row bits; absent rows repeat a present row of their 4-row gather group; groups with a row are compacted
        int4 t = cur;
        const uint32_t nib = (t.x >= 0 ? 1u : 0u) | (t.y >= 0 ? 2u : 0u) | (t.z >= 0 ? 4u : 0u) | (t.w >= 0 ? 8u : 0u);
        const bool have = own && nib != 0u;
        const uint32_t bal = __ballot_sync(0xffffffffu, have);
        // word q of the tile's mask = rows 32q .. 32q+31 = lanes 8q .. 8q+7: every lane ends up with the word of ITS chunk
        uint32_t w8 = nib << (4 * (lane & 7));
        w8 |= __shfl_xor_sync(0xffffffffu, w8, 1);
        w8 |= __shfl_xor_sync(0xffffffffu, w8, 2);
        w8 |= __shfl_xor_sync(0xffffffffu, w8, 4);
        if (have) {
          const int rep = t.x >= 0 ? t.x : t.y >= 0 ? t.y : t.z >= 0 ? t.z : t.w;
          t.x = t.x >= 0 ? t.x : rep;
          t.y = t.y >= 0 ? t.y : rep;
          t.z = t.z >= 0 ? t.z : rep;
          t.w = t.w >= 0 ? t.w : rep;
          const int pos = __popc(bal & ((1u << lane) - 1u));
          my_rows[pos] = t;
          my_dst[pos] = (uint8_t)lane;
        }
        if (own && (lane & 7) == 0) my_words[lane >> 3] = w8;
        TRC(const long long t_p3 = clock64(); t_ph[2] += t_p3 - t_p2;)
        __syncwarp();
        if (elect_one()) {          // elect.sync, not lane == 0: ptxas then keeps the copy operands in uniform registers (no vote loops)
          TRC(const long long tw = clock64(); t_pre += tw - t_top;)
          mbar_wait(empty_bar + 8 * s, ph);
          TRC(const long long ti = clock64(); w_empty += ti - tw; ++n_items;)
          const uint32_t st = base + s * p.stage_bytes;
          int count = __popc(bal);
          uint32_t *mw = reinterpret_cast<uint32_t *>(s_masks + s * MAX_MT + m);
          for (int q = lane_lo >> 3; q < (lane_lo >> 3) + p.MT; ++q) mw[q] = my_words[q];
          if (p.dbg & 16) count = 0;
          const uint32_t bytes = (uint32_t)(count * 512) + ((part == 0 && !(p.dbg & 1)) ? (uint32_t)p.b_stage : 0u);
          if (bytes) mbar_expect_tx(full_bar + 8 * s, bytes);
          else mbar_arrive(full_bar + 8 * s);
          if (part == 0 && !(p.dbg & 1)) {
            int wtap = p.mirror ? p.V - 1 - trow : trow;
            if (p.item_off) {           // tile groups are tap-pure: the tap whose item range holds this group
              const int item = (tg * p.MT * TM) / p.rows_per_item;
              wtap = 0;
              for (int k = 1; k < p.n_taps; ++k) wtap += (__ldg(&p.item_off[k]) <= item) ? 1 : 0;
            }
            tma_tile_2d(st + p.MT * A_STAGE, &map_w, kc * p.kelems, wtap * p.c_out + n0, full_bar + 8 * s);
            TRC(t_b += clock64() - ti;)
          }
          TRC(n_copies += count;)
          const uint32_t a0 = st + m * A_STAGE;
          const int col = kc * p.kelems;
#pragma unroll 2
          for (int i = 0; i < count; ++i) {
            const int4 r = my_rows[i];
            tma_gather4(a0 + (uint32_t)my_dst[i] * 512u, &map_x, col, r.x, r.y, r.z, r.w, full_bar + 8 * s);
          }
          TRC(t_issue += clock64() - ti;)
        }
        __syncwarp();
        cur = nxt;
        nxt = far;
        tg = tg1; kc = kc1; trow = trow1;
        tg1 = tg2; kc1 = kc2; trow1 = trow2;
        s += p.nprod;
        if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
      }
      TRC(if (p.trace && lane == 0 && pw == 0) {
        unsigned long long *t = p.trace + 4 + 8 * part;
        atomicAdd(t + 0, (unsigned long long)(clock64() - t_begin));
        atomicAdd(t + 1, (unsigned long long)w_empty);
        atomicAdd(t + 2, (unsigned long long)t_issue);
        atomicAdd(t + 3, (unsigned long long)n_items);
        atomicAdd(t + 4, (unsigned long long)n_copies);
        atomicAdd(t + 5, (unsigned long long)t_pre);
        atomicAdd(t + 6, (unsigned long long)t_b);
        if (part == 0) { atomicAdd(p.trace + 44, (unsigned long long)t_ph[0]); atomicAdd(p.trace + 45, (unsigned long long)t_ph[1]); atomicAdd(p.trace + 46, (unsigned long long)t_ph[2]); }
      })
    }
  } else {
    // =========================== epilogue warps ===========================
    const int quarter = warp & 3;                     // TMEM lanes 32*quarter .. +31 belong to this warp
    const int half = (warp - 1 - p.nprod * p.ni) >> 2; // two warps per quarter: this one takes the 16-column chunks c0/16 % 2 == half
    const uint32_t tq = tmem + ((uint32_t)(quarter * 32) << 16);
    const float *aux = p.residual ? p.residual : p.bnb_x;   // the per-row fp32 operand of the epilogue (never both)
    for (int c = 32 * half; c < 2 * acc_cols; c += 32 * EPI_SPLIT) tmem_zero32(tq + c);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(acce_bar);
      mbar_arrive(acce_bar + 8);
    }
    float *my_stat = s_stat + (half * 4 + quarter) * 512;
    for (int c = lane; c < 512; c += 32) my_stat[c] = 0.f;
    __syncwarp();
    int gi = 0;
    TRC(long long e_wait = 0; const long long e_begin = clock64();)
    for (int tg = blockIdx.x; tg < p.n_groups; tg += G, ++gi) {
      const int buf = gi & 1;
      // rows of this group (lane = row), resolved BEFORE the wait so that the lines of the per-row operand of the epilogue
      // (residual or BatchNorm input: fp32 rows that nobody has touched recently) can be pulled into L2 while the
      // accumulators are still being computed: the dependent loads below then cost an L2 hit, not an HBM round trip
      int r_m[MAX_MT];
#pragma unroll
      for (int m = 0; m < MAX_MT; ++m) {
        int r = (tg * p.MT + m) * TM + quarter * 32 + lane;
        bool live = m < p.MT && r < p.n_rows;
        if (live && p.out_rows) {         // scattered result rows (each written exactly once)
          r = __ldg(&p.out_rows[r]);
          live = r >= 0 && r < p.out_limit;
        }
        r_m[m] = live ? r : -1;
        if (aux && live && half == 0) {
          const char *line = reinterpret_cast<const char *>(aux + (long long)r * p.c_out + n0);
          for (int b = 0; b < p.TN * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + b));
        }
      }
      TRC(const long long ew = clock64();)
      mbar_wait_warp(accf_bar + 8 * buf, (uint32_t)(gi >> 1) & 1u, lane);
      TRC(e_wait += clock64() - ew;)
      tc_fence_after();
      // Epilogue, 16 accumulator columns at a time.  tcgen05.ld hands every lane one ROW (16 consecutive columns); writing
      // rows from that layout makes each store instruction touch 32 different lines, and the load/store unit -- shared
      // with the producers' table loads and the MMA thread's mask loads -- was what this kernel queued on.  So the chunk
      // is transposed through a conflict-free 2 KB shared-memory tile: afterwards 4 consecutive lanes hold the 64
      // contiguous bytes of one row, a warp instruction covers 8 rows x 64 B, and the residual / BatchNorm-input reads
      // and the result stores are coalesced.  The column statistics fall out of this layout for free: a lane keeps the
      // same 4 columns for all its rows, so it sums them in registers and only the 8 lanes that share a column group
      // are folded with shuffles, once per chunk.
      float4 *my_t = s_tile + (half * 4 + quarter) * 128;    // [32 rows][4 chunks of 16 B], chunk index XOR-swizzled by row
      const int cj = lane & 3, rg = lane >> 2;                // after the transpose: this lane's column chunk and row slot
#pragma unroll 1
      for (int c0 = 16 * half; c0 < p.TN; c0 += 16 * EPI_SPLIT) {
        const int col = n0 + c0 + 4 * cj;                     // first of this lane's 4 columns
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sc = bias4, sh = bias4, wv = bias4, bv = bias4;
        if (p.bias) bias4 = __ldg(reinterpret_cast<const float4 *>(p.bias + col));
        if (p.ep_scale) {
          sc = __ldg(reinterpret_cast<const float4 *>(p.ep_scale + col));
          sh = __ldg(reinterpret_cast<const float4 *>(p.ep_shift + col));
        }
        if (p.bnb_x) {
          // w = invstd*gamma, b = beta - mean*w: the very expressions (and roundings) of the BatchNorm forward kernel
          const float4 mu = __ldg(reinterpret_cast<const float4 *>(p.bnb_mean + col));
          wv = __ldg(reinterpret_cast<const float4 *>(p.bnb_invstd + col));
          if (p.bnb_gamma) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(p.bnb_gamma + col));
            wv.x *= g.x; wv.y *= g.y; wv.z *= g.z; wv.w *= g.w;
          }
          if (p.bnb_beta) bv = __ldg(reinterpret_cast<const float4 *>(p.bnb_beta + col));
          bv.x = -mu.x * wv.x + bv.x; bv.y = -mu.y * wv.y + bv.y; bv.z = -mu.z * wv.z + bv.z; bv.w = -mu.w * wv.w + bv.w;
        }
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;   // column statistics of this lane's 4 columns
#pragma unroll
        for (int m = 0; m < MAX_MT; ++m) {
          if (m >= p.MT) break;
          // this lane's 4 rows of the tile and their per-row operand: all four loads are in flight before anything waits
          int r4[4];
          float4 a4[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            r4[it] = __shfl_sync(0xffffffffu, r_m[m], 8 * it + rg);
            a4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (aux && r4[it] >= 0) a4[it] = __ldg(reinterpret_cast<const float4 *>(aux + (long long)r4[it] * p.c_out + col));
          }
          float v[16];
          tmem_ld16(tq + buf * acc_cols + m * p.TN + c0, v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            my_t[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = 8 * it + rg;
            const int r = r4[it];
            float4 o = my_t[row * 4 + (cj ^ ((row >> 1) & 3))];
            if (r >= 0 && !(p.dbg & 8)) {
              const long long at = (long long)r * p.c_out + col;
              if (p.bias) { o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w; }
              if (p.residual) {
                const float4 rv = a4[it];
                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
              }
              if (p.ep_scale) {
                o.x = fmaf(sc.x, o.x, sh.x); o.y = fmaf(sc.y, o.y, sh.y); o.z = fmaf(sc.z, o.z, sh.z); o.w = fmaf(sc.w, o.w, sh.w);
                o.x = o.x > 0.f ? o.x : o.x * p.ep_leak; o.y = o.y > 0.f ? o.y : o.y * p.ep_leak;
                o.z = o.z > 0.f ? o.z : o.z * p.ep_leak; o.w = o.w > 0.f ? o.w : o.w * p.ep_leak;
              }
              float4 sq = make_float4(o.x * o.x, o.y * o.y, o.z * o.z, o.w * o.w);       // second statistic
              if (p.bnb_x) {
                // BatchNorm backward: mask recomputed from the BatchNorm's input exactly as its forward pass evaluated it
                const float4 xv = a4[it];
                o.x = fmaf(wv.x, xv.x, bv.x) > 0.f ? o.x : o.x * p.bnb_leak;
                o.y = fmaf(wv.y, xv.y, bv.y) > 0.f ? o.y : o.y * p.bnb_leak;
                o.z = fmaf(wv.z, xv.z, bv.z) > 0.f ? o.z : o.z * p.bnb_leak;
                o.w = fmaf(wv.w, xv.w, bv.w) > 0.f ? o.w : o.w * p.bnb_leak;
                sq = make_float4(o.x * xv.x, o.y * xv.y, o.z * xv.z, o.w * xv.w);
              }
              *reinterpret_cast<float4 *>(p.out + at) = o;
              if (p.out_bf16) {
                uint2 h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(o.y), "f"(o.x));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(o.w), "f"(o.z));
                *reinterpret_cast<uint2 *>(p.out_bf16 + at) = h;
              }
              s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
              s2.x += sq.x; s2.y += sq.y; s2.z += sq.z; s2.w += sq.w;
            }
          }
          __syncwarp();                  // the tile is rewritten by the next chunk
        }
        if (p.stats) {
          // fold the 8 lanes that hold the same 4 columns (lane bits 2..4), then lanes 0..3 add to the fp64 totals
#pragma unroll
          for (int sft = 4; sft <= 16; sft <<= 1) {
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, sft); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, sft);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, sft); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, sft);
            s2.x += __shfl_xor_sync(0xffffffffu, s2.x, sft); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, sft);
            s2.z += __shfl_xor_sync(0xffffffffu, s2.z, sft); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, sft);
          }
          // this warp's slice of the CTA's running totals (fp32: ~50 additions per column per CTA); one fp64 atomic per column
          // per CTA at the very end -- per-group global atomics on 2*C addresses from 148 CTAs were a serialisation point
          if (lane < 4) {
            float4 *t1 = reinterpret_cast<float4 *>(my_stat + c0 + 4 * cj), *t2 = reinterpret_cast<float4 *>(my_stat + 256 + c0 + 4 * cj);
            float4 a = *t1, b = *t2;
            a.x += s1.x; a.y += s1.y; a.z += s1.z; a.w += s1.w;
            b.x += s2.x; b.y += s2.y; b.z += s2.z; b.w += s2.w;
            *t1 = a; *t2 = b;
          }
        }
      }
      for (int c = 16 * half; c < acc_cols; c += 16 * EPI_SPLIT) tmem_zero16(tq + buf * acc_cols + c);     // exactly the columns this warp read
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar + 8 * buf);
    }
    if (p.stats) {
      __syncwarp();
      for (int c = lane; c < p.TN; c += 32) {
        if (((c >> 4) % EPI_SPLIT) != half) continue;          // the chunks this warp accumulated
        double *st = p.stats + (long long)(blockIdx.x * 4 + quarter) * p.stats_stride;
        atomicAdd(st + n0 + c, (double)my_stat[c]);
        atomicAdd(st + p.c_out + n0 + c, (double)my_stat[256 + c]);
      }
    }
    TRC(if (p.trace && lane == 0 && quarter == 0 && half == 0) {
      atomicAdd(p.trace + 41, (unsigned long long)(clock64() - e_begin));
      atomicAdd(p.trace + 42, (unsigned long long)e_wait);
    })
  }
  tc_fence_before();
  __syncthreads();
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
constexpr int KR = PAIR_ITEM;     // rules (GEMM K) per pipeline item (64)
constexpr int SUB = KR * 128;     // bytes of one [KR rules x 128 B] channel block (32 tf32 or 64 bf16 channels)
constexpr int WG_THREADS = 288;   // warp 0: TMEM + MMA; warps 1-4: producers; warps 5-8: epilogue (TMEM quarter = warp & 3)

struct WgParams {
  float *dw;
  long long part_stride;       // deterministic mode: CTAs of row range r accumulate into dw + r * part_stride (zeroed scratch)
  const int *gi, *si, *blk_item;
  int V, Cg, Cs, transpose_out, n_blk, blk_per_range;
  int N, m_tiles, stages, nprod, stage_bytes, tmem_cols;
  int bf16, cpb;                    // operand type; channels per 128-byte block (32 tf32 / 64 bf16)
};

__global__ void __launch_bounds__(WG_THREADS) k_wgrad_tma(const __grid_constant__ CUtensorMap map_g,
                                                          const __grid_constant__ CUtensorMap map_s, WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  int *s_pidx = reinterpret_cast<int *>(smem + p.stages * p.stage_bytes);     // [4 producers][2 * KR]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_pidx + 4 * 2 * KR);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t accum_bar = empty_bar + 8 * p.stages;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g_atoms = p.Cg / p.cpb, s_atoms = p.N / p.cpb;
  const int apm = 128 / p.cpb;                 // blocks per 128-channel accumulator tile
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
      const uint32_t idesc = idesc_make(p.N, 1, 1, p.bf16 != 0);
      int s = 0;
      uint32_t ph = 0;
      for (int item = ib; item < ie; ++item) {
        mbar_wait(full_bar + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = base + s * p.stage_bytes;
        if (p.bf16) {
          const uint64_t bd = desc_mn128(st + g_atoms * SUB, SUB);
          for (int mt = 0; mt < p.m_tiles; ++mt) {
            const uint64_t ad = desc_mn128(st + mt * apm * SUB, SUB);
#pragma unroll
            for (int kk = 0; kk < KR / 16; ++kk)   // K = 16 rules per MMA = two 1024-byte K groups of every block
              mma_bf16(tmem + mt * p.N, ad + (uint64_t)(kk * 128), bd + (uint64_t)(kk * 128), idesc,
                       (item > ib || kk) ? 1u : 0u);
          }
        } else {
          const uint64_t bd = desc_mn32(st + g_atoms * SUB, SUB);
          for (int mt = 0; mt < p.m_tiles; ++mt) {
            const uint64_t ad = desc_mn32(st + mt * apm * SUB, SUB);
#pragma unroll
            for (int kk = 0; kk < KR / 8; ++kk)    // K = 8 rules per MMA = two 512-byte K groups of every block
              mma_tf32(tmem + mt * p.N, ad + (uint64_t)(kk * 64), bd + (uint64_t)(kk * 64), idesc,
                       (item > ib || kk) ? 1u : 0u);
          }
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
      int *my = s_pidx + pw * 2 * KR;
      int it = ib + pw;
      int s = pw;                       // stage of item `it`
      uint32_t ph = 1;
      int g_next[2] = {0, 0}, s_next[2] = {0, 0};
      if (it < ie) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          g_next[h] = __ldg(&p.gi[(long long)it * KR + h * 32 + lane]);
          s_next[h] = __ldg(&p.si[(long long)it * KR + h * 32 + lane]);
        }
      }
      for (; it < ie; it += p.nprod) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          my[h * 32 + lane] = g_next[h];
          my[KR + h * 32 + lane] = s_next[h];
        }
        if (it + p.nprod < ie) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            g_next[h] = __ldg(&p.gi[(long long)(it + p.nprod) * KR + h * 32 + lane]);
            s_next[h] = __ldg(&p.si[(long long)(it + p.nprod) * KR + h * 32 + lane]);
          }
        }
        __syncwarp();
        if (elect_one()) {
          mbar_wait(empty_bar + 8 * s, ph);
          mbar_expect_tx(full_bar + 8 * s, (uint32_t)p.stage_bytes);
          const uint32_t dst = base + s * p.stage_bytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {            // 8 gather groups (32 rules) at a time
            int4 rg[8], rs[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              rg[g] = reinterpret_cast<const int4 *>(my)[h * 8 + g];
              rs[g] = reinterpret_cast<const int4 *>(my + KR)[h * 8 + g];
            }
            for (int a = 0; a < g_atoms; ++a) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                tma_gather4(dst + a * SUB + (h * 8 + g) * 512, &map_g, a * p.cpb, rg[g].x, rg[g].y, rg[g].z, rg[g].w,
                            full_bar + 8 * s);
            }
            for (int a = 0; a < s_atoms; ++a) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                tma_gather4(dst + (g_atoms + a) * SUB + (h * 8 + g) * 512, &map_s, n0 + a * p.cpb, rs[g].x, rs[g].y, rs[g].z,
                            rs[g].w, full_bar + 8 * s);
            }
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
    mbar_wait_warp(accum_bar, 0, lane);
    tc_fence_after();
    const uint32_t acc = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int mt = 0; mt < p.m_tiles; ++mt) {
      const int cg = mt * 128 + quarter * 32 + lane;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(acc + mt * p.N + c0, v);
        if (cg < p.Cg) {
          if (!p.transpose_out) {
            float *dst = p.dw + blockIdx.y * p.part_stride + ((long long)tap * p.Cg + cg) * p.Cs + n0 + c0;
#pragma unroll
            for (int q = 0; q < 32; q += 4) red_add_v4(dst + q, v[q], v[q + 1], v[q + 2], v[q + 3]);
          } else {
            float *dst = p.dw + blockIdx.y * p.part_stride + ((long long)tap * p.Cs + n0 + c0) * p.Cg + cg;
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
  const int kel = a.bf16 ? 64 : 32;
  return !a.scatter && a.c_in >= kel && a.c_in % kel == 0 && a.c_out >= 32 && a.c_out % 32 == 0 && a.V <= 32 &&
         tma::pick_tn(a.c_out) > 0 && a.in_rows > 0 && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}

void conv_tma(const ConvArgs &a, cudaStream_t s) {
  using namespace tma;
  SCN_CHECK(a.weight_nk != nullptr, "conv_tma needs the [V][Cout][Cin] weight layout");
  if (a.n_rows == 0) return;
  ConvParams p;
  p.out_rows = a.out_rows; p.item_off = a.item_off; p.rows_per_item = a.rows_per_item; p.n_taps = a.n_taps;
  p.out_limit = a.out_limit;
  p.tile_mask = a.tile_mask;
  p.residual = a.residual;
  p.ep_scale = a.ep_scale; p.ep_shift = a.ep_shift; p.ep_leak = a.ep_leak; p.out_bf16 = a.out_bf16;
  p.bnb_x = a.bnb_x; p.bnb_mean = a.bnb_mean; p.bnb_invstd = a.bnb_invstd; p.bnb_gamma = a.bnb_gamma; p.bnb_beta = a.bnb_beta;
  p.bnb_leak = a.bnb_leak;
  SCN_CHECK(!a.bnb_x || a.stats, "conv_tma: the fused BatchNorm backward needs the statistics buffer");
  SCN_CHECK(!(a.bnb_x && a.residual), "conv_tma: residual and fused BatchNorm backward are mutually exclusive");
  p.stats = a.stats;
  p.stats_stride = 0;
  if (a.stats) SCN_CUDA(cudaMemsetAsync(a.stats, 0, sizeof(double) * 2 * (size_t)a.c_out, s));
  p.in = a.in; p.bias = a.bias; p.out = a.out; p.tbl = a.tbl; p.tbl_stride = a.tbl_stride; p.n_rows = a.n_rows;
  p.V = a.V; p.c_in = a.c_in; p.c_out = a.c_out; p.mirror = a.mirror ? 1 : 0;
  p.bf16 = a.bf16 ? 1 : 0;
  p.kelems = a.bf16 ? 64 : 32;
  p.TN = pick_tn(a.c_out);
  p.b_stage = p.TN * 128;
  // row tiles per group: as many as two TMEM accumulator buffers allow (the weight tile of an item is shared by them)
  static const int force_mt = env_int("SCN_CONV_MT", 0);
  int mt = 512 / (2 * p.TN);
  if (mt > 2) mt = 2;
  if (force_mt > 0) mt = force_mt;
  if (mt > MAX_MT) mt = MAX_MT;
  while (mt > 1 && 2 * mt * p.TN > 512) --mt;
  if (mt < 1) mt = 1;
  const int tiles = (a.n_rows + TM - 1) / TM;
  if (tiles < mt * sm_count()) mt = 1;                 // small levels: more, smaller groups keep every SM busy
  p.MT = mt;
  p.stage_bytes = p.MT * A_STAGE + p.b_stage;
  // alignment slack + masks + producers' gather groups (4 owner slots x MT x 32 int4) + barriers + transpose tiles of the epilogue warps
  // alignment slack + masks + producers' gather lists (4 owner slots x MT x 32 int4) + their words + barriers + transpose tiles and
  // statistics of the epilogue warps
  const int fixed = 1024 + 8 * MAX_MT * 16 + 4 * p.MT * 32 * 16 + MAX_PROD_WARPS * 32 + 8 * (2 * 8 + 4) + 64 + EPI_WARPS * 2048 + EPI_WARPS * 2048;
  static const int smem_kb = env_int("SCN_CONV_SMEM_KB", 227);
  int st = (smem_kb * 1024 - fixed) / p.stage_bytes;
  if (st > 8) st = 8;
  SCN_CHECK(st >= 2, "conv_tma: shared memory budget");
  p.stages = st;
  p.nprod = st < 4 ? st : 4;                         // owner slots (<= stages: a parity wait can never be two phases off)
  p.ni = 4;                          // shares per item: each owns MT consecutive 32-row chunks of one row tile
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.MT * p.TN) p.tmem_cols <<= 1;
  p.n_groups = (tiles + p.MT - 1) / p.MT;
  p.n_tiles = tiles;
  static const int dbg = env_int("SCN_CONV_DBG", 0);
  p.dbg = dbg;
  static const int prefetch = env_int("SCN_CONV_PREFETCH", 0);
  p.prefetch = prefetch;
  const size_t smem = (size_t)fixed + (size_t)p.stages * p.stage_bytes;
  CUtensorMap mx = make_map(a.in, a.bf16, (uint64_t)a.c_in, (uint64_t)a.in_rows, p.kelems, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap mw = make_map(a.weight_nk, a.bf16, (uint64_t)a.c_in, (uint64_t)(a.n_taps ? a.n_taps : a.V) * a.c_out, p.kelems, (uint32_t)p.TN,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  static SmemAttrCache smem_attr;       // per device: the attribute is device state
  smem_attr.ensure(k_conv_tma, smem);
  const int n_tiles_n = a.c_out / p.TN;
  int gx = sm_count() / n_tiles_n;
  if (gx < 1) gx = 1;
  if (gx > p.n_groups) gx = p.n_groups;
  dim3 grid(gx, n_tiles_n);
  DevBuf<double> stats_part;
  if (a.stats && deterministic()) {      // one zeroed [2][c_out] slice per (CTA, epilogue warp), summed in that order after the launch
    static_assert(EPI_WARPS == 4, "deterministic statistics: one slice per TMEM quarter");
    stats_part.alloc((size_t)gx * 4 * 2 * a.c_out, s);
    SCN_CUDA(cudaMemsetAsync(stats_part.p, 0, sizeof(double) * stats_part.n, s));
    p.stats = stats_part.p;
    p.stats_stride = 2ll * a.c_out;
  }
  static const int trace = env_int("SCN_TRACE", 0);
  p.trace = nullptr;
#ifndef SCN_TRACE_BUILD
  if (trace) fprintf(stderr, "[conv trace] rebuild with -DSCN_TRACE_BUILD to collect the clock64 breakdown\n");
#else
  if (trace) {
    SCN_CUDA(cudaMalloc((void **)&p.trace, 8 * 64));
    SCN_CUDA(cudaMemset(p.trace, 0, 8 * 64));
  }
#endif
  k_conv_tma<<<grid, 32 + 32 * p.nprod * p.ni + 32 * EPI_WARPS, smem, s>>>(mx, mw, p);
  SCN_LAUNCH_CHECK();
  if (stats_part.p) sum_partials(stats_part.p, gx * 4, 2ll * a.c_out, a.stats, s);
  stats_part.release(s);
  if (p.trace) {
    unsigned long long h[64];
    SCN_CUDA(cudaStreamSynchronize(s));
    SCN_CUDA(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    const double n = (double)grid.x * grid.y;
    fprintf(stderr, "[conv trace] rows %d C %d->%d V %d TN %d MT %d stages %d nprod %d ni %d ctas %.0f | per CTA: mma loop %.0f clk, "
            "wait-full %.0f, wait-acc %.0f, in-MMA-issue %.0f, tiles with MMAs %.0f of %.0f | epilogue loop %.0f wait %.0f |", a.n_rows, a.c_in, a.c_out, a.V, p.TN, p.MT, p.stages, p.nprod, p.ni, n, h[0] / n,
            h[1] / n, h[2] / n, h[40] / n, h[3] / n, (double)p.n_groups * p.MT * a.V * (a.c_in / p.kelems) / n, h[41] / n, h[42] / n);
    fprintf(stderr, " commit %.0f | share0 pre phases: table-load %.0f prefetch %.0f ballots+smem %.0f |", h[43] / n, h[44] / n, h[45] / n, h[46] / n);
    for (int q = 0; q < p.ni; ++q) {
      const unsigned long long *t = h + 4 + 8 * q;
      fprintf(stderr, " share%d: loop %.0f pre %.0f wait-empty %.0f issue %.0f (masks+B %.0f) items %.0f copies %.0f |", q, t[0] / n,
              t[5] / n, t[1] / n, t[2] / n, t[6] / n, t[3] / n, t[4] / n);
    }
    fprintf(stderr, "\n");
  }
}

bool wgrad_tma_supported(const WgradArgs &a) {
  const int cg = a.table_on_a ? a.c_a : a.c_b, cs = a.table_on_a ? a.c_b : a.c_a;
  const int cpb = a.bf16 ? 64 : 32;
  // two pipeline stages of (all G blocks + one S block) must fit shared memory: tf32 with more than ~416 gathered channels
  // does not (e.g. the 640 -> 320 block of UNet m=64), and takes the exact-fp32 kernel instead
  const long long min_stage = (long long)(cg / cpb + 1) * tma::SUB;
  const long long tail = ((cg + 127) / 128) * 128 > cg ? (128 / cpb) * tma::SUB : 0;
  if (2 * min_stage + tail + 4096 > 225 * 1024) return false;
  return a.gi != nullptr && a.g_rows > 0 && a.s_rows > 0 && cg >= cpb && cg % cpb == 0 && cs >= cpb && cs % cpb == 0 &&
         a.V <= 32 && ((uintptr_t)a.a % 16 == 0) && ((uintptr_t)a.b % 16 == 0);
}

void wgrad_tma(const WgradArgs &a, cudaStream_t s) {
  using namespace tma;
  SCN_CUDA(cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.V * a.c_a * a.c_b, s));
  if (a.n_rows == 0 || a.n_blk == 0) return;
  WgParams p;
  const float *G = a.table_on_a ? a.a : a.b;
  const float *S = a.table_on_a ? a.b : a.a;
  p.bf16 = a.bf16 ? 1 : 0;
  p.cpb = a.bf16 ? 64 : 32;
  p.Cg = a.table_on_a ? a.c_a : a.c_b;
  p.Cs = a.table_on_a ? a.c_b : a.c_a;
  p.transpose_out = a.table_on_a ? 0 : 1;
  p.dw = a.dw; p.gi = a.gi; p.si = a.si; p.blk_item = a.blk_item; p.V = a.V; p.n_blk = a.n_blk;
  p.m_tiles = (p.Cg + 127) / 128;
  // every stage holds the item's G blocks, then its S blocks; an accumulator tile always spans 128 channels of G, so the
  // last tile of a Cg that is not a multiple of 128 reads on into the S blocks (finite data; rows >= Cg are ignored)
  const int fixed = 1024 + 4 * 2 * KR * 4 + 8 * (2 * 8 + 1) + 64;
  const int tail = p.m_tiles * 128 > p.Cg ? (128 / p.cpb) * SUB : 0;
  // column tile: the widest multiple of a block that divides Cs, fits TMEM and leaves at least 3 pipeline stages
  p.N = 0;
  int budget = 225 * 1024, st = 0;
  for (int pass = 0; pass < 2 && p.N == 0; ++pass)
    for (int n = 256; n >= p.cpb; n -= p.cpb) {
      if (p.Cs % n != 0 || p.m_tiles * n > 512) continue;
      const int sb = (p.Cg / p.cpb + n / p.cpb) * SUB;
      if ((225 * 1024 - fixed - tail) / sb < (pass == 0 ? 3 : 2)) continue;
      p.N = n;
      break;
    }
  SCN_CHECK(p.N > 0, "wgrad_tma: no column tile fits shared memory");
  p.tmem_cols = 32;
  while (p.tmem_cols < p.m_tiles * p.N) p.tmem_cols <<= 1;
  p.stage_bytes = (p.Cg / p.cpb + p.N / p.cpb) * SUB;
  budget = 113 * 1024;                               // two CTAs per SM when TMEM and >= 3 stages allow
  if (512 / p.tmem_cols < 2 || (budget - fixed - tail) / p.stage_bytes < 3) budget = 225 * 1024;
  st = (budget - fixed - tail) / p.stage_bytes;
  SCN_CHECK(st >= 2, "wgrad_tma: shared memory budget");
  if (st > 8) st = 8;
  p.stages = st;
  p.nprod = st < 4 ? st : 4;
  const size_t smem = (size_t)fixed + (size_t)p.stages * p.stage_bytes + tail;
  // row range per CTA: about 8 MB of G + S rows, so the ranges in flight (SMs x CTAs/SM / V of them) fit in L2
  static const int range_kb = env_int("SCN_WG_RANGE_KB", 8192);
  long long rows = (long long)range_kb * 1024 / ((long long)(p.Cg + p.Cs) * 4);
  int bpr = (int)(rows / a.blk_rows);
  if (bpr < 1) bpr = 1;
  p.blk_per_range = bpr;
  const int ranges = (a.n_blk + bpr - 1) / bpr;
  const CUtensorMapSwizzle sw = a.bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUtensorMap mg = make_map(G, a.bf16, (uint64_t)p.Cg, (uint64_t)a.g_rows, p.cpb, 1, sw);
  CUtensorMap ms = make_map(S, a.bf16, (uint64_t)p.Cs, (uint64_t)a.s_rows, p.cpb, 1, sw);
  static SmemAttrCache smem_attr;
  smem_attr.ensure(k_wgrad_tma, smem);
  dim3 grid(a.V, ranges, p.Cs / p.N);
  DevBuf<float> part;
  const size_t n_dw = (size_t)a.V * a.c_a * a.c_b;
  p.part_stride = 0;
  if (deterministic() && ranges > 1) {     // one zeroed slice per row range: a single addend per address, summed in range order below
    part.alloc(n_dw * ranges, s);
    SCN_CUDA(cudaMemsetAsync(part.p, 0, sizeof(float) * n_dw * ranges, s));
    p.dw = part.p;
    p.part_stride = (long long)n_dw;
  }
  k_wgrad_tma<<<grid, WG_THREADS, smem, s>>>(mg, ms, p);
  SCN_LAUNCH_CHECK();
  if (part.p) sum_partials(part.p, ranges, (long long)n_dw, a.dw, s);
  part.release(s);
}

}  // namespace scn

// Micro-benchmarks behind the pipeline design in DESIGN.md (run on the GPU box):
//   1. mbarrier ping-pong cost: producer thread <-> consumer thread, release by plain arrive vs tcgen05.commit
//   2. TMA gather4 throughput per SM as a function of bytes in flight and of the row pattern
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_pipe tools/ubench_pipe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"((uint64_t)map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------- 1. ping-pong
// mode 0: consumer releases with mbarrier.arrive; mode 1: with tcgen05.commit (no MMAs outstanding)
__global__ void k_pingpong(int items, int stages, int mode, long long *out) {
  __shared__ uint64_t bars[32];
  const uint32_t full = smem_u32(bars), empty = full + 8 * stages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    if (elect_one())
      for (int it = 0; it < items; ++it) {
        const int s = it % stages;
        mbar_wait(empty + 8 * s, ((it / stages) & 1) ^ 1);
        mbar_expect_tx(full + 8 * s, 0);
      }
  } else if (warp == 1) {
    if (elect_one())
      for (int it = 0; it < items; ++it) {
        const int s = it % stages;
        mbar_wait(full + 8 * s, (it / stages) & 1);
        if (mode == 0) mbar_arrive(empty + 8 * s);
        else mma_commit(empty + 8 * s);
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------- 2. gather4
// One producer thread per warp (nprod warps) issues gather4 copies of 128-byte rows into `stages` stages of
// `stage_rows` rows; a consumer thread waits for each stage and releases it immediately.
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__global__ void k_gather(const __grid_constant__ CUtensorMap map, const int *rows, int n_rows_list, int items, int stages,
                         int stage_rows, int nprod, long long *out, int tile_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  int *s_rows = reinterpret_cast<int *>(smem_raw + (base - raw) + stages * stage_rows * 128);
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_rows + 4096);
  const uint32_t full = smem_u32(bars), empty = full + 8 * stages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_rows[i] = rows[((long long)blockIdx.x * 4096 + i) % n_rows_list];
  __syncthreads();
  long long t0 = clock64();
  const int per_warp = stage_rows / nprod;          // rows per producer warp per stage
  if (warp == 0) {
    if (elect_one())
      for (int it = 0; it < items; ++it) {
        const int s = it % stages;
        mbar_wait(empty + 8 * s, ((it / stages) & 1) ^ 1);
        mbar_expect_tx(full + 8 * s, stage_rows * 128);
      }
  } else if (warp == 1) {
    if (elect_one())
      for (int it = 0; it < items; ++it) {
        const int s = it % stages;
        mbar_wait(full + 8 * s, (it / stages) & 1);
        mbar_arrive(empty + 8 * s);
      }
  } else if (warp - 2 < nprod) {
    const int pw = warp - 2;
    if (elect_one()) {
      long long pos = ((long long)blockIdx.x * items) * stage_rows;
      for (int it = 0; it < items; ++it) {
        const int s = it % stages;
        mbar_wait(empty + 8 * s, ((it / stages) & 1) ^ 1);
        const uint32_t dst = base + s * stage_rows * 128 + pw * per_warp * 128;
        const int4 *r4 = reinterpret_cast<const int4 *>(s_rows + ((it * stage_rows + pw * per_warp) & 4095));
        if (per_warp == 32) {           // indices first (as the real kernels do), then back-to-back copies
          int4 r[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) r[g] = r4[g];
          if (tile_mode) {
            tma_tile_2d(dst, &map, 0, r[0].x, full + 8 * s);
          } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) tma_gather4(dst + g * 512, &map, 0, r[g].x, r[g].y, r[g].z, r[g].w, full + 8 * s);
          }
        } else {
          int4 r[2];
          r[0] = r4[0]; r[1] = r4[1];
          tma_gather4(dst, &map, 0, r[0].x, r[0].y, r[0].z, r[0].w, full + 8 * s);
          tma_gather4(dst + 512, &map, 0, r[1].x, r[1].y, r[1].z, r[1].w, full + 8 * s);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------- 3. item-interleaved producers
// producer warp p owns items p, p+nprod, ...: one thread waits for the stage, posts the byte count and issues ALL the
// item's gather4 copies (stage_rows/4 of them); a consumer thread releases stages.  skip_pct: percentage of 4-row groups not issued.
__global__ void k_gather_il(const __grid_constant__ CUtensorMap map, const int *rows, int n_rows_list, int items, int stages,
                            int stage_rows, int nprod, long long *out, int skip_pct) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  int *s_rows = reinterpret_cast<int *>(smem_raw + (base - raw) + stages * stage_rows * 128);
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_rows + 4096);
  const uint32_t full = smem_u32(bars), empty = full + 8 * stages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_rows[i] = rows[((long long)blockIdx.x * 4096 + i) % n_rows_list];
  __syncthreads();
  long long t0 = clock64();
  const int ngroups = stage_rows / 4;
  const int nissue = ngroups - ngroups * skip_pct / 100;
  if (warp == 0) {
    if (elect_one())
      for (int it = 0; it < items; ++it) {
        const int s = it % stages;
        mbar_wait(full + 8 * s, (it / stages) & 1);
        mbar_arrive(empty + 8 * s);
      }
  } else if (warp - 1 < nprod) {
    const int pw = warp - 1;
    if (elect_one()) {
      for (int it = pw; it < items; it += nprod) {
        const int s = it % stages;
        mbar_wait(empty + 8 * s, ((it / stages) & 1) ^ 1);
        mbar_expect_tx(full + 8 * s, nissue * 512);
        const uint32_t dst = base + s * stage_rows * 128;
        const int4 *r4 = reinterpret_cast<const int4 *>(s_rows + ((it * stage_rows) & 4095));
#pragma unroll 8
        for (int g = 0; g < nissue; ++g) {
          int4 r = r4[g];
          tma_gather4(dst + g * 512, &map, 0, r.x, r.y, r.z, r.w, full + 8 * s);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------- 4. tcgen05.mma issue / execution rate
// one thread issues `n` MMAs (M=128, N=n_cols, one K step of 32 bytes per row) from shared memory, commit every 4, waits at the end
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
template <int KIND>   // 0 tf32, 1 bf16
__global__ void k_mma_rate(int n, int n_cols, int masked, int n_cta_stages, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw + (base - raw))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x == 0) {
    // idesc: c fp32 (1<<4); a/b format: tf32 = 2, bf16 = 1 at bits 7 and 10; N>>3 at 17; M>>4 at 24
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t mk = masked ? 0x0F0F0F0Fu : 0u;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint32_t st = base + (i % n_cta_stages) * 0;     // same stage: execution rate, not data movement
      const uint64_t ad = desc_k128(st) + (uint64_t)((i & 3) * 2);
      const uint64_t bd = desc_k128(st + 16384) + (uint64_t)((i & 3) * 2);
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u), "r"(mk) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u), "r"(mk) : "memory");
      if ((i & 3) == 3) mma_commit(smem_u32(&bar) + 0 * 8);
    }
    t1 = clock64();
  }
  __syncthreads();
  // drain: wait until the tensor pipe is idle (commit phases completed = n/4; just spin on time via a final commit)
  if (threadIdx.x == 0) {
    // parity of the (n/4)-th completion
    mbar_wait(smem_u32(&bar), ((n / 4) - 1) & 1);
    t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  long long *d_out; CK(cudaMalloc(&d_out, 148 * 8));
  std::vector<long long> h(148);
  printf("== ping-pong (1 CTA/SM, 148 CTAs): clk per item\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int stages : {1, 2, 4, 8}) {
      const int items = 4000;
      k_pingpong<<<148, 64>>>(items, stages, mode, d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost));
      printf("mode %s stages %d: %.1f clk/item\n", mode ? "tcgen05.commit" : "arrive", stages, (double)h[0] / items);
    }

  {
    long long *d2; CK(cudaMalloc(&d2, 148 * 16));
    std::vector<long long> h2(296);
    CK(cudaFuncSetAttribute(k_mma_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_mma_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    printf("== tcgen05.mma SS, M=128, K step = 32 B/row, 148 CTAs: clk per MMA (issue loop / until complete)\n");
    for (int kind = 0; kind < 2; ++kind)
      for (int ncols : {64, 128, 256})
        for (int masked = 0; masked < 2; ++masked) {
          const int n = 4096;
          if (kind == 0) k_mma_rate<0><<<148, 128, 50 * 1024>>>(n, ncols, masked, 1, d2);
          else k_mma_rate<1><<<148, 128, 50 * 1024>>>(n, ncols, masked, 1, d2);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h2.data(), d2, 148 * 16, cudaMemcpyDeviceToHost));
          printf("  %s N=%3d %s: issue %.1f clk/MMA, complete %.1f clk/MMA\n", kind ? "bf16" : "tf32", ncols,
                 masked ? "masked  " : "unmasked", (double)h2[0] / n, (double)h2[1] / n);
        }
  }
  if (getenv("UBENCH_MMA_ONLY")) return 0;

  // ---- gather4
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  const int C = 64; const long long NR = 2000000;
  float *x; CK(cudaMalloc(&x, NR * C * 4)); CK(cudaMemset(x, 0, NR * C * 4));
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)NR}; cuuint64_t strides[1] = {C * 4};
  cuuint32_t box[2] = {32, 1}; cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int LIST = 1 << 22;
  std::vector<int> rows(LIST);
  int *d_rows; CK(cudaMalloc(&d_rows, LIST * 4));
  CK(cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const char *pat_name[4] = {"sequential", "local-random(+-4096)", "all out-of-bounds", "60% OOB + local"};
  for (int pat = 0; pat < 4; ++pat) {
    srand(1);
    for (int i = 0; i < LIST; ++i) {
      int base_row = (int)(((long long)i * 7) % NR);
      if (pat == 0) rows[i] = i % NR;
      else if (pat == 1) rows[i] = (int)((base_row + rand() % 8192) % NR);
      else if (pat == 2) rows[i] = (int)NR;
      else rows[i] = (rand() % 100 < 60) ? (int)NR : (int)((base_row + rand() % 8192) % NR);
    }
    CK(cudaMemcpy(d_rows, rows.data(), LIST * 4, cudaMemcpyHostToDevice));
    printf("== gather4 pattern %s: bytes/clk/SM (requested), GB/s chip at measured clk\n", pat_name[pat]);
    for (int nprod : {1, 4})
      for (int stage_rows : {32, 128}) {
        if (stage_rows / nprod != 32 && stage_rows / nprod != 8) continue;
        for (int stages : {2, 4, 8}) {
          const int items = 2000;
          size_t smem = 1024 + (size_t)stages * stage_rows * 128 + 256 + 16384;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          k_gather<<<148, 64 + 32 * nprod, smem>>>(map, d_rows, LIST - 4096, items, stages, stage_rows, nprod, d_out, 0);
          cudaEventRecord(e0);
          k_gather<<<148, 64 + 32 * nprod, smem>>>(map, d_rows, LIST - 4096, items, stages, stage_rows, nprod, d_out, 0);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          CK(cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost));
          double bytes = (double)items * stage_rows * 128;
          printf("  nprod %d stage_rows %3d stages %d: %.1f B/clk/SM, %.0f clk/item, chip %.0f GB/s\n", nprod, stage_rows, stages,
                 bytes / h[0], (double)h[0] / items, bytes * 148 / (ms * 1e-3) / 1e9);
        }
      }
  }
  for (int pat : {1, 3}) {
    srand(1);
    for (int i = 0; i < LIST; ++i) {
      int base_row = (int)(((long long)i * 7) % NR);
      rows[i] = (int)((base_row + rand() % 8192) % NR);
    }
    CK(cudaMemcpy(d_rows, rows.data(), LIST * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_gather_il, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int skip = pat == 1 ? 0 : 50;
    printf("== item-interleaved producers, local-random rows, %d%% of 4-row groups skipped: issued bytes/clk/SM\n", skip);
    for (int nprod : {1, 2, 4, 8})
      for (int stages : {4, 8}) {
        if (nprod > stages) continue;      // parity waits alias when a producer can run two phases ahead
        const int items = 4000, stage_rows = 128;
        size_t smem = 1024 + (size_t)stages * stage_rows * 128 + 256 + 16384;
        k_gather_il<<<148, 32 + 32 * nprod, smem>>>(map, d_rows, LIST - 4096, items, stages, stage_rows, nprod, d_out, skip);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_gather_il<<<148, 32 + 32 * nprod, smem>>>(map, d_rows, LIST - 4096, items, stages, stage_rows, nprod, d_out, skip);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost));
        double bytes = (double)items * (stage_rows / 4 - stage_rows / 4 * skip / 100) * 512;
        printf("  nprod %d stages %d: %.1f B/clk/SM, %.0f clk/item, chip %.0f GB/s\n", nprod, stages, bytes / h[0],
               (double)h[0] / items, bytes * 148 / (ms * 1e-3) / 1e9);
      }
  }
  {
    // reference: plain 2-D tile loads of 32 consecutive rows x 128 B per producer (box {32,32})
    CUtensorMap map2;
    cuuint32_t box2[2] = {32, 32};
    ((EncodeFn)fn)(&map2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int i = 0; i < LIST; ++i) rows[i] = (i % (int)(NR - 64));
    CK(cudaMemcpy(d_rows, rows.data(), LIST * 4, cudaMemcpyHostToDevice));
    printf("== plain 2-D tile loads (32 rows x 128 B per copy)\n");
    for (int nprod : {1, 4})
      for (int stages : {2, 4, 8}) {
        const int items = 2000, stage_rows = 32 * nprod;
        size_t smem = 1024 + (size_t)stages * stage_rows * 128 + 256 + 16384;
        k_gather<<<148, 64 + 32 * nprod, smem>>>(map2, d_rows, LIST - 4096, items, stages, stage_rows, nprod, d_out, 1);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_gather<<<148, 64 + 32 * nprod, smem>>>(map2, d_rows, LIST - 4096, items, stages, stage_rows, nprod, d_out, 1);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost));
        double bytes = (double)items * stage_rows * 128;
        printf("  nprod %d stages %d: %.1f B/clk/SM, %.0f clk/item, chip %.0f GB/s\n", nprod, stages, bytes / h[0],
               (double)h[0] / items, bytes * 148 / (ms * 1e-3) / 1e9);
      }
  }
  return 0;
}

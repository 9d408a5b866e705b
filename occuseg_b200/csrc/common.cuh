// Shared declarations for libscn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <atomic>
#include <stdexcept>

namespace scn {

// ---- error plumbing: exceptions inside, int status + scn_last_error() at the C boundary ----------
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
void set_last_error(const std::string &s);

#define SCN_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::scn::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +         \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")");                          \
  } while (0)

#define SCN_CHECK(cond, msg)                                                                        \
  do {                                                                                              \
    if (!(cond)) throw ::scn::Error(std::string(msg) + " [" #cond "]");                             \
  } while (0)

extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
#define SCN_LAUNCH_CHECK()                                                                          \
  do {                                                                                              \
    ::scn::count_launch();                                                                          \
    SCN_CUDA(cudaGetLastError());                                                                   \
  } while (0)

int sm_count();                  // SM count of the CURRENT device (cached per device)
int current_device();
constexpr int MAX_DEVICES = 64;
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: call sites keep one of these per kernel
struct SmemAttrCache {
  std::atomic<size_t> configured[MAX_DEVICES];
  SmemAttrCache() { for (auto &c : configured) c.store(0); }
  template <typename K> void ensure(K kernel, size_t smem) {
    const int dev = current_device();
    if (dev >= 0 && dev < MAX_DEVICES && configured[dev].load(std::memory_order_acquire) >= smem) return;
    SCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < MAX_DEVICES) {
      size_t cur = configured[dev].load();
      while (cur < smem && !configured[dev].compare_exchange_weak(cur, smem)) {}
    }
  }
};

// ---- optional per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline) ----
enum ProfKind { PK_RULEBOOK = 0, PK_CONV_TC, PK_CONV_FP32, PK_WGRAD_TC, PK_WGRAD_FP32, PK_BN, PK_IO, PK_CAST, PK_COUNT };
struct ProfScope {   // no-op unless scn_profile(1) was called
  int slot = -1;
  cudaStream_t s;
  ProfScope(ProfKind kind, double bytes, double flops, cudaStream_t stream);
  ~ProfScope();
};
void prof_enable(bool on);
int prof_read(double *out, int max_kinds);   // out[kind*4 + {launches, ms, bytes, flops}]
const char *prof_name(int kind);

// ---- stream-ordered device buffers -------------------------------------------------------------------
// Buffers that are not released explicitly (error paths, the members of a handle being destroyed) are freed on the
// stream the library last worked on in this thread, so the free is ordered after the kernels that used them even
// when the caller runs on a non-default stream.
extern thread_local cudaStream_t t_last_stream;
inline cudaStream_t note_stream(void *stream) { return t_last_stream = (cudaStream_t)stream; }

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  void alloc(size_t count, cudaStream_t s) {
    release(s);
    n = count;
    if (count) SCN_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), s));
  }
  void release(cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { if (p) cudaFreeAsync(p, t_last_stream); }
};

// ---- voxel key: 16 bits per field, (batch, z, y, x) most->least significant -------------------------
// Sorting by this key reproduces the reference's row order: per sample, rank of the 31-bit key
// z<<21|y<<10|x (CUDA/CUDPPWrapper.cu:80-81) with samples concatenated (IOLayersRules.h:164-169),
// but without the reference's 10/11/10-bit aliasing.
__host__ __device__ inline uint64_t make_key(uint32_t b, uint32_t z, uint32_t y, uint32_t x) {
  return ((uint64_t)b << 48) | ((uint64_t)z << 32) | ((uint64_t)y << 16) | (uint64_t)x;
}
constexpr int COORD_LIMIT = 65535;          // exclusive upper bound on coordinates and batch index
constexpr uint64_t EMPTY_KEY = ~0ull;       // never a valid key (batch 65535 is rejected)

// ---- per-tap compacted rule lists of a [V][stride] table: the reference's rulebook (27 lists of (in,out) pairs
// ordered by out, SubmanifoldRules_cuda.cpp:167-187) in device memory, each list padded to a multiple of PAIR_ITEM
// rules with an out-of-range index.  gi = table entry (gathered row), si = table column (stationary row).
constexpr int PAIR_PAD = 0x7F7F7F7F;
constexpr int PAIR_ITEM = 64;   // rules per pipeline item of the weight-gradient kernel
struct PairList {
  DevBuf<int> gi, si;        // [PAIR_ITEM * n_items_ub]
  DevBuf<int> item_off;      // [V+1] first item (PAIR_ITEM rules) of every tap; item_off[V] = number of items
  DevBuf<int> blk_item;      // [V][n_blk+1] item holding tap k's first rule whose column is >= b*BLK_ROWS
  int n_blk = 0;
  int unit = PAIR_ITEM;      // rules per item (64 for the weight gradient, 256 = one tile group for the "up" products)
  long long n_items_ub = 0;  // host-side upper bound of item_off[V]
};
constexpr int BLK_ROWS = 512;
void build_pair_list(PairList &out, const int *tbl, int V, int stride, long long n_rules, int unit, int pad_byte,
                     cudaStream_t s);

// ---- one scale of one batch -------------------------------------------------------------------------
struct Level {
  ~Level();
  // Dilated 3x3x3 neighbourhoods of this scale (SubmanifoldConvolution(dilated_rate = d): taps at offsets d*(dx,dy,dz),
  // Metadata/SubmanifoldConvolutionRules.h:39-75,114-153): each is a Level of its own that shares this scale's rows and
  // hash (borrowed through `base`) and owns its table, rule lists and tile order.
  std::vector<Level *> dilated;
  Level *base = nullptr;
  int dilation = 1;
  int64_t size[3] = {0, 0, 0};
  int n = 0;                 // active rows
  int n_pad = 0;             // row stride of the [V][n_pad] tables (multiple of 128)
  DevBuf<uint64_t> keys;     // [n] sorted unique voxel keys; row id == index
  // open-addressing hash: key -> row
  DevBuf<ulonglong2> htab;   // {key, row}: one 16-byte entry = one 32-byte sector per probe (keys and rows in separate arrays cost two)
  uint32_t hmask = 0;
  // submanifold 3x3x3 neighbour table (output-stationary form of the reference's 27 rule lists)
  DevBuf<int> nbr;           // [27][n_pad], -1 = absent
  long long n_rules = -1;    // sum_k n_k, centre offset included
  // Tile order of the tensor-core convolutions: rows of every SORT_BLOCK-row block reordered by their 27-bit occupancy
  // pattern, so that the 128-row tiles hold rows with similar neighbourhoods -- more (tile, tap) pairs are empty (skipped)
  // and the present ones are denser (fewer wasted rows in 4-row gather groups).  Results are written through `perm`, so
  // the row order the caller sees is unchanged.
  DevBuf<unsigned long long> row_key;   // [n] (block << 27 | pattern), written by k_neighbours
  DevBuf<int> perm;          // [n_pad] tile position -> row (-1 in the padding)
  DevBuf<int> nbr_sorted;    // [27][n_pad] = nbr[k][perm[j]]
  DevBuf<uint32_t> tile_mask; // [n_pad / 128] taps present in every 128-row tile of the order in use (sorted or natural)
  bool tile_mask_sorted = false;
  PairList nbr_pairs;        // the same rules as 27 compacted (in, out) lists, built on first weight-gradient use
  // size-2/stride-2 link to the next coarser scale
  Level *coarse = nullptr;
  DevBuf<int> parent;        // [n]   coarse row of every fine row
  DevBuf<uint8_t> off8;      // [n]   (x&1)*4+(y&1)*2+(z&1)
  DevBuf<int> child;         // [8][coarse->n_pad] fine row or -1
  PairList child_pairs;      // 8 compacted (fine, coarse) lists
  PairList up_pairs;         // the same lists padded to whole 256-row tile groups (-1), for the one-tap-per-row products
  // Normal-guided kernels (OccuSeg's `use_normal`): every voxel carries an averaged surface normal; its orientation class
  // OrientedFilter(normal) in {0, 2, 4} (Metadata/RectangularRegions.h:12-31) permutes the weight taps its OUTPUT row uses
  // (SubmanifoldConvolutionRules.h:213-245 remap_rules_with_normal; strided: ConvolutionRules.h:27-90).
  bool guided = false;
  DevBuf<float> normal;      // [n][3]
  DevBuf<uint8_t> ori;       // [n] orientation class
  DevBuf<int> nbr_g;         // [27][n_pad] forward table with the taps of every output row permuted by its class
  DevBuf<int> nbr_t[3];      // [27][n_pad] per class c: tap k' -> the output row of class c this INPUT row feeds through weight tap k' (dgrad)
  PairList nbr_g_pairs;      // rule lists of the guided table (weight gradients)
};

struct Meta {
  int device = 0;
  std::vector<std::pair<Level *, int>> pending_counts;   // (scale, slot) whose n_rules is still on its way
  cudaStream_t last_stream = nullptr;   // stream of the most recent entry that used this handle (scn_meta_destroy frees on it)
  int batch = 0;
  int mode = 0;
  long long n_points = 0;
  std::vector<Level *> levels;
  // InputLayer rules, CSR form of the reference's [N][1+maxRepeat] table (CUDPPWrapper.cu:53-64)
  DevBuf<int> row_of_point;  // [P]
  DevBuf<int> rule_ptr;      // [N+1]
  DevBuf<int> rule_pts;      // [P] point ids grouped by row, original order inside a row
  // scn_bf16_operand(): a caller-owned bf16 copy of the fp32 matrix at hint_src, for the next convolution entry
  const float *hint_src = nullptr;
  void *hint_bf16 = nullptr;
  int hint_ready = 0;
  const float *ghint_src = nullptr;       // scn_grad_bf16(): a ready bf16 copy of the (dense) d_out of the next backward entry (one use)
  const void *ghint_bf16 = nullptr;
  const float *point_normals = nullptr;   // scn_input_normals(): [P,3] normals of the points of the next scn_input_layer_build
  int normal_guide_scale = 1 << 30;        // strided layers propagate normals / permute taps only from scales >= this size
  double *next_stats = nullptr;   // scn_out_stats(): column statistics wanted from the next scn_conv_fwd / scn_deconv_fwd (one use)
  int next_dilation = 1;     // scn_subm_dilation(): dilation of the next submanifold entry (one use)
  long long next_grad_ld = 0;  // scn_grad_stride(): row stride (floats) of d_out of the next backward entry (one use; 0 = dense)
  // scn_bn_bwd_fusion(): the BatchNorm whose backward the next *_bwd entry folds into its dgrad epilogue (one use)
  struct BnBwdHint {
    const float *x = nullptr, *mean = nullptr, *invstd = nullptr, *gamma = nullptr, *beta = nullptr;
    float leak = 0.f;
    double *acc = nullptr;
  } bnb;
  ~Meta();
};

Level *find_level(Meta *m, const int64_t size[3]);

// meta.cu
void build_input_level(Meta *m, const int64_t size[3], const int64_t *coords, bool on_device, long long P, int batch,
                       int mode, cudaStream_t s);
// defer = true (only while a whole chain of scales is built, prebuild_scales): the rule count is copied to pinned host memory
// without waiting; resolve_rule_counts() reads it after the next synchronisation the chain performs anyway
void ensure_neighbour_table(Meta *m, Level *L, cudaStream_t s, bool defer = false);
void resolve_rule_counts(Meta *m, cudaStream_t s, bool synchronise);
Level *dilated_level(Meta *m, Level *L, int rate, cudaStream_t s);    // rate 1 = L itself; table built on return
void ensure_guided_tables(Meta *m, Level *L, cudaStream_t s);          // nbr_g / nbr_t of a scale that carries normals
constexpr int SORT_BLOCK_DEFAULT = 262144;
// Deterministic mode (scn_deterministic / SCN_DETERMINISTIC=1): every reduction whose partial results are otherwise merged with
// floating-point atomics (weight gradients, column statistics) writes one partial per row range / CTA into scratch and
// sum_partials adds them in range order -- run-to-run bit-identical results for a few per cent of extra traffic.
bool deterministic();
int set_deterministic(int on);   // returns the previous setting
// out[e] = sum over p = 0 .. parts-1 (in that order) of partial[p * n + e]
void sum_partials(const float *partial, int parts, long long n, float *out, cudaStream_t s);
void sum_partials(const double *partial, int parts, long long n, double *out, cudaStream_t s);
bool tile_sort_enabled();
int set_tile_sort(int block);   // returns the previous setting
void ensure_sorted_table(Level *L, cudaStream_t s);     // builds perm / nbr_sorted / tile_mask (no-op when already built)
Level *ensure_coarse_level(Meta *m, Level *fine, const int64_t coarse_size[3], cudaStream_t s);

// io.cu
void input_layer_fwd(Meta *m, const float *feats, int C, float *out, cudaStream_t s);
void input_layer_bwd(Meta *m, const float *d_out, int C, float *d_feats, cudaStream_t s);
void output_layer_fwd(Meta *m, const float *in, int C, float *out, cudaStream_t s);
void output_layer_bwd(Meta *m, const float *d_out, int C, float *d_in, cudaStream_t s);
void float_coords(const float *xyz, long long n, const float offset[3], int batch_index, float full_scale, long long *coords,
                  uint8_t *keep, cudaStream_t s);

// conv_simt.cu / conv_small.cu / conv_tma.cu ---------------------------------------------------------------------------
// Generic "table convolution".  tbl is [V][stride] (row index or -1).
//  GATHER : out[o,:]        = sum_k in[tbl[wk(k)][o],:] * Wk      for o in [0,n_rows)
//  SCATTER: out[tbl[k][p],:] =        in[p,:]            * Wk      for p in [0,n_rows), tbl>=0
// Wk is weight[k] ([Cin][Cout]) or, when transpose_w, weight[k]^T read from the original [V][Cout'][Cin'] array.
struct ConvArgs {
  const float *in = nullptr;
  const float *weight = nullptr;   // [V][c_in][c_out] as seen by THIS product (fp32 FMA kernels)
  const void *weight_nk = nullptr; // [V][c_out][c_in] as seen by THIS product (tensor-core kernel: K-major B operand), fp32 or bf16
  bool bf16 = false;               // tensor-core kernel: `in` and `weight_nk` are bf16 ([rows][c] uint16), fp32 accumulate and output
  const float *bias = nullptr;     // GATHER only
  const float *residual = nullptr; // tensor-core kernel only: [n_rows, c_out] fp32 added to the result
  double *stats = nullptr;         // tensor-core kernel only: [2][c_out] column sums / sums of squares of the result
  // tensor-core kernel only: inference BatchNorm + (leaky) ReLU of the following layer applied in the epilogue
  // (out = leaky(scale * acc + shift)), optional bf16 copy of the result
  const float *ep_scale = nullptr, *ep_shift = nullptr;
  float ep_leak = 1.f;
  uint16_t *out_bf16 = nullptr;
  // tensor-core kernel only (dgrad products): fused backward of the BatchNorm that produced this product's input operand --
  // see ConvParams in conv_tma.cu.  bnb_x [n_rows, c_out] fp32 + the BatchNorm's saved statistics and affine parameters;
  // `stats` receives (sum d', sum d'*x)
  const float *bnb_x = nullptr, *bnb_mean = nullptr, *bnb_invstd = nullptr, *bnb_gamma = nullptr, *bnb_beta = nullptr;
  float bnb_leak = 0.f;
  float *out = nullptr;
  const int *tbl = nullptr;
  int tbl_stride = 0;
  int n_rows = 0;                  // rows iterated (outputs for GATHER, inputs for SCATTER)
  int in_rows = 0;                 // rows of the `in` matrix (TMA out-of-bounds index = zero row)
  int V = 27;
  int c_in = 0, c_out = 0;
  long long n_rules = 0;           // live table entries (for the algorithmic-bytes model only)
  bool mirror = false;             // GATHER: use table row V-1-k for weight tap k (dgrad of a submanifold conv)
  // tensor-core kernel, one-tap-per-row products (deconvolution forward, strided dgrad): V = 1, tbl = the gathered row of
  // every rule, rules grouped by tap in items of `rows_per_item` rows (item_off[k] = first item of weight tap k, n_taps + 1
  // entries), result row j goes to out_rows[j] (< 0 or >= out_limit: padding)
  const int *out_rows = nullptr;
  const int *item_off = nullptr;
  const uint32_t *tile_mask = nullptr;   // tensor-core kernel: per 128-row tile, bit t = table row t has a row in the tile
  int rows_per_item = 0, n_taps = 0, out_limit = 0;
  bool scatter = false;
};
void conv_simt(const ConvArgs &a, cudaStream_t s);
bool conv_small_supported(const ConvArgs &a);
void conv_small(const ConvArgs &a, cudaStream_t s);   // c_in <= 4: one thread per output row (conv_small.cu)
bool conv_tma_supported(const ConvArgs &a);            // tcgen05 + TMA gather4 (conv_tma.cu); a.bf16 selects the operand type
void conv_tma(const ConvArgs &a, cudaStream_t s);

// weight preparation: dst[k][co][ci] = src[k][ci][co]   (per-tap transpose, used by every dgrad)
void transpose_weight(const float *src, float *dst, int V, int c_in, int c_out, cudaStream_t s);
// dst[k][co][ci] = bf16(src[k][ci][co]): transpose and operand rounding in one pass
void transpose_weight_bf16(const float *src, uint16_t *dst, int V, int c_in, int c_out, cudaStream_t s);
// bf16 operand copies (round to nearest even): dst[i] = bf16(src[i]); n must be even
void cast_bf16(const float *src, uint16_t *dst, long long n, cudaStream_t s);
void axpy(const float *x, float *y, long long n, cudaStream_t s);     // y += x
// the same from a row-strided source (rows `ld` floats apart) into a dense copy
void cast_bf16_rows(const float *src, long long ld, long long rows, int cols, uint16_t *dst, cudaStream_t s);

// dW[k] = sum_r A[ia(k,r),:]^T * B[ib(k,r),:]    A: [*,c_a], B: [*,c_b], dW: [V][c_a][c_b] (zeroed here)
//   table_on_a:  ia = tbl[k][r], ib = r      (submanifold / strided conv: A = input, B = d_out)
//   !table_on_a: ia = r, ib = tbl[k][r]      (deconvolution: A = coarse input, B = fine d_out)
struct WgradArgs {
  const float *a = nullptr;
  const float *b = nullptr;
  float *dw = nullptr;
  const int *tbl = nullptr;
  int tbl_stride = 0;
  int n_rows = 0;
  int g_rows = 0;                  // rows of the gathered matrix (a when table_on_a, else b)
  int V = 27;
  int c_a = 0, c_b = 0;
  long long n_rules = 0;
  int s_rows = 0;                  // rows of the other ("stationary") matrix
  bool bf16 = false;               // a/b are bf16 copies ([rows][c] uint16) instead of fp32
  // compacted rule lists (tensor-core kernel): see PairList
  const int *gi = nullptr, *si = nullptr, *blk_item = nullptr;
  int n_blk = 0, blk_rows = 0;
  bool table_on_a = true;
  long long part_stride = 0;       // deterministic mode: CTAs of row range r accumulate into dw + r * part_stride (zeroed scratch)
};
void wgrad_simt(const WgradArgs &a, cudaStream_t s);
bool wgrad_small_supported(const WgradArgs &a);
void wgrad_small(const WgradArgs &a, cudaStream_t s);
bool wgrad_tma_supported(const WgradArgs &a);
void wgrad_tma(const WgradArgs &a, cudaStream_t s);

void bias_grad(const float *d_out, float *d_bias, long long n_rows, int C, cudaStream_t s);

// scatter.cu
void resolution_scatter(const int *lr_xyz, long long n_lr, const int *hr_xyz, long long n_hr, int stride, int *hr2lr,
                        cudaStream_t s);

// bn.cu
void bn_fwd(const float *in, float *out, uint16_t *out_bf16, const double *stats_in, float *save_mean, float *save_invstd, float *running_mean,
            float *running_var, const float *gamma, const float *beta, long long n, int C, float eps, float momentum,
            bool train, float leakiness, cudaStream_t s);
// inference coefficients exactly as bn_fwd(train = false) uses them: scale = invstd*gamma, shift = beta - mean*scale
void bn_eval_coeffs(const float *running_mean, const float *running_var, const float *gamma, const float *beta, int C,
                    float eps, float *scale, float *shift, cudaStream_t s);
void bn_bwd(const float *in, const float *out, const float *d_out, const float *save_mean, const float *save_invstd,
            const float *gamma, const float *beta, const float *d_in_add, float *d_in, float *d_gamma, float *d_beta, long long n, int C, float leakiness,
            cudaStream_t s);
// coefficients of the activation mask for a fused BatchNorm backward: coef[0][c] = w = invstd*gamma, coef[1][c] = b = beta - mean*w
void bn_mask_coeffs(const float *save_mean, const float *save_invstd, const float *gamma, const float *beta, int C, float *coef,
                    cudaStream_t s);
// second half of bn_bwd when a convolution epilogue already produced the masked gradient d' and acc = (sum d', sum d'*x):
// d_in = (d' - mean(d') - (x - mean) * k) * invstd * gamma (+ d_in_add); d_gamma, d_beta.  d_in may alias d_masked.
void bn_bwd_apply(const float *in, const float *d_masked, const double *acc, const float *save_mean, const float *save_invstd,
                  const float *gamma, const float *d_in_add, long long ld_add, float *d_in, uint16_t *d_in_bf16, float *d_gamma,
                  float *d_beta, long long n, int C, cudaStream_t s);

}  // namespace scn

// Rulebook construction on the GPU: voxelisation, open-addressing coordinate hash, 27-offset
// neighbour query, size-2/stride-2 links.  Replaces the reference's Metadata layer + cudpp:
//   Metadata/IOLayersRules.h:136-202, CUDA/CUDPPWrapper.{cu,hpp}, extra/cudpp/src/cudpp_hash/*,
//   CUDA/SubmanifoldRules_cuda.{cu,cpp}:20-203, Metadata/ConvolutionRules.h:344-378.
// Everything stays device-resident; the reference copies every rule list to the host and back.
#include "common.cuh"
#include <cub/cub.cuh>

namespace scn {

Meta::~Meta() {
  for (Level *L : levels) delete L;
}
Level::~Level() {
  for (Level *D : dilated) delete D;
}

Level *find_level(Meta *m, const int64_t size[3]) {
  for (Level *L : m->levels)
    if (L->size[0] == size[0] && L->size[1] == size[1] && L->size[2] == size[2]) return L;
  return nullptr;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : g);
}

// -----------------------------------------------------------------------------------------------------
// keys
// -----------------------------------------------------------------------------------------------------
__global__ void k_point_keys(const int64_t *__restrict__ coords, long long P, int batch, uint64_t *__restrict__ keys,
                             int *__restrict__ idx, int *__restrict__ err) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= P) return;
  // one 32-byte row per point: two 16-byte loads
  const longlong2 *row = reinterpret_cast<const longlong2 *>(coords + 4 * i);
  longlong2 xy = row[0], zb = row[1];
  bool ok = xy.x >= 0 && xy.x < COORD_LIMIT && xy.y >= 0 && xy.y < COORD_LIMIT && zb.x >= 0 && zb.x < COORD_LIMIT &&
            zb.y >= 0 && zb.y < batch;
  if (!ok) atomicExch(err, 1);
  keys[i] = ok ? make_key((uint32_t)zb.y, (uint32_t)zb.x, (uint32_t)xy.y, (uint32_t)xy.x) : EMPTY_KEY - 1;
  idx[i] = (int)i;
}

__global__ void k_coarse_keys(const uint64_t *__restrict__ fine, int n, uint64_t *__restrict__ ckeys,
                              int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = fine[i];
  // halve x, y, z fields in place (out = in/2, ConvolutionRules.h:364); batch field untouched
  uint64_t lo = (k >> 1) & 0x00007FFF7FFF7FFFull;
  ckeys[i] = (k & 0xFFFF000000000000ull) | lo;
  idx[i] = i;
}

// rows = runs of equal keys in the sorted list; fan the row id back out to the members of each run
// run (= row) holding sorted position j: the last r with run_ptr[r] <= j
__device__ __forceinline__ int run_of(const int *__restrict__ run_ptr, int n_runs, int j) {
  int lo = 0, hi = n_runs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&run_ptr[mid]) <= j) lo = mid;
    else hi = mid;
  }
  return lo;
}

// One thread per POINT (a binary search for its run), so a voxel holding thousands of duplicate points costs no more than any other
__global__ void k_rows_of_points(const int *__restrict__ run_ptr, const int *__restrict__ sorted_idx, int n_runs, int n_points,
                                 int *__restrict__ row_of_point) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_points) return;
  row_of_point[sorted_idx[j]] = run_of(run_ptr, n_runs, j);
}

// stride-2 link: parent / offset of every fine row, child table of every coarse row
__global__ void k_link_levels(const int *__restrict__ run_ptr, const int *__restrict__ sorted_idx,
                              const uint64_t *__restrict__ fine_keys, int n_runs, int n_fine, int child_stride,
                              int *__restrict__ parent, uint8_t *__restrict__ off8, int *__restrict__ child) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per fine row (sorted position)
  if (j >= n_fine) return;
  {
    const int r = run_of(run_ptr, n_runs, j);
    int i = sorted_idx[j];
    uint64_t k = fine_keys[i];
    // offset index ((x&1)*2 + (y&1))*2 + (z&1): SubmanifoldRules_cuda.cu:549-554
    int off = (int)(((k & 1ull) << 2) | (((k >> 16) & 1ull) << 1) | ((k >> 32) & 1ull));
    parent[i] = r;
    off8[i] = (uint8_t)off;
    child[off * child_stride + r] = i;
  }
}

// -----------------------------------------------------------------------------------------------------
// normal-guided rules
// -----------------------------------------------------------------------------------------------------
// tap permutations per orientation class: SubmanifoldRules_cuda.cu:8-15 (27 taps), ConvolutionRules.h:28-33 (8 taps)
__device__ __constant__ unsigned char c_rot27[27 * 6] = {
    0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,
    24,25,26,21,22,23,18,19,20,15,16,17,12,13,14,9,10,11,6,7,8,3,4,5,0,1,2,
    6,7,8,15,16,17,24,25,26,3,4,5,12,13,14,21,22,23,0,1,2,9,10,11,18,19,20,
    18,19,20,9,10,11,0,1,2,21,22,23,12,13,14,3,4,5,24,25,26,15,16,17,6,7,8,
    2,11,20,5,14,23,8,17,26,1,10,19,4,13,22,7,16,25,0,9,18,3,12,21,6,15,24,
    18,9,0,21,12,3,24,15,6,19,10,1,22,13,4,25,16,7,20,11,2,23,14,5,26,17,8};
__device__ __constant__ unsigned char c_rot8[8 * 6] = {0,1,2,3,4,5,6,7, 6,7,4,5,2,3,0,1, 2,3,6,7,0,1,4,5,
                                                       4,5,0,1,6,7,2,3, 1,5,3,7,0,4,2,6, 4,0,6,2,5,1,7,3};

// OrientedFilter, Metadata/RectangularRegions.h:12-31: the dominant axis of the normal -> class 0 (x), 2 (y), 4 (z)
__device__ __forceinline__ int oriented_filter(float nx, float ny, float nz) {
  const float x = fabsf(nx), y = fabsf(ny), z = fabsf(nz);
  if (x >= y && x >= z) return 0;
  if (y >= x && y >= z) return 2;
  if (z >= x && z >= y) return 4;
  return 0;
}
__device__ __forceinline__ void normalize3(float &x, float &y, float &z) {      // Float3::normalize, Metadata.h:94-100
  // no fused multiply-add: the host code this mirrors rounds every product and sum
  float mag = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  if (mag < 1e-8f) return;
  mag = __fdiv_rn(1.f, mag);
  x *= mag; y *= mag; z *= mag;
}

// per-voxel normal = normalised mean of its points' normals, summed in rule order (CUDA/IOLayers.cpp:39-66)
__global__ void k_voxel_normals(const float *__restrict__ pn, const int *__restrict__ rule_ptr, const int *__restrict__ rule_pts,
                                int n, float *__restrict__ normal, uint8_t *__restrict__ ori) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float x = 0.f, y = 0.f, z = 0.f;
  const int b = rule_ptr[r], e = rule_ptr[r + 1];
  for (int j = b; j < e; ++j) {
    const float *q = pn + 3ll * rule_pts[j];
    x += q[0]; y += q[1]; z += q[2];
  }
  if (e > b) { x = __fdiv_rn(x, (float)(e - b)); y = __fdiv_rn(y, (float)(e - b)); z = __fdiv_rn(z, (float)(e - b)); }
  normalize3(x, y, z);
  normal[3 * r] = x; normal[3 * r + 1] = y; normal[3 * r + 2] = z;
  ori[r] = (uint8_t)oriented_filter(x, y, z);
}

// coarse normal = normalised mean of the children's normals (ConvolutionRules.h:56-72); then the 8 taps of the coarse row are
// permuted by its class (:80-88) -- in place in the child table, with the offsets of the children updated to match
__global__ void k_coarse_normals_and_taps(const float *__restrict__ fine_normal, int nc, int child_stride, int *__restrict__ child,
                                          uint8_t *__restrict__ off8, float *__restrict__ normal, uint8_t *__restrict__ ori) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nc) return;
  int c[8];
  float x = 0.f, y = 0.f, z = 0.f;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    c[k] = child[k * child_stride + p];
    if (c[k] >= 0) {
      x += fine_normal[3 * c[k]]; y += fine_normal[3 * c[k] + 1]; z += fine_normal[3 * c[k] + 2];
      ++cnt;
    }
  }
  if (cnt > 0) { x = __fdiv_rn(x, (float)cnt); y = __fdiv_rn(y, (float)cnt); z = __fdiv_rn(z, (float)cnt); }
  normalize3(x, y, z);
  normal[3 * p] = x; normal[3 * p + 1] = y; normal[3 * p + 2] = z;
  const int o = oriented_filter(x, y, z);
  ori[p] = (uint8_t)o;
  int t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) t[k] = -1;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int k2 = c_rot8[o * 8 + k];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j == k2) t[j] = c[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    child[k * child_stride + p] = t[k];
    if (t[k] >= 0) off8[t[k]] = (uint8_t)k;
  }
}

// guided tables of a scale: forward nbr_g[rot[ori(o)][k]][o] = nbr[k][o]; dgrad, per class c, nbr_t[c][rot[c][k]][i] = the output
// row o = nbr[26-k][i] (the row for which i sits at offset k) if ori(o) == c
__global__ void k_guide_tables(const int *__restrict__ nbr, const uint8_t *__restrict__ ori, int n, int stride,
                               int *__restrict__ nbr_g, int *__restrict__ t0, int *__restrict__ t1, int *__restrict__ t2) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int c = ori[r];
  for (int k = 0; k < 27; ++k) {
    nbr_g[(long long)c_rot27[c * 27 + k] * stride + r] = nbr[(long long)k * stride + r];
    const int o = nbr[(long long)(26 - k) * stride + r];
    if (o >= 0) {
      const int co = ori[o];
      int *t = co == 0 ? t0 : co == 2 ? t1 : t2;
      t[(long long)c_rot27[co * 27 + k] * stride + r] = o;
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// open-addressing hash (linear probing, load <= 0.5).  Replaces cudpp's cuckoo tables
// (extra/cudpp/src/cudpp_hash/hash_table.cuh:94-295): no stash, no rebuild loop, one CAS per insert.
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_key(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

// An entry is {key, row} in 16 bytes: a probe is ONE 16-byte load (one 32-byte sector).  The neighbour query is bound by the
// sectors its random probes pull from L2 (52 M probes at level 0), and separate key / row arrays cost two per hit.
__global__ void k_hash_insert(const uint64_t *__restrict__ keys, int n, ulonglong2 *__restrict__ htab, uint32_t mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = keys[i];
  uint32_t slot = hash_key(k) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&htab[slot].x, (unsigned long long)EMPTY_KEY, (unsigned long long)k);
    if (prev == EMPTY_KEY) {  // keys are unique, so a slot is never claimed twice for the same key
      htab[slot].y = (unsigned long long)i;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ int hash_find(uint64_t k, const ulonglong2 *__restrict__ htab, uint32_t mask) {
  uint32_t slot = hash_key(k) & mask;
  while (true) {
    const ulonglong2 e = __ldg(&htab[slot]);
    if (e.x == k) return (int)e.y;
    if (e.x == EMPTY_KEY) return -1;
    slot = (slot + 1) & mask;
  }
}

// Bit of tap t in the 27-bit sort key of the tile order.  Rows are sorted by this key inside every sort block, so the most
// significant bits are constant inside a tile and only the least significant ones vary; a tap costs a tile a pipeline item as
// soon as ONE of its rows has it.  Taps that are rarely present (the 8 corners, then the 12 edges) therefore go to the top --
// tiles then either have them in every row or in none -- and the taps present in most rows (the 6 faces) to the bottom, where
// mixing costs little.  tools/sim_tile_order.py: items per 256-row group on S250k level 0 19.7 (tap order) -> 17.9, S1M 17.0 ->
// 15.9; an order adapted to the measured presence of every tap gains another 1 %.
__host__ __device__ constexpr int tap_key_bit(int t) {
  // class of a tap = number of non-zero offsets: 0 centre, 1 face, 2 edge, 3 corner; key order (most significant first):
  // centre, corners, edges, faces; inside a class by tap index
  int cls_rank[4] = {0, 3, 2, 1};
  auto cls = [](int u) { return (u / 9 != 1) + ((u / 3) % 3 != 1) + (u % 3 != 1); };
  int before = 0;
  for (int u = 0; u < 27; ++u)
    if (cls_rank[cls(u)] < cls_rank[cls(t)] || (cls_rank[cls(u)] == cls_rank[cls(t)] && u < t)) ++before;
  return 26 - before;
}
__device__ __forceinline__ uint32_t pattern_of_key(uint32_t key) {
  uint32_t pat = 0;
#pragma unroll
  for (int t = 0; t < 27; ++t) pat |= ((key >> tap_key_bit(t)) & 1u) << t;
  return pat;
}

// One thread per output voxel, 27 probes.  Offset order k=(dx+1)*9+(dy+1)*3+(dz+1) follows the GPU builder
// (CUDA/SubmanifoldRules_cuda.cu:63-73), NOT the dormant CPU-grid enumeration.  A neighbour exists only
// inside the same sample (the reference keeps one hash per sample, Metadata.h:110-122).  Writes are
// coalesced: consecutive threads -> consecutive rows of nbr[k][*].
__global__ void k_neighbours(const uint64_t *__restrict__ keys, int n, int stride, const ulonglong2 *__restrict__ htab,
                             uint32_t mask, int *__restrict__ nbr,
                             unsigned long long *__restrict__ n_rules, unsigned long long *__restrict__ row_key,
                             int sort_block, int dil) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int hits = 0;
  if (i < n) {
    uint32_t pattern = 0;
    uint64_t k = keys[i];
    int x = (int)(k & 0xFFFF), y = (int)((k >> 16) & 0xFFFF), z = (int)((k >> 32) & 0xFFFF);
    uint32_t b = (uint32_t)(k >> 48);
    int t = 0;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz, ++t) {
          int r;
          if (t == 13) {
            r = i;
          } else if (dil == 1 && dy == 0 && dz == 0) {
            // the x neighbours of a row are its neighbours in the sorted key array, if they exist (x is the lowest key field
            // and 65535 is not a coordinate, so key -+ 1 never reaches into another (y, z)): no probe
            const int j = i + dx;
            r = (j >= 0 && j < n && __ldg(&keys[j]) == k + (uint64_t)(long long)dx) ? j : -1;
          } else {
            int qx = x + dx * dil, qy = y + dy * dil, qz = z + dz * dil;
            bool inside = qx >= 0 && qy >= 0 && qz >= 0 && qx < COORD_LIMIT && qy < COORD_LIMIT && qz < COORD_LIMIT;
            r = inside ? hash_find(make_key(b, qz, qy, qx), htab, mask) : -1;
          }
          nbr[t * stride + i] = r;
          hits += (r >= 0);
          pattern |= (r >= 0 ? 1u : 0u) << tap_key_bit(t);
        }
    if (row_key) row_key[i] = ((unsigned long long)(i / sort_block) << 27) | pattern;
  }
  // block-level count -> one atomic per block
  typedef cub::BlockReduce<int, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  int total = BR(tmp).Sum(hits);
  if (threadIdx.x == 0 && total) atomicAdd(n_rules, (unsigned long long)total);
}

// ---- per-tap compaction of a [V][stride] table into padded (entry, column) lists ------------------------------
// Ranks come from per-block counts: cnt[k][b] = rules of tap k in columns [b*BLK_ROWS, (b+1)*BLK_ROWS); an exclusive scan over
// the V * n_blk counts (tap-major) gives base[k][b] = number of rules before that block, and the scatter kernel adds the rank
// inside the block.  (Until round 2 a scan over all V * stride flags wrote a rank per table entry: 4 reads + 1 write of the
// table's size instead of 2 reads.)  Tap k owns rules base[k][0] .. base[k+1][0]-1; its list starts at item_off[k] (`unit`
// rules per item, every tap rounded up).
constexpr int PB_THREADS = BLK_ROWS / 4;      // one int4 (4 columns) per thread
static_assert(PB_THREADS % 32 == 0 && PB_THREADS <= 1024, "block geometry of the pair-list kernels");

__device__ __forceinline__ int4 load_cols4(const int *__restrict__ tbl, int k, int stride, int col) {
  return col < stride ? __ldg(reinterpret_cast<const int4 *>(tbl + (long long)k * stride + col)) : make_int4(-1, -1, -1, -1);
}
// exclusive prefix of `c` over the thread block (PB_THREADS threads); *total receives the block sum
__device__ __forceinline__ int block_exclusive(int c, int *total) {
  __shared__ int warp_sum[PB_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += v;
  }
  if (lane == 31) warp_sum[w] = inc;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int j = 0; j < PB_THREADS / 32; ++j) {
    const int v = warp_sum[j];
    if (j < w) before += v;
    all += v;
  }
  *total = all;
  return before + inc - c;
}
// grid (n_blk, V)
__global__ void __launch_bounds__(PB_THREADS) k_count_rules(const int *__restrict__ tbl, int stride, int n_blk,
                                                            int *__restrict__ cnt) {
  const int k = blockIdx.y, b = blockIdx.x;
  const int4 t = load_cols4(tbl, k, stride, b * BLK_ROWS + 4 * threadIdx.x);
  const int c = (t.x >= 0) + (t.y >= 0) + (t.z >= 0) + (t.w >= 0);
  int total;
  block_exclusive(c, &total);
  if (threadIdx.x == 0) cnt[k * n_blk + b] = total;
}
// base = exclusive scan of cnt, V * n_blk + 1 entries (the last one = all rules)
__global__ void k_pair_offsets(const int *__restrict__ base, int V, int n_blk, int unit, int *__restrict__ item_off,
                               int *__restrict__ rank_base) {
  if (threadIdx.x != 0) return;
  int items = 0;
  for (int k = 0; k < V; ++k) {
    const int b = base[k * n_blk], e = base[(k + 1) * n_blk];
    item_off[k] = items;
    rank_base[k] = b;
    items += (e - b + unit - 1) / unit;
  }
  item_off[V] = items;
}
__global__ void k_block_items(const int *__restrict__ base, int V, int n_blk, int unit, const int *__restrict__ item_off,
                              const int *__restrict__ rank_base, int *__restrict__ blk_item) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * (n_blk + 1)) return;
  const int k = i / (n_blk + 1), b = i - k * (n_blk + 1);
  blk_item[i] = b == n_blk ? item_off[k + 1] : item_off[k] + (base[k * n_blk + b] - rank_base[k]) / unit;
}
// grid (n_blk, V)
__global__ void __launch_bounds__(PB_THREADS) k_scatter_pairs(const int *__restrict__ tbl, const int *__restrict__ base,
                                                              int stride, int n_blk, int unit,
                                                              const int *__restrict__ item_off,
                                                              const int *__restrict__ rank_base, int *__restrict__ gi,
                                                              int *__restrict__ si) {
  const int k = blockIdx.y, b = blockIdx.x;
  const int col = b * BLK_ROWS + 4 * threadIdx.x;
  const int4 t = load_cols4(tbl, k, stride, col);
  const int c = (t.x >= 0) + (t.y >= 0) + (t.z >= 0) + (t.w >= 0);
  int total;
  const int pre = block_exclusive(c, &total);
  if (!c) return;
  long long pos = (long long)item_off[k] * unit + (base[k * n_blk + b] + pre - rank_base[k]);
  if (t.x >= 0) { gi[pos] = t.x; si[pos] = col; ++pos; }
  if (t.y >= 0) { gi[pos] = t.y; si[pos] = col + 1; ++pos; }
  if (t.z >= 0) { gi[pos] = t.z; si[pos] = col + 2; ++pos; }
  if (t.w >= 0) { gi[pos] = t.w; si[pos] = col + 3; ++pos; }
}

__global__ void k_fill_int(int *p, long long n, int v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// -----------------------------------------------------------------------------------------------------
// sort + run-length encode: (keys, payload) -> unique keys, run pointers, payload grouped by run.
// Radix sort is stable, so members of a run keep their original order -- this is what fixes the
// summation order of duplicate points in InputLayer mode 3/4.
// -----------------------------------------------------------------------------------------------------
__global__ void k_fill_int(int *p, long long n, int v);

struct Runs {
  DevBuf<uint64_t> unique;   // [n_runs] (allocated n)
  DevBuf<int> ptr;           // [n_runs+1] (allocated n+1)
  DevBuf<int> sorted_idx;    // [n]
  int n_runs = 0;
};

static void sort_and_group(DevBuf<uint64_t> &keys, DevBuf<int> &idx, long long n, int end_bit, Runs &out,
                           cudaStream_t s) {
  DevBuf<uint64_t> keys_sorted;
  DevBuf<int> counts, d_nruns;
  keys_sorted.alloc(n, s);
  out.sorted_idx.alloc(n, s);
  out.unique.alloc(n, s);
  out.ptr.alloc(n + 1, s);
  counts.alloc(n, s);
  d_nruns.alloc(1, s);

  size_t t1 = 0, t2 = 0, t3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, keys.p, keys_sorted.p, idx.p, out.sorted_idx.p, (int)n, 0, end_bit, s);
  cub::DeviceRunLengthEncode::Encode(nullptr, t2, keys_sorted.p, out.unique.p, counts.p, d_nruns.p, (int)n, s);
  cub::DeviceScan::ExclusiveSum(nullptr, t3, counts.p, out.ptr.p, (int)n, s);
  DevBuf<uint8_t> tmp;
  tmp.alloc(std::max(t1, std::max(t2, t3)), s);
  size_t tb = tmp.n;
  SCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys_sorted.p, idx.p, out.sorted_idx.p, (int)n, 0,
                                           end_bit, s));
  count_launch(4);
  tb = tmp.n;
  SCN_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tb, keys_sorted.p, out.unique.p, counts.p, d_nruns.p, (int)n, s));
  count_launch(2);
  int h_runs = 0;
  SCN_CUDA(cudaMemcpyAsync(&h_runs, d_nruns.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));   // the caller needs the row count to size its tensors
  out.n_runs = h_runs;
  tb = tmp.n;
  // exclusive sum over n_runs+1 entries: counts[n_runs] is garbage but only feeds ptr[n_runs+1..]; write the
  // terminal pointer explicitly instead
  SCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, out.ptr.p, h_runs, s));
  count_launch(2);
  k_fill_int<<<1, 32, 0, s>>>(out.ptr.p + h_runs, 1, (int)n);
  SCN_LAUNCH_CHECK();
  keys_sorted.release(s);
  counts.release(s);
  d_nruns.release(s);
  tmp.release(s);
}

void build_pair_list(PairList &out, const int *tbl, int V, int stride, long long n_rules, int unit, int pad_byte,
                     cudaStream_t s) {
  if (out.item_off.p) return;
  out.unit = unit;
  const long long n_flat = (long long)V * stride;
  // table read twice (count + scatter, 4 B each) + 8 B per rule written
  ProfScope ps(PK_RULEBOOK, 8.0 * (double)n_flat + 8.0 * (double)n_rules, 0.0, s);
  SCN_CHECK(n_flat > 0 && n_flat < (1ll << 31), "rule table too large for 32-bit ranks");
  out.n_items_ub = (n_rules + (long long)(unit - 1) * V) / unit + 1;
  out.gi.alloc((size_t)out.n_items_ub * unit, s);
  out.si.alloc((size_t)out.n_items_ub * unit, s);
  out.item_off.alloc(V + 1, s);
  SCN_CUDA(cudaMemsetAsync(out.gi.p, pad_byte, sizeof(int) * out.gi.n, s));   // 0x7F -> PAIR_PAD, 0xFF -> -1
  SCN_CUDA(cudaMemsetAsync(out.si.p, pad_byte, sizeof(int) * out.si.n, s));
  out.n_blk = (stride + BLK_ROWS - 1) / BLK_ROWS;
  const int n_cnt = V * out.n_blk;
  DevBuf<int> cnt, base, rank_base;
  cnt.alloc((size_t)n_cnt + 1, s);
  base.alloc((size_t)n_cnt + 1, s);
  rank_base.alloc(V, s);
  SCN_CUDA(cudaMemsetAsync(cnt.p + n_cnt, 0, sizeof(int), s));
  const dim3 grid(out.n_blk, V);
  k_count_rules<<<grid, PB_THREADS, 0, s>>>(tbl, stride, out.n_blk, cnt.p);
  SCN_LAUNCH_CHECK();
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, base.p, n_cnt + 1, s);
  DevBuf<uint8_t> tmp;
  tmp.alloc(tb, s);
  SCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, base.p, n_cnt + 1, s));
  count_launch(2);
  k_pair_offsets<<<1, 32, 0, s>>>(base.p, V, out.n_blk, unit, out.item_off.p, rank_base.p);
  SCN_LAUNCH_CHECK();
  k_scatter_pairs<<<grid, PB_THREADS, 0, s>>>(tbl, base.p, stride, out.n_blk, unit, out.item_off.p, rank_base.p, out.gi.p,
                                              out.si.p);
  SCN_LAUNCH_CHECK();
  out.blk_item.alloc((size_t)V * (out.n_blk + 1), s);
  k_block_items<<<grid_for((long long)V * (out.n_blk + 1), 256), 256, 0, s>>>(base.p, V, out.n_blk, unit, out.item_off.p,
                                                                             rank_base.p, out.blk_item.p);
  SCN_LAUNCH_CHECK();
  cnt.release(s);
  base.release(s);
  rank_base.release(s);
  tmp.release(s);
}

static int bits_for(unsigned v) {
  int b = 0;
  while ((1u << b) <= v && b < 16) ++b;
  return b;
}

void build_input_level(Meta *m, const int64_t size[3], const int64_t *coords, bool on_device, long long P, int batch,
                       int mode, cudaStream_t s) {
  SCN_CHECK(mode == 3 || mode == 4, "InputLayer: only modes 3 (sum) and 4 (mean) are supported, like the reference GPU path");
  SCN_CHECK(m->levels.empty(), "InputLayer: this handle already holds a batch");
  SCN_CHECK(P > 0 && P < (1ll << 31) - 1, "InputLayer: point count must be in (0, 2^31)");
  SCN_CHECK(batch > 0 && batch < COORD_LIMIT, "InputLayer: bad batch size");
  m->batch = batch;
  m->mode = mode;
  m->n_points = P;
  // coordinates read (32 B per point), keys sorted with the point index (12 B in, 12 B out), row of every point + grouped list (8 B)
  ProfScope ps(PK_RULEBOOK, (32.0 + 24 + 8) * (double)P, 0.0, s);

  DevBuf<int64_t> dcoords;
  const int64_t *dc = coords;
  if (!on_device) {
    dcoords.alloc((size_t)P * 4, s);
    SCN_CUDA(cudaMemcpyAsync(dcoords.p, coords, sizeof(int64_t) * 4 * P, cudaMemcpyHostToDevice, s));
    dc = dcoords.p;
  }
  DevBuf<uint64_t> keys;
  DevBuf<int> idx, err;
  keys.alloc(P, s);
  idx.alloc(P, s);
  err.alloc(1, s);
  SCN_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), s));
  k_point_keys<<<grid_for(P, 256), 256, 0, s>>>(dc, P, batch, keys.p, idx.p, err.p);
  SCN_LAUNCH_CHECK();
  int h_err = 0;
  SCN_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));

  Runs runs;
  sort_and_group(keys, idx, P, 48 + bits_for((unsigned)(batch > 1 ? batch - 1 : 1)), runs, s);
  SCN_CHECK(h_err == 0, "InputLayer: coordinate outside [0,65535) or batch index outside [0,batch_size)");

  Level *L = new Level();
  for (int d = 0; d < 3; ++d) L->size[d] = size[d];
  L->n = runs.n_runs;
  L->n_pad = round_up(L->n, 128);
  L->keys.alloc(L->n, s);
  SCN_CUDA(cudaMemcpyAsync(L->keys.p, runs.unique.p, sizeof(uint64_t) * L->n, cudaMemcpyDeviceToDevice, s));
  m->levels.push_back(L);

  m->row_of_point.alloc(P, s);
  k_rows_of_points<<<grid_for(P, 256), 256, 0, s>>>(runs.ptr.p, runs.sorted_idx.p, L->n, (int)P, m->row_of_point.p);
  SCN_LAUNCH_CHECK();
  m->rule_ptr.alloc(L->n + 1, s);
  SCN_CUDA(cudaMemcpyAsync(m->rule_ptr.p, runs.ptr.p, sizeof(int) * (L->n + 1), cudaMemcpyDeviceToDevice, s));
  // hand the grouped point list over without a copy
  m->rule_pts.p = runs.sorted_idx.p;
  m->rule_pts.n = runs.sorted_idx.n;
  runs.sorted_idx.p = nullptr;
  runs.sorted_idx.n = 0;

  if (m->point_normals) {          // scn_input_normals(): this batch carries surface normals
    L->normal.alloc((size_t)3 * L->n, s);
    L->ori.alloc((size_t)L->n, s);
    k_voxel_normals<<<grid_for(L->n, 256), 256, 0, s>>>(m->point_normals, m->rule_ptr.p, m->rule_pts.p, L->n, L->normal.p, L->ori.p);
    SCN_LAUNCH_CHECK();
    L->guided = true;
    m->point_normals = nullptr;
  }
  keys.release(s);
  idx.release(s);
  err.release(s);
  dcoords.release(s);
  runs.unique.release(s);
  runs.ptr.release(s);
}

static void build_hash(Level *L, cudaStream_t s) {
  if (L->htab.p) return;
  uint32_t cap = 1024;
  while (cap < 2u * (uint32_t)L->n) cap <<= 1;
  L->hmask = cap - 1;
  L->htab.alloc(cap, s);
  SCN_CUDA(cudaMemsetAsync(L->htab.p, 0xFF, sizeof(ulonglong2) * cap, s));
  if (L->n) {
    k_hash_insert<<<grid_for(L->n, 256), 256, 0, s>>>(L->keys.p, L->n, L->htab.p, L->hmask);
    SCN_LAUNCH_CHECK();
  }
}

// rows per sort block; 0 = natural order.  SCN_TILE_SORT overrides the default at load, scn_tile_sort() at run time.
static std::atomic<int> g_sort_block{[] {
  const char *e = getenv("SCN_TILE_SORT");
  const int b = e ? atoi(e) : SORT_BLOCK_DEFAULT;
  return b < 0 ? 0 : b;
}()};
static int sort_block() {
  const int v = g_sort_block.load(std::memory_order_relaxed);
  return v > 0 ? v : SORT_BLOCK_DEFAULT;
}
bool tile_sort_enabled() { return g_sort_block.load(std::memory_order_relaxed) > 0; }
int set_tile_sort(int block) { return g_sort_block.exchange(block < 0 ? 0 : block); }

__global__ void k_iota(int *p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
__global__ void k_permute_table(const int *__restrict__ nbr, const int *__restrict__ perm, int n, int stride,
                                int *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= stride) return;
  const int r = j < n ? perm[j] : -1;
#pragma unroll
  for (int k = 0; k < 27; ++k) out[(long long)k * stride + j] = r >= 0 ? __ldg(&nbr[(long long)k * stride + r]) : -1;
}

// taps present in every 128-row tile (OR of the row patterns), for the order given by perm (NULL = natural)
__global__ void k_tile_masks(const unsigned long long *__restrict__ row_key, const int *__restrict__ perm, int n,
                             uint32_t *__restrict__ out) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  int r = j < n ? (perm ? perm[j] : j) : -1;
  uint32_t pat = r >= 0 ? pattern_of_key((uint32_t)(row_key[r] & 0x7FFFFFFull)) : 0u;
  pat = __reduce_or_sync(0xffffffffu, pat);
  __shared__ uint32_t acc;
  if (threadIdx.x == 0) acc = 0u;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicOr(&acc, pat);
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// Sort the rows of every block by occupancy pattern (keys from k_neighbours) and gather the table into that order.
void ensure_sorted_table(Level *L, cudaStream_t s) {
  if (L->tile_mask.p || !L->row_key.p || L->n == 0) return;
  const int n = L->n;
  // pattern keys read + sorted (8+4 B in, 4 B out), table gathered into tile order (27 x 4 B read + written), tile masks
  ProfScope ps(PK_RULEBOOK, (tile_sort_enabled() ? 16.0 + 2 * 27.0 * 4 + 8 : 8.0) * (double)n, 0.0, s);
  L->tile_mask.alloc((size_t)(L->n_pad / 128), s);
  if (!tile_sort_enabled()) {          // natural order: only the per-tile tap masks
    k_tile_masks<<<L->n_pad / 128, 128, 0, s>>>(L->row_key.p, nullptr, n, L->tile_mask.p);
    SCN_LAUNCH_CHECK();
    L->row_key.release(s);
    return;
  }
  DevBuf<unsigned long long> keys_out;
  DevBuf<int> idx;
  keys_out.alloc((size_t)n, s);
  idx.alloc((size_t)n, s);
  L->perm.alloc((size_t)L->n_pad, s);
  if (L->n_pad != n) SCN_CUDA(cudaMemsetAsync(L->perm.p + n, 0xFF, sizeof(int) * (size_t)(L->n_pad - n), s));
  k_iota<<<grid_for(n, 256), 256, 0, s>>>(idx.p, n);
  SCN_LAUNCH_CHECK();
  int end_bit = 27;
  while (end_bit < 64 && ((unsigned long long)((n - 1) / sort_block()) >> (end_bit - 27)) != 0) ++end_bit;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, L->row_key.p, keys_out.p, idx.p, L->perm.p, n, 0, end_bit, s);
  DevBuf<uint8_t> tmp;
  tmp.alloc(tb, s);
  SCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, L->row_key.p, keys_out.p, idx.p, L->perm.p, n, 0, end_bit, s));
  count_launch(4);
  L->nbr_sorted.alloc((size_t)27 * L->n_pad, s);
  k_permute_table<<<grid_for(L->n_pad, 256), 256, 0, s>>>(L->nbr.p, L->perm.p, n, L->n_pad, L->nbr_sorted.p);
  SCN_LAUNCH_CHECK();
  k_tile_masks<<<L->n_pad / 128, 128, 0, s>>>(L->row_key.p, L->perm.p, n, L->tile_mask.p);
  SCN_LAUNCH_CHECK();
  L->tile_mask_sorted = true;
  keys_out.release(s);
  idx.release(s);
  tmp.release(s);
  L->row_key.release(s);
}

// Deferred rule counts: valid once the stream has been synchronised after the copies were enqueued.  The pinned slots belong
// to the calling thread and live for the life of the process (cudaHostAlloc / cudaFreeHost synchronise the device: one pair
// per handle cost more than the waits it saved); they are in use only inside one prebuild_scales call at a time.
static unsigned long long *pinned_count_slots() {
  static thread_local unsigned long long *slots = nullptr;
  if (!slots) SCN_CUDA(cudaHostAlloc((void **)&slots, 16 * sizeof(unsigned long long), cudaHostAllocPortable));
  return slots;
}
void resolve_rule_counts(Meta *m, cudaStream_t s, bool synchronise) {
  if (m->pending_counts.empty()) return;
  if (synchronise) SCN_CUDA(cudaStreamSynchronize(s));
  const unsigned long long *slots = pinned_count_slots();
  for (auto &pc : m->pending_counts) pc.first->n_rules = (long long)slots[pc.second];
  m->pending_counts.clear();
}

void ensure_neighbour_table(Meta *m, Level *L, cudaStream_t s, bool defer) {
  if (L->nbr.p) return;
  Level *K = L->base ? L->base : L;       // the scale whose rows and hash this table is built on
  // algorithmic bytes: 27 key probes x 8 B + 12 B of coordinates + 27 x 4 B of table per row, + the hash insert
  ProfScope ps(PK_RULEBOOK, (27.0 * 8 + 12 + 27.0 * 4 + (K->htab.p ? 0.0 : 8 + 12)) * (double)L->n, 0.0, s);
  build_hash(K, s);
  L->nbr.alloc((size_t)27 * L->n_pad, s);
  DevBuf<unsigned long long> cnt;
  cnt.alloc(1, s);
  SCN_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), s));
  if (L->n_pad != L->n)   // padding rows read as "absent" so tile kernels need no tail checks on the table
    for (int k = 0; k < 27; ++k)
      SCN_CUDA(cudaMemsetAsync(L->nbr.p + (size_t)k * L->n_pad + L->n, 0xFF, sizeof(int) * (size_t)(L->n_pad - L->n), s));
  if (L->n) {
    L->row_key.alloc((size_t)L->n, s);
    k_neighbours<<<grid_for(L->n, 256), 256, 0, s>>>(K->keys.p, L->n, L->n_pad, K->htab.p, K->hmask,
                                                     L->nbr.p, cnt.p, L->row_key.p, sort_block(), L->dilation);
    SCN_LAUNCH_CHECK();
  }
  if (defer && m->pending_counts.size() < 16) {
    // the count travels to pinned memory behind the kernel; the caller resolves it after its next synchronisation
    const int slot = (int)m->pending_counts.size();
    SCN_CUDA(cudaMemcpyAsync(pinned_count_slots() + slot, cnt.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    m->pending_counts.emplace_back(L, slot);
    L->n_rules = 0;
    cnt.release(s);
    return;
  }
  unsigned long long h = 0;
  SCN_CUDA(cudaMemcpyAsync(&h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));   // the MAC count is part of the API
  L->n_rules = (long long)h;
  cnt.release(s);
}

void ensure_guided_tables(Meta *m, Level *L, cudaStream_t s) {
  if (L->nbr_g.p || !L->guided) return;
  ensure_neighbour_table(m, L, s);
  const size_t sz = (size_t)27 * L->n_pad;
  ProfScope ps(PK_RULEBOOK, (2.0 * 27 * 4 + 4.0 * 27 * 4 + 1) * (double)L->n, 0.0, s);
  L->nbr_g.alloc(sz, s);
  SCN_CUDA(cudaMemsetAsync(L->nbr_g.p, 0xFF, sizeof(int) * sz, s));
  for (int c = 0; c < 3; ++c) {
    L->nbr_t[c].alloc(sz, s);
    SCN_CUDA(cudaMemsetAsync(L->nbr_t[c].p, 0xFF, sizeof(int) * sz, s));
  }
  if (L->n) {
    k_guide_tables<<<grid_for(L->n, 256), 256, 0, s>>>(L->nbr.p, L->ori.p, L->n, L->n_pad, L->nbr_g.p, L->nbr_t[0].p,
                                                       L->nbr_t[1].p, L->nbr_t[2].p);
    SCN_LAUNCH_CHECK();
  }
}

Level *dilated_level(Meta *m, Level *L, int rate, cudaStream_t s) {
  SCN_CHECK(rate >= 1 && rate < 4096, "SubmanifoldConvolution: bad dilation rate");
  Level *D = L;
  if (rate != 1) {
    D = nullptr;
    for (Level *c : L->dilated)
      if (c->dilation == rate) D = c;
    if (!D) {
      D = new Level();
      for (int d = 0; d < 3; ++d) D->size[d] = L->size[d];
      D->n = L->n;
      D->n_pad = L->n_pad;
      D->base = L;
      D->dilation = rate;
      L->dilated.push_back(D);
    }
  }
  ensure_neighbour_table(m, D, s);
  return D;
}

Level *ensure_coarse_level(Meta *m, Level *F, const int64_t coarse_size[3], cudaStream_t s) {
  if (F->coarse) {
    SCN_CHECK(F->coarse->size[0] == coarse_size[0], "Convolution: a different output size was already linked to this scale");
    return F->coarse;
  }
  for (int d = 0; d < 3; ++d)
    SCN_CHECK((coarse_size[d] - 1) * 2 + 2 == F->size[d], "Convolution: only filter size 2 / stride 2 is supported (reference FastDownSampleMode)");
  SCN_CHECK(find_level(m, coarse_size) == nullptr, "Convolution: output scale already exists in this handle");
  // fine keys read (8 B), coarse keys sorted with their fine row (12 B in, 12 B out), parent + offset written (5 B); child table below
  ProfScope ps(PK_RULEBOOK, (8.0 + 24 + 5 + 4) * (double)F->n, 0.0, s);
  DevBuf<uint64_t> ckeys;
  DevBuf<int> idx;
  ckeys.alloc(F->n, s);
  idx.alloc(F->n, s);
  k_coarse_keys<<<grid_for(F->n, 256), 256, 0, s>>>(F->keys.p, F->n, ckeys.p, idx.p);
  SCN_LAUNCH_CHECK();
  Runs runs;
  sort_and_group(ckeys, idx, F->n, 48 + bits_for((unsigned)(m->batch > 1 ? m->batch - 1 : 1)), runs, s);

  Level *C = new Level();
  for (int d = 0; d < 3; ++d) C->size[d] = coarse_size[d];
  C->n = runs.n_runs;
  C->n_pad = round_up(C->n, 128);
  C->keys.alloc(C->n, s);
  SCN_CUDA(cudaMemcpyAsync(C->keys.p, runs.unique.p, sizeof(uint64_t) * C->n, cudaMemcpyDeviceToDevice, s));
  F->parent.alloc(F->n, s);
  F->off8.alloc(F->n, s);
  F->child.alloc((size_t)8 * C->n_pad, s);
  SCN_CUDA(cudaMemsetAsync(F->child.p, 0xFF, sizeof(int) * (size_t)8 * C->n_pad, s));      // -1 = no child at this offset
  k_link_levels<<<grid_for(F->n, 256), 256, 0, s>>>(runs.ptr.p, runs.sorted_idx.p, F->keys.p, C->n, F->n, C->n_pad,
                                                    F->parent.p, F->off8.p, F->child.p);
  SCN_LAUNCH_CHECK();
  if (F->guided && F->size[0] >= m->normal_guide_scale) {     // ConvolutionRules.h:774: below the guide scale the plain rules are used
    C->normal.alloc((size_t)3 * C->n, s);
    C->ori.alloc((size_t)C->n, s);
    k_coarse_normals_and_taps<<<grid_for(C->n, 256), 256, 0, s>>>(F->normal.p, C->n, C->n_pad, F->child.p, F->off8.p, C->normal.p,
                                                                C->ori.p);
    SCN_LAUNCH_CHECK();
    C->guided = true;
  }
  F->coarse = C;
  m->levels.push_back(C);
  ckeys.release(s);
  idx.release(s);
  runs.unique.release(s);
  runs.ptr.release(s);
  runs.sorted_idx.release(s);
  return C;
}

}  // namespace scn

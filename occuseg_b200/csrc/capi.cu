// extern "C" boundary of libscn_b200.so -- see include/scn_b200.h for the contract and the reference
// interface each entry replaces.
#include "../../include/scn_b200.h"
#include "common.cuh"
#include <mutex>
#include <cstdlib>
#include <cstdio>

namespace scn {

static thread_local std::string t_last_error;
thread_local cudaStream_t t_last_stream = nullptr;
void set_last_error(const std::string &s) { t_last_error = s; }
std::atomic<long long> g_launches{0};

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

int sm_count() {
  static std::atomic<int> cache[MAX_DEVICES];
  const int dev = current_device();
  int n = (dev >= 0 && dev < MAX_DEVICES) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev);
    if (n <= 0) n = 148;
    // SCN_SM_RESERVE: SMs left to other resident kernels (e.g. a communication library's); the persistent kernels size their
    // grids to the rest, so none of their CTAs has to wait for an SM
    if (const char *e = getenv("SCN_SM_RESERVE")) n = n - atoi(e) > 8 ? n - atoi(e) : n;
    if (dev >= 0 && dev < MAX_DEVICES) cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---- profiling ------------------------------------------------------------------------------------------
struct ProfRec { int kind; cudaEvent_t e0, e1; double bytes, flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static double g_prof_acc[PK_COUNT][4];
static cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
ProfScope::ProfScope(ProfKind kind, double bytes, double flops, cudaStream_t stream) : s(stream) {
  if (!g_prof_on) return;
  ProfRec r{(int)kind, prof_event(), prof_event(), bytes, flops};
  cudaEventRecord(r.e0, s);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, s);
}
static void prof_drain() {
  static const bool log_each = getenv("SCN_PROF_LOG") != nullptr;      // one line per instrumented launch (debugging aid)
  for (ProfRec &r : g_prof) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (log_each) fprintf(stderr, "[scn prof] %-10s %8.4f ms  %10.1f MB  %9.2f GFLOP\n", prof_name(r.kind), ms, r.bytes / 1e6, r.flops / 1e9);
    g_prof_acc[r.kind][0] += 1;
    g_prof_acc[r.kind][1] += ms;
    g_prof_acc[r.kind][2] += r.bytes;
    g_prof_acc[r.kind][3] += r.flops;
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
}
void prof_enable(bool on) {
  prof_drain();
  if (on) for (auto &k : g_prof_acc) for (double &v : k) v = 0;
  g_prof_on = on;
}
int prof_read(double *out, int max_kinds) {
  prof_drain();
  int n = max_kinds < PK_COUNT ? max_kinds : PK_COUNT;
  for (int k = 0; k < n; ++k) for (int j = 0; j < 4; ++j) out[k * 4 + j] = g_prof_acc[k][j];
  return n;
}
const char *prof_name(int kind) {
  static const char *names[PK_COUNT] = {"rulebook", "conv_tc", "conv_fp32", "wgrad_tc", "wgrad_fp32", "bn", "io", "cast"};
  return (kind >= 0 && kind < PK_COUNT) ? names[kind] : "";
}

static Level *need_level(Meta *m, const int64_t size[3], const char *what) {
  m->last_stream = t_last_stream;       // every entry that names a scale passes through here after note_stream()
  Level *L = find_level(m, size);
  if (!L) throw Error(std::string(what) + ": no such scale in this handle (size " + std::to_string(size[0]) + ")");
  return L;
}

// the (possibly dilated) neighbourhood a submanifold entry works on: consumes the one-use hint of scn_subm_dilation
static Level *subm_level(Meta *m, const int64_t size[3], cudaStream_t s) {
  Level *L = need_level(m, size, "SubmanifoldConvolution");
  const int rate = m->next_dilation;
  m->next_dilation = 1;
  return dilated_level(m, L, rate, s);
}

// ---- precision plumbing ----------------------------------------------------------------------------------
// SCN_BF16: operands of the tensor-core kernels are bf16 COPIES of the caller's fp32 matrices (made here, one
// streaming pass each, or handed in by the caller when two products share one), products accumulate in fp32 and
// outputs stay fp32.  Needs K % 64 == 0 (128-byte swizzle rows of bf16); other shapes fall to SCN_TF32 tiles
// (K % 32 == 0) and then to the exact fp32 kernels.
static bool bf16_conv_shape(int k, int n, int precision) { return precision == SCN_BF16 && k % 64 == 0 && n % 32 == 0; }
static bool bf16_wgrad_shape(int cg, int cs, int precision) { return precision == SCN_BF16 && cg % 64 == 0 && cs % 64 == 0; }

struct Bf16Copy {
  DevBuf<uint16_t> buf;
  const uint16_t *make(const float *src, long long n, cudaStream_t s) {
    buf.alloc((size_t)n, s);
    ProfScope ps(PK_CAST, 6.0 * (double)n, 0.0, s);
    cast_bf16(src, buf.p, n, s);
    return buf.p;
  }
  // from rows `ld` floats apart (0 = dense)
  const uint16_t *make_rows(const float *src, long long ld, long long rows, int cols, cudaStream_t s) {
    buf.alloc((size_t)rows * cols, s);
    ProfScope ps(PK_CAST, 6.0 * (double)rows * cols, 0.0, s);
    cast_bf16_rows(src, ld, rows, cols, buf.p, s);
    return buf.p;
  }
  void release(cudaStream_t s) { buf.release(s); }
};

// row stride of d_out registered for this backward entry (scn_grad_stride); 0 = dense.  A strided d_out is only ever read by
// the bf16 cast, so every product of the entry must run on bf16 copies (and there must be no bias gradient).
static long long take_grad_ld(Meta *m, int cols, bool all_bf16, const float *d_bias) {
  long long ld = m->next_grad_ld;
  m->next_grad_ld = 0;
  if (ld == cols) ld = 0;
  SCN_CHECK(ld == 0 || (all_bf16 && !d_bias && ld > cols),
            "row-strided d_out needs the bf16 tensor-core path for every product of this entry (scn_grad_stride)");
  return ld;
}

// `w` is the caller's weight array; native_kn says whether, for THIS product, it already reads as
// [V][K=c_in][N=c_out] (true) or as [V][N][K] (false).  The fp32 kernels want KN, the tensor-core kernel
// wants NK (K-major B operand); whichever is missing is produced by one small per-tap transpose.
// in16: optional bf16 copy of a.in made by the caller.
static void run_conv(ConvArgs a, const float *w, bool native_kn, int precision, cudaStream_t s,
                     const uint16_t *in16 = nullptr) {
  a.bf16 = bf16_conv_shape(a.c_in, a.c_out, precision);
  const bool tcore = precision != SCN_FP32 && conv_tma_supported(a);
  if (!tcore) a.bf16 = false;
  const bool want_kn = !tcore;
  DevBuf<float> tmp;
  DevBuf<uint16_t> w16;
  Bf16Copy x16;
  const float *use = w;
  const int wv = a.n_taps ? a.n_taps : a.V;        // weight taps (the one-tap-per-row form has V = 1 table row, n_taps weights)
  const size_t wn = (size_t)wv * a.c_in * a.c_out;
  const bool fused_prep = tcore && a.bf16 && want_kn != native_kn;    // bf16 tiles: transpose + rounding in one kernel
  if (want_kn != native_kn && !fused_prep) {
    tmp.alloc(wn, s);
    // source rows/cols: native_kn -> [c_in][c_out], else [c_out][c_in]
    if (native_kn) transpose_weight(w, tmp.p, wv, a.c_in, a.c_out, s);
    else transpose_weight(w, tmp.p, wv, a.c_out, a.c_in, s);
    use = tmp.p;
  }
  const double es = a.bf16 ? 2.0 : 4.0;      // bytes per gathered element
  // algorithmic work (gather/scatter model): R*Cin*s + N*Cout*4 + 4*R + V*Cin*Cout*s
  // + the per-row fp32 operand an epilogue fusion reads in the same pass (residual shortcut, or the BatchNorm input of a fused
  //   BatchNorm backward): N*Cout*4 -- algorithmic traffic of the fused elementwise layer, which no longer has a pass of its own
  const double out_rows = (double)(a.scatter || a.item_off ? a.n_rules : a.n_rows);    // one-rule-per-row forms write n_rules rows
  const double bytes = es * (double)a.n_rules * a.c_in + 4.0 * out_rows * a.c_out + 4.0 * (double)a.n_rules +
                       es * (double)wv * a.c_in * a.c_out + ((a.residual || a.bnb_x) ? 4.0 * out_rows * a.c_out : 0.0) +
                       (a.out_bf16 ? 2.0 * out_rows * a.c_out : 0.0);
  const double flops = 2.0 * (double)a.n_rules * a.c_in * a.c_out;
  if (a.bf16) {
    w16.alloc(wn, s);
    if (fused_prep) transpose_weight_bf16(w, w16.p, wv, a.c_in, a.c_out, s);      // native [K][N] -> [N][K] bf16
    else cast_bf16(use, w16.p, (long long)wn, s);
    if (!in16) in16 = x16.make(a.in, (long long)a.in_rows * a.c_in, s);
  }
  {
    ProfScope ps(tcore ? PK_CONV_TC : PK_CONV_FP32, bytes, flops, s);
    if (tcore) {
      a.weight_nk = a.bf16 ? (const void *)w16.p : (const void *)use;
      if (a.bf16) a.in = reinterpret_cast<const float *>(in16);
      conv_tma(a, s);
    } else {
      a.weight = use;
      if (conv_small_supported(a)) conv_small(a, s);
      else conv_simt(a, s);
    }
  }
  tmp.release(s);
  w16.release(s);
  x16.release(s);
}

// One-rule-per-fine-row products (Deconvolution forward, strided-Convolution dgrad):
//   out[i] = in[parent[i]] * Wk(off[i]).   fp32: input-stationary scatter over the child table;
//   tensor cores: the fine rows regrouped by tap (see below).
// The BatchNorm-backward hint of a handle (scn_bn_bwd_fusion) applied to the dgrad product `a` of the next backward entry:
// returns the scratch buffer holding the mask coefficients (released by the caller after the launch).  Throws when the
// product cannot honour it (not on the tensor-core path) -- callers ask scn_bn_bwd_fusable first.
struct BnbScratch {
  DevBuf<float> coef;
  void release(cudaStream_t s) { coef.release(s); }
};
static void apply_bnb_hint(Meta *m, ConvArgs &a, int precision, BnbScratch &scr, cudaStream_t s) {
  Meta::BnBwdHint h = m->bnb;
  m->bnb = Meta::BnBwdHint();
  if (!h.x) return;
  ConvArgs probe = a;
  probe.bf16 = bf16_conv_shape(a.c_in, a.c_out, precision);
  if (probe.n_rows == 0) probe.n_rows = 1;
  SCN_CHECK(precision != SCN_FP32 && conv_tma_supported(probe) && a.out != nullptr && (uintptr_t)h.x % 16 == 0,
            "fused BatchNorm backward: this layer's dgrad does not run on the tensor-core path (see scn_bn_bwd_fusable)");
  (void)scr;
  a.bnb_x = h.x;
  a.bnb_mean = h.mean; a.bnb_invstd = h.invstd; a.bnb_gamma = h.gamma; a.bnb_beta = h.beta;
  a.bnb_leak = h.leak;
  a.stats = h.acc;
}

static void run_up(Level *F, Level *C, const float *in, const float *w, bool native_kn, float *out, int c_in, int c_out,
                   int precision, cudaStream_t s, const uint16_t *in16 = nullptr, Meta *bnb_from = nullptr,
                   double *stats = nullptr) {
  // tensor cores: the fine rows grouped by tap (whole 256-row tile groups per tap), one gathered row and one weight tap
  // per row -- no work is spent on the 7 taps a row does not have
  ConvArgs g;
  g.in = in; g.out = out; g.V = 1; g.c_in = c_in; g.c_out = c_out; g.n_rules = F->n; g.in_rows = C->n;
  g.n_rows = 1;   // placeholder for the support check; set below
  g.bf16 = bf16_conv_shape(c_in, c_out, precision);
  if (precision != SCN_FP32 && conv_tma_supported(g)) {
    build_pair_list(F->up_pairs, F->child.p, 8, C->n_pad, F->n, 256, 0xFF, s);
    PairList &P = F->up_pairs;
    g.tbl = P.si.p;                     // gathered (coarse) row of every rule
    g.tbl_stride = (int)(P.n_items_ub * 256);
    g.n_rows = (int)(P.n_items_ub * 256);
    g.out_rows = P.gi.p;                // fine row every result goes to
    g.item_off = P.item_off.p;
    g.rows_per_item = 256;
    g.n_taps = 8;
    g.out_limit = F->n;
    BnbScratch scr;
    if (bnb_from) apply_bnb_hint(bnb_from, g, precision, scr, s);
    if (stats) g.stats = stats;
    run_conv(g, w, native_kn, precision, s, in16);
    scr.release(s);
    return;
  }
  SCN_CHECK(!stats, "column statistics need the tensor-core path");
  SCN_CHECK(!bnb_from || !bnb_from->bnb.x, "fused BatchNorm backward: this layer's dgrad does not run on the tensor-core path");
  ConvArgs a;
  a.in = in; a.out = out; a.tbl = F->child.p; a.tbl_stride = C->n_pad; a.n_rows = C->n; a.V = 8;
  a.c_in = c_in; a.c_out = c_out; a.scatter = true; a.n_rules = F->n;
  run_conv(a, w, native_kn, SCN_FP32, s);
}

// Tensor-core submanifold products walk the level in its pattern-sorted tile order (Level::perm / nbr_sorted) and write
// their result rows through the permutation; the exact-fp32 kernels keep the natural order.
static void use_sorted_tiles(ConvArgs &a, Level *L, int precision, cudaStream_t s) {
  if (precision == SCN_FP32) return;
  ConvArgs probe = a;
  probe.bf16 = bf16_conv_shape(a.c_in, a.c_out, precision);
  if (!conv_tma_supported(probe)) return;
  ensure_sorted_table(L, s);
  a.tile_mask = L->tile_mask.p;         // taps with no row in a tile group never enter the pipeline
  if (!L->tile_mask_sorted) return;     // natural order
  a.tbl = L->nbr_sorted.p;
  a.out_rows = L->perm.p;
  a.out_limit = L->n;
}

// a16 / b16: optional bf16 copies of a.a / a.b made by the caller
static void run_wgrad(WgradArgs a, PairList &pairs, int precision, cudaStream_t s, const uint16_t *a16 = nullptr,
                      const uint16_t *b16 = nullptr) {
  const int cg = a.table_on_a ? a.c_a : a.c_b, cs = a.table_on_a ? a.c_b : a.c_a;
  a.bf16 = bf16_wgrad_shape(cg, cs, precision);
  if (precision != SCN_FP32) {
    build_pair_list(pairs, a.tbl, a.V, a.tbl_stride, a.n_rules, PAIR_ITEM, 0x7F, s);
    a.gi = pairs.gi.p; a.si = pairs.si.p; a.blk_item = pairs.blk_item.p; a.n_blk = pairs.n_blk; a.blk_rows = BLK_ROWS;
  }
  const bool tcore = precision != SCN_FP32 && wgrad_tma_supported(a);
  if (!tcore) a.bf16 = false;
  Bf16Copy ca, cb;
  if (a.bf16) {
    const long long rows_a = a.table_on_a ? a.g_rows : a.s_rows, rows_b = a.table_on_a ? a.s_rows : a.g_rows;
    if (!a16) a16 = ca.make(a.a, rows_a * a.c_a, s);
    if (!b16) b16 = cb.make(a.b, rows_b * a.c_b, s);
    a.a = reinterpret_cast<const float *>(a16);
    a.b = reinterpret_cast<const float *>(b16);
  }
  const double es = a.bf16 ? 2.0 : 4.0;
  // R*(Cin+Cout)*s + 8*R + V*Cin*Cout*4
  const double bytes = es * (double)a.n_rules * (a.c_a + a.c_b) + 8.0 * a.n_rules + 4.0 * (double)a.V * a.c_a * a.c_b;
  {
    ProfScope ps(tcore ? PK_WGRAD_TC : PK_WGRAD_FP32, bytes, 2.0 * (double)a.n_rules * a.c_a * a.c_b, s);
    if (tcore) wgrad_tma(a, s);
    else if (wgrad_small_supported(a)) wgrad_small(a, s);
    else wgrad_simt(a, s);
  }
  ca.release(s);
  cb.release(s);
}

}  // namespace scn

using namespace scn;

#define SCN_TRY try {
#define SCN_CATCH                                                                     \
  }                                                                                   \
  catch (const std::exception &e) {                                                   \
    set_last_error(e.what());                                                         \
    return 1;                                                                         \
  }                                                                                   \
  catch (...) {                                                                       \
    set_last_error("unknown error");                                                  \
    return 1;                                                                         \
  }                                                                                   \
  return 0;

struct scn_meta {
  Meta m;
};

// the bf16 copy registered for `src` (filled here if the caller only supplied the buffer), or NULL; one use only
static const uint16_t *take_bf16_hint(Meta *m, const float *src, long long n, cudaStream_t s) {
  if (!m->hint_bf16 || m->hint_src != src) return nullptr;
  uint16_t *p = (uint16_t *)m->hint_bf16;
  if (!m->hint_ready) {
    ProfScope ps(PK_CAST, 6.0 * (double)n, 0.0, s);
    cast_bf16(src, p, n, s);
  }
  m->hint_bf16 = nullptr;
  m->hint_src = nullptr;
  return p;
}

// bf16 copy of d_out for the backward products: the one the producer of d_out left (scn_grad_bf16; dense rows only), else a cast pass
// (the hint is consumed by every backward entry, wanted or not: a stale one could match a later d_out at a recycled address)
static const uint16_t *grad_bf16(Meta *m, bool wanted, Bf16Copy &g16, const float *d_out, long long ld, long long rows, int cols,
                                 cudaStream_t s) {
  const uint16_t *held = (m->ghint_src == d_out && ld == 0) ? (const uint16_t *)m->ghint_bf16 : nullptr;
  m->ghint_src = nullptr;
  m->ghint_bf16 = nullptr;
  if (!wanted) return nullptr;
  return held ? held : g16.make_rows(d_out, ld, rows, cols, s);
}

static void check_channels(int c_in, int c_out) {
  SCN_CHECK(c_in > 0 && c_out > 0 && c_in <= 4096 && c_out <= 4096, "channel counts out of range");
}

extern "C" {

int scn_version(void) { return 100; }
const char *scn_last_error(void) { return t_last_error.c_str(); }
int64_t scn_launch_count(void) { return (int64_t)g_launches.load(); }
void scn_profile(int enable) { prof_enable(enable != 0); }
int scn_profile_read(double *out, int max_kinds) { return prof_read(out, max_kinds); }
const char *scn_profile_kind_name(int kind) { return prof_name(kind); }

scn_meta *scn_meta_create(int device) {
  try {
    SCN_CUDA(cudaSetDevice(device));
    {
      // keep freed blocks in the stream-ordered pool across synchronisation points; with the default threshold (0)
      // every sync hands the memory back to the driver and the next batch pays for cudaMalloc.  The amount kept is
      // BOUNDED (SCN_POOL_KEEP_MB, default 4096: a step of the 8 x 250k-voxel workload holds < 2 GB of rulebooks and
      // operand copies) so the embedding process -- PyTorch's caching allocator cannot see this pool -- gets the rest
      // back at the next synchronisation; scn_pool_trim() returns everything that is free right now.
      static std::atomic<bool> pool_ready[MAX_DEVICES];
      if (device >= 0 && device < MAX_DEVICES && !pool_ready[device].load()) {
        cudaMemPool_t pool;
        SCN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        const char *e = getenv("SCN_POOL_KEEP_MB");
        uint64_t keep = (uint64_t)(e ? atoll(e) : 4096) << 20;
        uint64_t cur = 0;
        SCN_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur));
        if (cur < keep) SCN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));   // never lower someone else's setting
        pool_ready[device].store(true);
      }
    }
    scn_meta *h = new scn_meta();
    h->m.device = device;
    return h;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return nullptr;
  }
}

int scn_tile_sort(int block_rows) { return set_tile_sort(block_rows); }
int scn_deterministic(int on) { return set_deterministic(on); }

int scn_pool_trim(int device, int64_t keep_bytes) {
  SCN_TRY
  cudaMemPool_t pool;
  SCN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  SCN_CUDA(cudaMemPoolTrimTo(pool, keep_bytes < 0 ? 0 : (size_t)keep_bytes));
  SCN_CATCH
}

void scn_meta_destroy(scn_meta *h) {
  if (!h) return;
  cudaSetDevice(h->m.device);
  t_last_stream = h->m.last_stream;     // the handle's buffers are freed in stream order behind its last kernels
  delete h;
}

// Depth of the scale hierarchy the previous batches ended up with.  Every scale costs two host synchronisations
// (row count, rule count); done lazily they sit between the layers of the forward pass and drain the launch queue
// right before the small deep levels, which then run launch-bound.  So once the depth is known, the whole chain of
// rulebooks is built (and all its synchronisations paid) inside the InputLayer call, and the rest of the step is
// enqueued without a single host wait.  A batch that needs fewer scales wastes a few tiny kernels; one that needs
// more builds the extra scales lazily, as before, and raises the hint.
static std::atomic<int> g_depth_hint{0};

static void prebuild_scales(Meta *m, cudaStream_t s) {
  const int depth = g_depth_hint.load();
  // rule counts are not waited for one by one: each copy is resolved after the synchronisation of the NEXT coarse scale's row
  // count (which the chain needs anyway), the last one by one final wait -- 7 host synchronisations per batch instead of 12
  struct Resolve {
    Meta *m; cudaStream_t s;
    ~Resolve() { try { resolve_rule_counts(m, s, true); } catch (...) { m->pending_counts.clear(); } }
  } resolve_at_exit{m, s};
  for (int i = 0; i + 1 <= depth; ++i) {
    Level *L = m->levels.back();
    ensure_neighbour_table(m, L, s, true);
    if (i + 1 == depth) break;
    int64_t coarse[3];
    bool ok = L->n > 0;
    for (int d = 0; d < 3; ++d) {
      ok = ok && L->size[d] % 2 == 0 && L->size[d] >= 2;
      coarse[d] = L->size[d] / 2;
    }
    if (!ok) break;
    ensure_coarse_level(m, L, coarse, s);
    resolve_rule_counts(m, s, false);        // ensure_coarse_level has just synchronised the stream
  }
}

int scn_input_normals(scn_meta *h, const float *point_normals, int normal_guide_scale) {
  SCN_TRY
  SCN_CHECK(h && point_normals, "scn_input_normals: null argument");
  h->m.point_normals = point_normals;
  h->m.normal_guide_scale = normal_guide_scale > 0 ? normal_guide_scale : (1 << 30);
  SCN_CATCH
}

int scn_guided(scn_meta *h, const int64_t size[3]) {
  if (!h) return 0;
  Level *L = find_level(&h->m, size);
  return L && L->guided ? 1 : 0;
}

int scn_subm_guided_table(scn_meta *h, const int64_t size[3], void *stream, int32_t *out, uint8_t *ori_out) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  Level *L = need_level(&h->m, size, "SubmanifoldConvolution");
  SCN_CHECK(L->guided, "this scale carries no normals (scn_input_normals)");
  ensure_guided_tables(&h->m, L, s);
  SCN_CUDA(cudaMemcpy2DAsync(out, sizeof(int) * L->n, L->nbr_g.p, sizeof(int) * L->n_pad, sizeof(int) * L->n, 27,
                             cudaMemcpyDeviceToHost, s));
  if (ori_out) SCN_CUDA(cudaMemcpyAsync(ori_out, L->ori.p, L->n, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  SCN_CATCH
}

int scn_normals(scn_meta *h, const int64_t size[3], void *stream, float *out) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  Level *L = need_level(&h->m, size, "normals");
  SCN_CHECK(L->guided && out, "this scale carries no normals (scn_input_normals)");
  SCN_CUDA(cudaMemcpyAsync(out, L->normal.p, sizeof(float) * 3 * L->n, cudaMemcpyDeviceToHost, s));
  SCN_CUDA(cudaStreamSynchronize(s));
  SCN_CATCH
}

int scn_input_layer_build(scn_meta *h, const int64_t size[3], const int64_t *coords, int on_device, int64_t P,
                          int batch, int mode, void *stream, int64_t *n_active) {
  SCN_TRY
  SCN_CHECK(h && coords && n_active, "null argument");
  h->m.last_stream = note_stream(stream);
  build_input_level(&h->m, size, coords, on_device != 0, P, batch, mode, note_stream(stream));
  *n_active = h->m.levels[0]->n;
  prebuild_scales(&h->m, note_stream(stream));
  SCN_CATCH
}

int scn_input_layer_fwd(scn_meta *h, const float *feats, int C, float *out, void *stream) {
  SCN_TRY
  const double n0 = h->m.levels.empty() ? 0.0 : (double)h->m.levels[0]->n;
  ProfScope ps(PK_IO, 4.0 * C * ((double)h->m.n_points + n0) + 4.0 * (double)h->m.n_points + 4.0 * n0, 0.0, note_stream(stream));
  input_layer_fwd(&h->m, feats, C, out, note_stream(stream));
  SCN_CATCH
}
int scn_input_layer_bwd(scn_meta *h, const float *d_out, int C, float *d_feats, void *stream) {
  SCN_TRY
  ProfScope ps(PK_IO, 8.0 * C * (double)h->m.n_points + 4.0 * (double)h->m.n_points, 0.0, note_stream(stream));
  input_layer_bwd(&h->m, d_out, C, d_feats, note_stream(stream));
  SCN_CATCH
}
int scn_output_layer_fwd(scn_meta *h, const float *in, int C, float *out, void *stream) {
  SCN_TRY
  ProfScope ps(PK_IO, 8.0 * C * (double)h->m.n_points + 4.0 * (double)h->m.n_points, 0.0, note_stream(stream));
  output_layer_fwd(&h->m, in, C, out, note_stream(stream));
  SCN_CATCH
}
int scn_output_layer_bwd(scn_meta *h, const float *d_out, int C, float *d_in, void *stream) {
  SCN_TRY
  const double n0 = h->m.levels.empty() ? 0.0 : (double)h->m.levels[0]->n;
  ProfScope ps(PK_IO, 4.0 * C * ((double)h->m.n_points + n0) + 4.0 * (double)h->m.n_points + 4.0 * n0, 0.0, note_stream(stream));
  output_layer_bwd(&h->m, d_out, C, d_in, note_stream(stream));
  SCN_CATCH
}
int scn_float_coords(const float *xyz, int64_t n_points, const float offset[3], int batch_index, float full_scale,
                     int64_t *coords, uint8_t *keep, void *stream) {
  SCN_TRY
  SCN_CHECK(xyz && offset && coords, "scn_float_coords: null argument");
  ProfScope ps(PK_IO, (12.0 + 32 + 1) * (double)n_points, 0.0, note_stream(stream));
  float_coords(xyz, n_points, offset, batch_index, full_scale, (long long *)coords, keep, note_stream(stream));
  SCN_CATCH
}

int64_t scn_n_points(scn_meta *h) { return h ? h->m.n_points : -1; }

int64_t scn_nactive(scn_meta *h, const int64_t size[3]) {
  if (!h) return -1;
  Level *L = find_level(&h->m, size);
  return L ? L->n : -1;
}

int scn_spatial_locations(scn_meta *h, const int64_t size[3], int64_t *out) {
  SCN_TRY
  Level *L = need_level(&h->m, size, "getSpatialLocations");
  std::vector<uint64_t> keys(L->n);
  // on the stream the handle last worked on (the caller's, possibly non-blocking): ordered behind the kernels that built the keys
  SCN_CUDA(cudaMemcpyAsync(keys.data(), L->keys.p, sizeof(uint64_t) * L->n, cudaMemcpyDeviceToHost, h->m.last_stream));
  SCN_CUDA(cudaStreamSynchronize(h->m.last_stream));
  for (int i = 0; i < L->n; ++i) {
    uint64_t k = keys[i];
    out[4 * i + 0] = (int64_t)(k & 0xFFFF);
    out[4 * i + 1] = (int64_t)((k >> 16) & 0xFFFF);
    out[4 * i + 2] = (int64_t)((k >> 32) & 0xFFFF);
    out[4 * i + 3] = (int64_t)(k >> 48);
  }
  SCN_CATCH
}

int scn_subm_dilation(scn_meta *h, int rate) {
  SCN_TRY
  SCN_CHECK(h && rate >= 1 && rate < 4096, "scn_subm_dilation: bad dilation rate");
  h->m.next_dilation = rate;
  SCN_CATCH
}

int scn_subm_rulebook(scn_meta *h, const int64_t size[3], void *stream, int64_t *n_rules) {
  SCN_TRY
  Level *L = subm_level(&h->m, size, note_stream(stream));
  if (n_rules) *n_rules = L->n_rules;
  SCN_CATCH
}

int scn_subm_neighbour_table(scn_meta *h, const int64_t size[3], int32_t *out) {
  SCN_TRY
  Level *L = need_level(&h->m, size, "SubmanifoldConvolution");
  if (h->m.next_dilation != 1) {                 // the table of a dilated neighbourhood (must have been built)
    Level *D = nullptr;
    for (Level *c : L->dilated)
      if (c->dilation == h->m.next_dilation) D = c;
    h->m.next_dilation = 1;
    SCN_CHECK(D, "dilated neighbour table not built yet (call scn_subm_dilation + scn_subm_rulebook)");
    L = D;
  }
  SCN_CHECK(L->nbr.p, "neighbour table not built yet (call scn_subm_rulebook)");
  SCN_CUDA(cudaMemcpy2DAsync(out, sizeof(int) * L->n, L->nbr.p, sizeof(int) * L->n_pad, sizeof(int) * L->n, 27,
                             cudaMemcpyDeviceToHost, h->m.last_stream));
  SCN_CUDA(cudaStreamSynchronize(h->m.last_stream));
  SCN_CATCH
}

static void note_depth(Meta *m) {
  int d = (int)m->levels.size(), cur = g_depth_hint.load();
  while (d > cur && !g_depth_hint.compare_exchange_weak(cur, d)) {}
}

int scn_strided_rulebook(scn_meta *h, const int64_t fine[3], const int64_t coarse[3], void *stream, int64_t *n_coarse) {
  SCN_TRY
  Level *F = need_level(&h->m, fine, "Convolution");
  Level *C = ensure_coarse_level(&h->m, F, coarse, note_stream(stream));
  note_depth(&h->m);
  if (n_coarse) *n_coarse = C->n;
  SCN_CATCH
}

int scn_strided_table(scn_meta *h, const int64_t fine[3], int32_t *parent, uint8_t *off) {
  SCN_TRY
  Level *F = need_level(&h->m, fine, "Convolution");
  SCN_CHECK(F->coarse, "strided rulebook not built yet (call scn_strided_rulebook)");
  SCN_CUDA(cudaMemcpyAsync(parent, F->parent.p, sizeof(int) * F->n, cudaMemcpyDeviceToHost, h->m.last_stream));
  SCN_CUDA(cudaMemcpyAsync(off, F->off8.p, F->n, cudaMemcpyDeviceToHost, h->m.last_stream));
  SCN_CUDA(cudaStreamSynchronize(h->m.last_stream));
  SCN_CATCH
}

int scn_resolution_scatter(const int32_t *lr_xyz, int64_t n_lr, const int32_t *hr_xyz, int64_t n_hr, int stride,
                           int32_t *hr2lr, void *stream) {
  SCN_TRY
  ProfScope ps(PK_RULEBOOK, 12.0 * (double)(n_lr + n_hr) + 4.0 * (double)n_hr, 0.0, note_stream(stream));
  resolution_scatter(lr_xyz, n_lr, hr_xyz, n_hr, stride, hr2lr, note_stream(stream));
  SCN_CATCH
}

// ---- submanifold ------------------------------------------------------------------------------------
int scn_bn_eval_coeffs(const float *running_mean, const float *running_var, const float *gamma, const float *beta,
                       int channels, float eps, float *scale, float *shift, void *stream) {
  SCN_TRY
  bn_eval_coeffs(running_mean, running_var, gamma, beta, channels, eps, scale, shift, note_stream(stream));
  SCN_CATCH
}

int scn_subm_fwd_bn(scn_meta *h, const int64_t size[3], const float *in, const float *weight, const float *bias,
                    const float *residual, const float *bn_scale, const float *bn_shift, float leakiness, float *out,
                    void *out_bf16, int c_in, int c_out, int precision, void *stream, double *macs) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  SCN_CHECK(bn_scale && bn_shift, "scn_subm_fwd_bn: null coefficients");
  Level *L = subm_level(&h->m, size, s);
  ConvArgs a;
  a.in = in; a.bias = bias; a.out = out;
  a.tbl = L->nbr.p; a.tbl_stride = L->n_pad; a.n_rows = L->n; a.V = 27; a.c_in = c_in; a.c_out = c_out; a.n_rules = L->n_rules; a.in_rows = L->n;
  a.residual = residual; a.ep_scale = bn_scale; a.ep_shift = bn_shift; a.ep_leak = leakiness; a.out_bf16 = (uint16_t *)out_bf16;
  ConvArgs probe = a;
  probe.bf16 = bf16_conv_shape(c_in, c_out, precision);
  SCN_CHECK(precision != SCN_FP32 && conv_tma_supported(probe), "scn_subm_fwd_bn needs the tensor-core path (see scn_fuses_residual)");
  if (L->guided) {
    ensure_guided_tables(&h->m, L, s);
    a.tbl = L->nbr_g.p;
  } else {
    use_sorted_tiles(a, L, precision, s);
  }
  run_conv(a, weight, true, precision, s, take_bf16_hint(&h->m, in, (long long)L->n * c_in, s));
  if (macs) *macs = (double)L->n_rules * c_in * c_out;
  SCN_CATCH
}

int scn_subm_fwd(scn_meta *h, const int64_t size[3], const float *in, const float *weight, const float *bias,
                 const float *residual, double *stats, float *out, int c_in, int c_out, int precision, void *stream,
                 double *macs) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  Level *L = subm_level(&h->m, size, s);
  ConvArgs a;
  a.in = in; a.bias = bias; a.out = out;
  a.tbl = L->nbr.p; a.tbl_stride = L->n_pad; a.n_rows = L->n; a.V = 27; a.c_in = c_in; a.c_out = c_out; a.n_rules = L->n_rules; a.in_rows = L->n;
  a.residual = residual;
  a.stats = stats;
  if (residual || stats) {
    ConvArgs probe = a;
    probe.bf16 = bf16_conv_shape(c_in, c_out, precision);
    SCN_CHECK(precision != SCN_FP32 && conv_tma_supported(probe) && (uintptr_t)residual % 16 == 0,
              "SubmanifoldConvolution: fused residual / statistics need the tensor-core path (see scn_fuses_residual)");
  }
  if (L->guided) {          // normal-guided taps: the forward table with every output row's taps permuted by its class
    ensure_guided_tables(&h->m, L, s);
    a.tbl = L->nbr_g.p;
  } else {
    use_sorted_tiles(a, L, precision, s);
  }
  run_conv(a, weight, true, precision, s, take_bf16_hint(&h->m, in, (long long)L->n * c_in, s));
  if (macs) *macs = (double)L->n_rules * c_in * c_out;   // flops += nRules*ip*op, CPU/Convolution.cpp:134
  SCN_CATCH
}

int scn_subm_bwd(scn_meta *h, const int64_t size[3], const float *in, const float *d_out, const float *weight,
                 float *d_in, float *d_weight, float *d_bias, int c_in, int c_out, int precision, void *stream) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  Level *L = subm_level(&h->m, size, s);
  // dgrad: d_in[i] = sum_k d_out[nbr[26-k][i]] * W[k]^T  (rule (i,o) at offset k  <=>  o sits at offset 26-k of i)
  // for this product K = c_out and N = c_in, so the caller's [27][c_in][c_out] array reads as [V][N][K]
  const bool dgrad16 = d_in && bf16_conv_shape(c_out, c_in, precision), wgrad16 = bf16_wgrad_shape(c_in, c_out, precision);
  Bf16Copy g16, x16;
  const long long ld_g = take_grad_ld(&h->m, c_out, (!d_in || dgrad16) && wgrad16, d_bias);
  const uint16_t *pg = grad_bf16(&h->m, dgrad16 || wgrad16, g16, d_out, ld_g, (long long)L->n, c_out, s);
  const uint16_t *px = take_bf16_hint(&h->m, in, (long long)L->n * c_in, s);
  if (!wgrad16) px = nullptr;
  else if (!px) px = x16.make(in, (long long)L->n * c_in, s);
  ConvArgs a;
  a.in = d_out; a.out = d_in;
  a.tbl = L->nbr.p; a.tbl_stride = L->n_pad; a.n_rows = L->n; a.V = 27; a.c_in = c_out; a.c_out = c_in; a.mirror = true; a.n_rules = L->n_rules; a.in_rows = L->n;
  BnbScratch scr;
  if (d_in && L->guided) {
    // the taps an input row feeds depend on the class of each OUTPUT row, so the transposed relation is one table per class
    // (a tap is used at most once per class); the three partial sums are accumulated through the epilogue's residual operand
    SCN_CHECK(!h->m.bnb.x, "fused BatchNorm backward is not available on normal-guided scales");
    ensure_guided_tables(&h->m, L, s);
    a.mirror = false;
    for (int c = 0; c < 3; ++c) {
      ConvArgs ac = a;
      ac.tbl = L->nbr_t[c].p;
      ConvArgs probe = ac;
      probe.bf16 = bf16_conv_shape(ac.c_in, ac.c_out, precision);
      const bool tc = precision != SCN_FP32 && conv_tma_supported(probe);
      if (c > 0) {
        if (tc) {
          ac.residual = d_in;
          run_conv(ac, weight, false, precision, s, dgrad16 ? pg : nullptr);
        } else {           // exact-fp32 kernels have no residual operand: accumulate through a temporary
          DevBuf<float> part;
          part.alloc((size_t)L->n * c_in, s);
          ac.out = part.p;
          run_conv(ac, weight, false, precision, s, nullptr);
          axpy(part.p, d_in, (long long)L->n * c_in, s);
          part.release(s);
        }
      } else {
        run_conv(ac, weight, false, precision, s, dgrad16 ? pg : nullptr);
      }
    }
  } else if (d_in) {
    use_sorted_tiles(a, L, precision, s);
    apply_bnb_hint(&h->m, a, precision, scr, s);
    run_conv(a, weight, false, precision, s, dgrad16 ? pg : nullptr);
  } else {
    SCN_CHECK(!h->m.bnb.x, "fused BatchNorm backward needs d_in");
  }
  scr.release(s);
  if (L->guided) ensure_guided_tables(&h->m, L, s);
  WgradArgs w;
  w.a = in; w.b = d_out; w.dw = d_weight; w.tbl = L->guided ? L->nbr_g.p : L->nbr.p; w.tbl_stride = L->n_pad; w.n_rows = L->n; w.V = 27;
  w.c_a = c_in; w.c_b = c_out; w.table_on_a = true; w.n_rules = L->n_rules; w.g_rows = L->n; w.s_rows = L->n;
  run_wgrad(w, L->guided ? L->nbr_g_pairs : L->nbr_pairs, precision, s, px, pg);
  g16.release(s);
  x16.release(s);
  if (d_bias) bias_grad(d_out, d_bias, L->n, c_out, s);
  SCN_CATCH
}

// ---- strided convolution: fine -> coarse ------------------------------------------------------------
int scn_conv_fwd(scn_meta *h, const int64_t in_size[3], const int64_t out_size[3], const float *in, const float *weight,
                 const float *bias, float *out, int c_in, int c_out, int precision, void *stream, double *macs) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  Level *F = need_level(&h->m, in_size, "Convolution");
  Level *C = ensure_coarse_level(&h->m, F, out_size, s);
  note_depth(&h->m);
  // out[p] = sum_k in[child[k][p]] * W[k]
  ConvArgs a;
  a.in = in; a.bias = bias; a.out = out;
  a.tbl = F->child.p; a.tbl_stride = C->n_pad; a.n_rows = C->n; a.V = 8; a.c_in = c_in; a.c_out = c_out; a.n_rules = F->n; a.in_rows = F->n;
  a.stats = h->m.next_stats;
  h->m.next_stats = nullptr;
  if (a.stats) {
    ConvArgs probe = a;
    probe.bf16 = bf16_conv_shape(c_in, c_out, precision);
    SCN_CHECK(precision != SCN_FP32 && conv_tma_supported(probe), "scn_out_stats: this layer does not run on the tensor-core path");
  }
  run_conv(a, weight, true, precision, s, take_bf16_hint(&h->m, in, (long long)F->n * c_in, s));
  if (macs) *macs = (double)F->n * c_in * c_out;   // every fine row has exactly one rule
  SCN_CATCH
}

int scn_conv_bwd(scn_meta *h, const int64_t in_size[3], const int64_t out_size[3], const float *in, const float *d_out,
                 const float *weight, float *d_in, float *d_weight, float *d_bias, int c_in, int c_out, int precision,
                 void *stream) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  Level *F = need_level(&h->m, in_size, "Convolution");
  Level *C = ensure_coarse_level(&h->m, F, out_size, s);
  // one bf16 copy of d_out serves both the dgrad (gathered operand) and the weight gradient
  const bool dgrad16 = bf16_conv_shape(c_out, c_in, precision), wgrad16 = bf16_wgrad_shape(c_in, c_out, precision);
  const long long ld_g = take_grad_ld(&h->m, c_out, dgrad16 && wgrad16, d_bias);
  Bf16Copy g16;
  const uint16_t *pg = grad_bf16(&h->m, dgrad16 || wgrad16, g16, d_out, ld_g, (long long)C->n, c_out, s);
  // dgrad: d_in[child[k][p]] = d_out[p] * W[k]^T  (scatter; each fine row has one parent)
  run_up(F, C, d_out, weight, false, d_in, c_out, c_in, precision, s, dgrad16 ? pg : nullptr, &h->m);
  WgradArgs w;
  w.a = in; w.b = d_out; w.dw = d_weight; w.tbl = F->child.p; w.tbl_stride = C->n_pad; w.n_rows = C->n; w.V = 8;
  w.c_a = c_in; w.c_b = c_out; w.table_on_a = true; w.n_rules = F->n; w.g_rows = F->n; w.s_rows = C->n;
  run_wgrad(w, F->child_pairs, precision, s, take_bf16_hint(&h->m, in, (long long)F->n * c_in, s), wgrad16 ? pg : nullptr);
  g16.release(s);
  if (d_bias) bias_grad(d_out, d_bias, C->n, c_out, s);
  SCN_CATCH
}

// ---- deconvolution: coarse -> fine, same rulebook with the roles swapped ------------------------------
int scn_deconv_fwd(scn_meta *h, const int64_t in_size[3], const int64_t out_size[3], const float *in,
                   const float *weight, const float *bias, float *out, int c_in, int c_out, int precision, void *stream,
                   double *macs) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  Level *F = need_level(&h->m, out_size, "Deconvolution");
  SCN_CHECK(F->coarse && find_level(&h->m, in_size) == F->coarse,
            "Deconvolution: the matching Convolution has not created this pair of scales");
  Level *C = F->coarse;
  SCN_CHECK(bias == nullptr, "Deconvolution: bias is not supported on this path (the UNet uses bias=False)");
  // out[child[k][p]] = in[p] * W[k]
  double *stats = h->m.next_stats;
  h->m.next_stats = nullptr;
  run_up(F, C, in, weight, true, out, c_in, c_out, precision, s, take_bf16_hint(&h->m, in, (long long)C->n * c_in, s), nullptr, stats);
  if (macs) *macs = (double)F->n * c_in * c_out;
  SCN_CATCH
}

int scn_deconv_bwd(scn_meta *h, const int64_t in_size[3], const int64_t out_size[3], const float *in,
                   const float *d_out, const float *weight, float *d_in, float *d_weight, float *d_bias, int c_in,
                   int c_out, int precision, void *stream) {
  SCN_TRY
  cudaStream_t s = note_stream(stream);
  check_channels(c_in, c_out);
  Level *F = need_level(&h->m, out_size, "Deconvolution");
  SCN_CHECK(F->coarse && find_level(&h->m, in_size) == F->coarse,
            "Deconvolution: the matching Convolution has not created this pair of scales");
  Level *C = F->coarse;
  // dgrad: d_in[p] = sum_k d_out[child[k][p]] * W[k]^T
  ConvArgs a;
  a.in = d_out; a.out = d_in;
  a.tbl = F->child.p; a.tbl_stride = C->n_pad; a.n_rows = C->n; a.V = 8; a.c_in = c_out; a.c_out = c_in; a.n_rules = F->n; a.in_rows = F->n;
  const bool dgrad16 = bf16_conv_shape(c_out, c_in, precision), wgrad16 = bf16_wgrad_shape(c_out, c_in, precision);
  const long long ld_g = take_grad_ld(&h->m, c_out, dgrad16 && wgrad16, d_bias);
  Bf16Copy g16;
  const uint16_t *pg = grad_bf16(&h->m, dgrad16 || wgrad16, g16, d_out, ld_g, (long long)F->n, c_out, s);
  BnbScratch scr;
  apply_bnb_hint(&h->m, a, precision, scr, s);
  run_conv(a, weight, false, precision, s, dgrad16 ? pg : nullptr);
  scr.release(s);
  // dW[k] = sum_p in[p]^T d_out[child[k][p]]
  WgradArgs w;
  w.a = in; w.b = d_out; w.dw = d_weight; w.tbl = F->child.p; w.tbl_stride = C->n_pad; w.n_rows = C->n; w.V = 8;
  w.c_a = c_in; w.c_b = c_out; w.table_on_a = false; w.n_rules = F->n; w.g_rows = F->n; w.s_rows = C->n;
  run_wgrad(w, F->child_pairs, precision, s, take_bf16_hint(&h->m, in, (long long)C->n * c_in, s), wgrad16 ? pg : nullptr);
  g16.release(s);
  if (d_bias) bias_grad(d_out, d_bias, F->n, c_out, s);
  SCN_CATCH
}

// ---- batch norm ---------------------------------------------------------------------------------------
int scn_bf16_operand(scn_meta *h, const float *fp32, void *bf16, int ready) {
  SCN_TRY
  SCN_CHECK(h, "null handle");
  h->m.hint_src = fp32;
  h->m.hint_bf16 = bf16;
  h->m.hint_ready = ready;
  SCN_CATCH
}

int scn_bn_bwd_fusion(scn_meta *h, const float *bn_in, const float *save_mean, const float *save_invstd, const float *gamma,
                      const float *beta, float leakiness, double *acc) {
  SCN_TRY
  SCN_CHECK(h && bn_in && save_mean && save_invstd && acc, "scn_bn_bwd_fusion: null argument");
  h->m.bnb.x = bn_in; h->m.bnb.mean = save_mean; h->m.bnb.invstd = save_invstd; h->m.bnb.gamma = gamma; h->m.bnb.beta = beta;
  h->m.bnb.leak = leakiness; h->m.bnb.acc = acc;
  SCN_CATCH
}

int scn_bn_bwd_fusable(int c_in, int c_out, int precision) {
  // the dgrad product of a [c_in -> c_out] layer contracts over c_out and produces c_in columns
  const int kel = bf16_conv_shape(c_out, c_in, precision) ? 64 : 32;
  return precision != SCN_FP32 && c_out >= kel && c_out % kel == 0 && c_in >= 32 && c_in % 32 == 0;
}

int scn_bn_bwd_apply(const float *in, const float *d_masked, const double *acc, const float *save_mean, const float *save_invstd,
                     const float *gamma, const float *d_in_add, int64_t ld_add, float *d_in, void *d_in_bf16, float *d_gamma,
                     float *d_beta, int64_t n, int C, void *stream) {
  SCN_TRY
  ProfScope ps(PK_BN, ((d_in_add ? 4.0 : 3.0) * 4.0 + (d_in_bf16 ? 2.0 : 0.0)) * (double)n * C, 0.0, note_stream(stream));
  bn_bwd_apply(in, d_masked, acc, save_mean, save_invstd, gamma, d_in_add, ld_add, d_in, (uint16_t *)d_in_bf16, d_gamma, d_beta,
               n, C, note_stream(stream));
  SCN_CATCH
}

int scn_grad_bf16(scn_meta *h, const float *d_out, const void *bf16) {
  SCN_TRY
  SCN_CHECK(h, "null handle");
  h->m.ghint_src = d_out;
  h->m.ghint_bf16 = bf16;
  SCN_CATCH
}

int scn_out_stats(scn_meta *h, double *stats) {
  SCN_TRY
  SCN_CHECK(h, "null handle");
  h->m.next_stats = stats;
  SCN_CATCH
}

int scn_grad_stride(scn_meta *h, int64_t ld) {
  SCN_TRY
  SCN_CHECK(h && ld >= 0, "scn_grad_stride: bad argument");
  h->m.next_grad_ld = ld;
  SCN_CATCH
}

int scn_fuses_residual(int c_in, int c_out, int precision) {
  const int kel = bf16_conv_shape(c_in, c_out, precision) ? 64 : 32;
  return precision != SCN_FP32 && c_in >= kel && c_in % kel == 0 && c_out >= 32 && c_out % 32 == 0;
}

int scn_bf16_plan(int c_in, int c_out, int precision) {
  return (bf16_conv_shape(c_in, c_out, precision) ? 1 : 0) | (bf16_wgrad_shape(c_in, c_out, precision) ? 2 : 0);
}

int scn_bn_fwd(const float *in, float *out, void *out_bf16, const double *stats_in, float *save_mean, float *save_invstd,
               float *running_mean, float *running_var, const float *gamma, const float *beta, int64_t n, int C, float eps,
               float momentum, int train, float leakiness, void *stream) {
  SCN_TRY
  ProfScope ps(PK_BN, ((out_bf16 ? 3.5 : 3.0) - (stats_in && train ? 1.0 : 0.0) - (out ? 0.0 : 1.0)) * 4.0 * (double)n * C, 0.0,
               note_stream(stream));
  bn_fwd(in, out, (uint16_t *)out_bf16, stats_in, save_mean, save_invstd, running_mean, running_var, gamma, beta, n, C, eps, momentum, train != 0,
         leakiness, note_stream(stream));
  SCN_CATCH
}

int scn_bn_bwd(const float *in, const float *out, const float *d_out, const float *save_mean, const float *save_invstd,
               const float *gamma, const float *beta, const float *d_in_add, float *d_in, float *d_gamma, float *d_beta,
               int64_t n, int C, float leakiness, void *stream) {
  SCN_TRY
  ProfScope ps(PK_BN, (d_in_add ? 6.0 : 5.0) * 4.0 * (double)n * C, 0.0, note_stream(stream));
  bn_bwd(in, out, d_out, save_mean, save_invstd, gamma, beta, d_in_add, d_in, d_gamma, d_beta, n, C, leakiness,
         note_stream(stream));
  SCN_CATCH
}

}  // extern "C"

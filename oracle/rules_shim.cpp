// TEST INFRASTRUCTURE ONLY -- never imported by the product path (occuseg_b200/).
//
// Pins the rulebook oracle to REFERENCE-COMPILED code.  The reference's GPU rule builders need cudpp / CUDA 9 and
// cannot be built, but its CPU `SparseGrid` builders -- the code the GPU path was written
// against and is self-checked against (Metadata/ConvolutionRules.h:786-815) -- are plain C++ templates:
//
//   inputLayerRules                                   Metadata/IOLayersRules.h:19-130
//   SubmanifoldConvolution_SgToRules (plain, dilated) Metadata/SubmanifoldConvolutionRules.h:114-153
//   SubmanifoldConvolution_SgToRules (normal-guided)  Metadata/SubmanifoldConvolutionRules.h:159-209
//   remap_rules_with_normal                           Metadata/SubmanifoldConvolutionRules.h:213-245
//   Convolution_InputSgToRulesAndOutputSg (2 forms)   Metadata/ConvolutionRules.h:18-119
//   RectangularRegion / OrientedFilter / region calculators   Metadata/RectangularRegions.h (whole file)
//   class Float3                                      Metadata/Metadata.h:75-103
//
// oracle/build_rules_ref.py cuts exactly those spans out of the reference headers AT BUILD TIME into
// oracle/_build/rules_extracted_*.inc (git-ignored; nothing from the reference is committed) and compiles this file
// around them.  What this file supplies is only what the spans need from the parts of Metadata.h that cannot be
// built: Int / Point (Metadata/32bits.h:11,15), RuleBook (Metadata.h:73), SparseGrid with google::dense_hash_map
// replaced by std::unordered_map (same find/insert/iteration interface; iteration ORDER differs, which is why every
// comparison goes through coordinates and canonical sorting), volume<D> (Metadata.cpp), and no-op profiler macros.
//
// extern "C" entry points (ctypes, see oracle/rules_ref.py) return rule RELATIONS through coordinates:
// each rule = (offset index, input voxel xyz, output voxel xyz), so that row numbering conventions do not matter.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>

using Int = int32_t;
template <Int dimension> using Point = std::array<Int, dimension>;
template <Int dimension> struct IntArrayHash {
  std::size_t operator()(Point<dimension> const &p) const {
    std::size_t h = 1469598103934665603ull;
    for (auto x : p) h = (h ^ (std::size_t)(uint32_t)x) * 1099511628211ull;
    return h;
  }
};
using namespace std;   // the reference headers rely on it (Metadata.h:51)

template <Int dimension>
using SparseGridMap = std::unordered_map<Point<dimension>, Int, IntArrayHash<dimension>, std::equal_to<Point<dimension>>>;
template <Int dimension> class SparseGrid {
public:
  Int ctr;
  SparseGridMap<dimension> mp;
  SparseGrid() : ctr(0) {}
};
template <Int dimension> using SparseGrids = std::vector<SparseGrid<dimension>>;
using RuleBook = std::vector<std::vector<Int>>;
template <Int dimension> Int volume(long *point) {
  Int v = 1;
  for (Int i = 0; i < dimension; i++) v *= (Int)point[i];
  return v;
}
#define EASY_FUNCTION(...)
#define EASY_VALUE(...)
#define EASY_BLOCK(...)
#define EASY_END_BLOCK

// ---- spans of the reference headers, cut at build time (see build_rules_ref.py) -------------------------------
#include "rules_extracted_float3.inc"          // class Float3
#include "rules_extracted_regions.inc"         // RectangularRegions.h
#include "rules_extracted_input.inc"           // inputLayerRules
#include "rules_extracted_submanifold.inc"     // SubmanifoldConvolution_SgToRules x2, remap_rules_with_normal
#include "rules_extracted_convolution.inc"     // Convolution_InputSgToRulesAndOutputSg x2

// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Scene {
  SparseGrids<3> grids;                 // one per sample, as the reference keeps them
  RuleBook input_rules;
  Int n_active = 0;
  std::vector<Point<3>> row_xyz;        // row -> coordinates
  std::vector<Int> row_batch;
};

Scene *g_scene = nullptr;

void index_rows(Scene &s) {
  s.row_xyz.assign(s.n_active, Point<3>{0, 0, 0});
  s.row_batch.assign(s.n_active, 0);
  for (size_t b = 0; b < s.grids.size(); ++b)
    for (auto const &it : s.grids[b].mp) {
      s.row_xyz[it.second + s.grids[b].ctr] = it.first;
      s.row_batch[it.second + s.grids[b].ctr] = (Int)b;
    }
}

}  // namespace

extern "C" {

// Build the reference's CPU grids from a point list (long [P][4] = x,y,z,batch) with inputLayerRules(mode).
// Returns nActive.  row_of_point[P] receives the reference's row of every point; rows are numbered in first-appearance
// order on this path (IOLayersRules.h:85-92), so callers compare through row_xyz.
int rules_ref_input(const long *coords, int n_points, int batch_size, int mode, int32_t *row_of_point, int32_t *max_active) {
  delete g_scene;
  g_scene = new Scene();
  Scene &s = *g_scene;
  inputLayerRules<3>(s.grids, s.input_rules, const_cast<long *>(coords), (Int)n_points, 4, (Int)batch_size, (Int)mode,
                     s.n_active);
  index_rows(s);
  if (max_active) *max_active = s.input_rules[0][1];
  if (row_of_point && s.input_rules.size() > 1) {
    const Int ma = s.input_rules[0][1];
    const std::vector<Int> &r = s.input_rules[1];
    for (Int row = 0; row < s.n_active; ++row) {
      const Int *e = &r[(size_t)row * (ma + 1)];
      for (Int j = 0; j < e[0]; ++j) row_of_point[e[1 + j]] = row;
    }
  }
  return s.n_active;
}

// the rule table of the InputLayer as the reference lays it out: [nActive][1 + maxActive] (count, then point ids in order)
int rules_ref_input_table(int32_t *out) {
  if (!g_scene || g_scene->input_rules.size() < 2) return -1;
  const std::vector<Int> &r = g_scene->input_rules[1];
  std::memcpy(out, r.data(), r.size() * sizeof(Int));
  return (int)r.size();
}

// coordinates (x,y,z,batch) of every row
int rules_ref_locations(int32_t *out) {
  if (!g_scene) return -1;
  for (Int i = 0; i < g_scene->n_active; ++i) {
    for (int d = 0; d < 3; ++d) out[4 * i + d] = g_scene->row_xyz[i][d];
    out[4 * i + 3] = g_scene->row_batch[i];
  }
  return g_scene->n_active;
}

// Submanifold 3x3x3 rules of the current scene.
//   variant 0: SubmanifoldConvolution_SgToRules(grid, rules, size, dilated_rate)          (CPU enumeration: z outermost, x innermost)
//   variant 1: SubmanifoldConvolution_SgToRules(grid, rules, size, normal, dilated_rate)  (region enumeration: x outermost, z innermost,
//              tap permuted by OrientedFilter(normal[row]))
//   variant 2: variant 0 followed by remap_rules_with_normal(rules, normal)
//   variant 3: variant 0, lists re-indexed to the GPU enumeration (x outermost), then remap_rules_with_normal -- what the
//              GPU path does (SubmanifoldConvolutionRules.h:486-490 on rules built by SubmanifoldRules_cuda.cu:63-73)
// normals: float [nActive][3] in THIS scene's row order (may be NULL for variant 0).
// out (may be NULL to size): int32 [n][4] = (offset, in_row, out_row, 0); returns n.
long rules_ref_submanifold(int variant, int dilated_rate, const float *normals, int32_t *out) {
  if (!g_scene) return -1;
  Scene &s = *g_scene;
  RuleBook rules(27);
  long size[3] = {3, 3, 3};
  std::vector<Float3> nv;
  if (normals)
    for (Int i = 0; i < s.n_active; ++i) nv.push_back(Float3(normals + 3 * i));
  for (auto &g : s.grids) {
    if (variant == 1) {
      // normal[] is indexed with the id stored in the grid (SubmanifoldConvolutionRules.h:177-178), which on this CPU
      // path is already the global row (IOLayersRules.h:88)
      SubmanifoldConvolution_SgToRules<3>(g, rules, size, nv, dilated_rate);
    } else {
      SubmanifoldConvolution_SgToRules<3>(g, rules, size, dilated_rate);
    }
  }
  if (variant == 3) {     // lists re-indexed from the CPU enumeration (z outermost) to the GPU one (x outermost), then remapped
    RuleBook t(27);
    for (int k = 0; k < 27; ++k) t[(k % 3) * 9 + ((k / 3) % 3) * 3 + k / 9].swap(rules[k]);
    rules.swap(t);
  }
  if (variant == 2 || variant == 3) remap_rules_with_normal<3>(rules, nv);
  long n = 0;
  for (int k = 0; k < 27; ++k) {
    for (size_t j = 0; j + 1 < rules[k].size(); j += 2) {
      if (out) {
        out[4 * n + 0] = k;
        out[4 * n + 1] = rules[k][j];
        out[4 * n + 2] = rules[k][j + 1];
        out[4 * n + 3] = 0;
      }
      ++n;
    }
  }
  return n;
}

// Size-2 / stride-2 Convolution rules of the current scene through Convolution_InputSgToRulesAndOutputSg
// (ConvolutionRules.h:95-119; with normals: :18-92).  Output rows are numbered in creation order, so the coarse
// coordinates are returned too.  out: int32 [n][3] = (offset, in_row, out_row); coarse_xyzb: int32 [nCoarse][4];
// out_normals (normal-guided form only): float [nCoarse][3].  Returns n; *n_coarse receives the coarse row count.
long rules_ref_strided(const float *normals, int32_t *out, int32_t *coarse_xyzb, float *out_normals, int32_t *n_coarse,
                       const long *in_size, const long *out_size) {
  if (!g_scene) return -1;
  Scene &s = *g_scene;
  long size[3] = {2, 2, 2}, stride[3] = {2, 2, 2};
  long isz[3] = {in_size[0], in_size[1], in_size[2]}, osz[3] = {out_size[0], out_size[1], out_size[2]};
  RuleBook rules;
  SparseGrids<3> outs(s.grids.size());
  std::vector<Float3> nv, onv;
  if (normals)
    for (Int i = 0; i < s.n_active; ++i) nv.push_back(Float3(normals + 3 * i));
  Int ctr = 0;
  for (size_t b = 0; b < s.grids.size(); ++b) {
    outs[b].ctr = ctr;      // as Convolution_InputSgsToRulesAndOutputSgs does: the output grid counts on from the previous sample
    if (normals)
      Convolution_InputSgToRulesAndOutputSg<3>(s.grids[b], outs[b], rules, size, stride, isz, osz, nv, onv);
    else
      Convolution_InputSgToRulesAndOutputSg<3>(s.grids[b], outs[b], rules, size, stride, isz, osz);
    ctr = outs[b].ctr;
  }
  if (n_coarse) *n_coarse = ctr;
  if (coarse_xyzb)
    for (size_t b = 0; b < outs.size(); ++b)
      for (auto const &it : outs[b].mp) {
        for (int d = 0; d < 3; ++d) coarse_xyzb[4 * it.second + d] = it.first[d];
        coarse_xyzb[4 * it.second + 3] = (Int)b;
      }
  if (out_normals)
    for (size_t i = 0; i < onv.size(); ++i) {
      out_normals[3 * i] = onv[i].x; out_normals[3 * i + 1] = onv[i].y; out_normals[3 * i + 2] = onv[i].z;
    }
  long n = 0;
  for (size_t k = 0; k < rules.size(); ++k)
    for (size_t j = 0; j + 1 < rules[k].size(); j += 2) {
      if (out) {
        out[3 * n + 0] = (Int)k;
        out[3 * n + 1] = rules[k][j];
        out[3 * n + 2] = rules[k][j + 1];
      }
      ++n;
    }
  return n;
}

int rules_ref_oriented_filter(float x, float y, float z) { return OrientedFilter(Float3(x, y, z)); }

void rules_ref_clear() {
  delete g_scene;
  g_scene = nullptr;
}

}  // extern "C"

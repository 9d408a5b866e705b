/*
 * scn_b200.h -- C ABI of libscn_b200.so: the B200 (sm_100a) replacement for the hot path behind
 * OccuSeg's pybind11 module `sparseconvnet.SCN` (reference: sparseconvnet/SCN/pybind.cpp:11-239,
 * sparseconvnet/SCN/sparseconvnet.h:9-239).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - every feature matrix is dense row-major fp32 [rows, channels] in DEVICE memory, one row per
 *    active voxel, samples concatenated (reference layout, SURVEY.md "Vocabulary");
 *  - weights are the reference's [V, Cin, Cout] fp32 (V=27 submanifold, k=(dx+1)*9+(dy+1)*3+(dz+1),
 *    CUDA/SubmanifoldRules_cuda.cu:63-73; V=8 strided, k=(x&1)*4+(y&1)*2+(z&1), :549-554);
 *  - a spatial size is the int64[3] the Python layer carries (cubes in practice); it names a scale
 *    inside a handle exactly as Metadata's maps are keyed (Metadata/Metadata.h:238-248);
 *  - all work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed as void*);
 *    the only host synchronisations are the ones that return a row count to the caller;
 *  - every entry returns 0 on success, non-zero on failure, message via scn_last_error().  The
 *    library never calls exit()/abort() (the reference does: CUDPPWrapper.hpp:15-25,
 *    Convolution.cu:14-24);
 *  - the library never allocates caller-visible memory: the caller asks for row counts, allocates,
 *    and passes pointers.  Outputs are OVERWRITTEN (d_weight too -- the Python wrapper hands in
 *    zeros like the reference does, submanifoldConvolution.py:113-115, so both conventions agree).
 *
 * precision: SCN_FP32 = exact fp32 FMA path (rel 1e-5 vs the reference CPU arithmetic);
 *            SCN_TF32 = tcgen05 kind::tf32 tiles, fp32 accumulate in TMEM (rel 2e-2), used only
 *                       when Cin and Cout are multiples of 32 and >= 32, else falls to SCN_FP32;
 *            SCN_BF16 = tcgen05 kind::f16 tiles on bf16 COPIES of the fp32 operands (made inside the call), fp32
 *                       accumulate, fp32 outputs (rel 2e-2); needs the contraction width to be a multiple of 64,
 *                       else falls to SCN_TF32.
 */
#ifndef SCN_B200_H
#define SCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scn_meta scn_meta; /* opaque; replaces Metadata<3> (Metadata/Metadata.h:218-364) */

enum { SCN_FP32 = 0, SCN_TF32 = 1, SCN_BF16 = 2 };

/* ---- library --------------------------------------------------------------------------------- */
int scn_version(void);
const char *scn_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t scn_launch_count(void);
/* optional per-kernel-family timing with CUDA events on the launching stream (used by bench.py for the roofline
 * line).  scn_profile(1) clears and starts, scn_profile_read fills out[kind*4 + {launches, milliseconds,
 * algorithmic bytes, flops}] for kinds 0..n-1 and returns n; scn_profile_kind_name(kind) names them. */
void scn_profile(int enable);
int scn_profile_read(double *out, int max_kinds);
const char *scn_profile_kind_name(int kind);

/* ---- handle: replaces py `Metadata_3()` / ~Metadata (pybind.cpp:11-13) ------------------------ */
scn_meta *scn_meta_create(int device);
void scn_meta_destroy(scn_meta *m);
/* The library's scratch (rulebooks, operand copies) lives in the device's default stream-ordered memory pool, of which it
 * keeps at most SCN_POOL_KEEP_MB (default 4096) cached across synchronisations.  scn_pool_trim releases the free part of
 * that cache down to keep_bytes right now (e.g. before another allocator needs the memory). */
int scn_pool_trim(int device, int64_t keep_bytes);
/* Tile order of the tensor-core submanifold products: rows of every `block_rows`-row block are processed sorted by their
 * 27-bit neighbourhood pattern (fewer non-empty (tile, tap) pairs, denser gathers); results are identical up to the
 * order of fp32 accumulation inside the BatchNorm statistics.  0 = natural row order.  Default 262144, or the value of the
 * SCN_TILE_SORT environment variable.  Applies to handles created afterwards.  Returns the previous setting. */
int scn_tile_sort(int block_rows);
/* Deterministic mode (default off, or the SCN_DETERMINISTIC environment variable): weight gradients and column statistics --
 * everywhere partial results of row ranges / CTAs are otherwise merged with floating-point atomics (as the reference merges them,
 * CUDA/Convolution.cu) -- are written one partial per range into scratch and added in range order, so that two runs on the same
 * inputs give bit-identical outputs and gradients.  Costs a few per cent of extra traffic.  Returns the previous setting. */
int scn_deterministic(int on);

/* ---- InputLayer: replaces InputLayer_updateOutput's Metadata::inputLayer -> inputLayerRulesSimple
 * (CUDA/IOLayers.cpp:17-80, Metadata/Metadata.cpp:425-437, Metadata/IOLayersRules.h:136-202).
 * coords: int64 [P,4] (x,y,z,batch), batch column ascending; host pointer unless coords_on_device.
 * Builds the finest scale: voxel row = rank of (z,y,x) among the sample's unique voxels + rows of
 * earlier samples.  *n_active receives the row count (one host sync). mode must be 3 (sum) or 4 (mean). */
int scn_input_layer_build(scn_meta *m, const int64_t spatial_size[3], const int64_t *coords, int coords_on_device,
                          int64_t n_points, int batch_size, int mode, void *stream, int64_t *n_active);
/* Normal-guided kernels (OccuSeg's `use_normal`; InputLayer_updateOutput's `normals` argument, CUDA/IOLayers.cpp:39-66).
 * Call BEFORE scn_input_layer_build: point_normals float [P,3] in DEVICE memory, one per point of that call.  Every voxel gets
 * the normalised mean of its points' normals and the orientation class OrientedFilter(normal) in {0,2,4}
 * (Metadata/RectangularRegions.h:12-31); submanifold rules of a scale that carries normals use, for every OUTPUT row, the tap
 * permutation of its class (remap_rules_with_normal, Metadata/SubmanifoldConvolutionRules.h:213-245, table
 * SubmanifoldRules_cuda.cu:8-15), forward, dgrad and wgrad.  Strided layers whose INPUT scale is >= normal_guide_scale
 * (Metadata::setNormalGuideScale, ConvolutionRules.h:774) hand the coarse scale the normalised mean of the children's normals
 * and permute their 8 taps by the coarse row's class, as the reference's CPU builder does (ConvolutionRules.h:18-92; its GPU
 * variant, :139-236, advances its query index twice per rule and is not reproduced); below that scale the plain rules are used
 * and the coarse scale carries no normals.  Not combined with dilation or with the fused training BatchNorm backward. */
int scn_input_normals(scn_meta *m, const float *point_normals, int normal_guide_scale);
int scn_guided(scn_meta *m, const int64_t spatial_size[3]);       /* 1 when the scale carries normals */
/* parity access: the guided forward table int32 [27, N] (HOST) and the orientation class of every row (uint8 [N], may be NULL) */
int scn_subm_guided_table(scn_meta *m, const int64_t spatial_size[3], void *stream, int32_t *out_host, uint8_t *ori_host);
/* parity access: the per-voxel normals float [N,3] (HOST) of a scale that carries them (Metadata::normals, Metadata.h) */
int scn_normals(scn_meta *m, const int64_t spatial_size[3], void *stream, float *out_host);
/* feature part of InputLayer_updateOutput (CUDA/IOLayers.cu:16-43): out[N,C] = sum/mean of the points of each voxel */
int scn_input_layer_fwd(scn_meta *m, const float *point_feats, int channels, float *out, void *stream);
/* InputLayer_updateGradInput (CUDA/IOLayers.cpp:81-105): d_point[P,C] = (1/n) * d_out[row(p)] */
int scn_input_layer_bwd(scn_meta *m, const float *d_out, int channels, float *d_point_feats, void *stream);
/* OutputLayer_updateOutput (CUDA/IOLayers.cpp:107-130): out[P,C] = in[row(p)] */
int scn_output_layer_fwd(scn_meta *m, const float *in, int channels, float *out_points, void *stream);
/* OutputLayer_updateGradInput (CUDA/IOLayers.cpp:131-154): d_in[N,C] = sum over the voxel's points of d_out[p] */
int scn_output_layer_bwd(scn_meta *m, const float *d_out_points, int channels, float *d_in, void *stream);
int64_t scn_n_points(scn_meta *m);
/* The step before the InputLayer, on the device: float point cloud xyz [P,3] (already augmented and scaled to voxel units)
 * -> coords int64 [P,4] = (trunc(x - offset), batch_index), the list scn_input_layer_build takes with coords_on_device = 1;
 * keep[P] (may be NULL) = 1 where the shifted point lies inside [0, full_scale)^3.  Replaces the host-side shift / crop /
 * LongTensor conversion of examples/ScanNet/datasets/scannet.py:133-137,160,210 and sparseconvnet/ioLayers.py:56. */
int scn_float_coords(const float *xyz, int64_t n_points, const float offset[3], int batch_index, float full_scale,
                     int64_t *coords, uint8_t *keep, void *stream);

/* ---- scale queries: Metadata::getNActive / getSpatialLocations (Metadata.cpp:89-92, :724-748) -- */
int64_t scn_nactive(scn_meta *m, const int64_t spatial_size[3]); /* -1 if the scale does not exist */
/* locations int64 [N,4] (x,y,z,batch) in row order, written to HOST memory */
int scn_spatial_locations(scn_meta *m, const int64_t spatial_size[3], int64_t *out_host);

/* ---- rulebooks (built lazily and cached per scale like Metadata::getSubmanifoldRuleBook /
 * getRuleBook, Metadata.cpp:503-529,597-625).  Exposed so parity tests can compare them bit-exactly
 * with the reference's rule lists after canonical sorting. ------------------------------------- */
/* Dilated submanifold convolution (SubmanifoldConvolution(..., dilated_rate), submanifoldConvolution.py:36-53): the NEXT
 * scn_subm_rulebook / scn_subm_neighbour_table / scn_subm_fwd / scn_subm_fwd_bn / scn_subm_bwd on this handle uses the 27 taps
 * at offsets rate*(dx,dy,dz) -- the rule relation of the reference's SubmanifoldConvolution_SgToRules(grid, rules, size, rate)
 * (Metadata/SubmanifoldConvolutionRules.h:39-75,114-153; its GPU_GRID builder accepts the argument and ignores it,
 * :248-275).  One use; rate 1 = the ordinary neighbourhood.  Tables are cached per (scale, rate) like Metadata's map keyed by
 * TwoLongTensorsToPoint_Dilation (Metadata.cpp:511). */
int scn_subm_dilation(scn_meta *m, int rate);
/* ensure the 27-offset neighbour table of a scale exists; *n_rules = total rule count (centre included) */
int scn_subm_rulebook(scn_meta *m, const int64_t spatial_size[3], void *stream, int64_t *n_rules);
/* copy the table to HOST: int32 [27, N]; entry = input row feeding output row o at offset k, or -1 */
int scn_subm_neighbour_table(scn_meta *m, const int64_t spatial_size[3], int32_t *out_host);
/* build (if needed) the size-2/stride-2 link fine -> coarse and the coarse scale; *n_coarse = its rows */
int scn_strided_rulebook(scn_meta *m, const int64_t fine_size[3], const int64_t coarse_size[3], void *stream,
                         int64_t *n_coarse);
/* copy to HOST: parent int32 [Nfine] (coarse row of each fine row) and offset uint8 [Nfine] (0..7) */
int scn_strided_table(scn_meta *m, const int64_t fine_size[3], int32_t *parent_host, uint8_t *offset_host);

/* ---- ResolutionBasedScattering (sparseconvnet.h:63, sparseconvnet_cuda.cpp:203-209, Metadata/ConvolutionRules.h:327-342;
 * caller: sparseconvnet/utils.py:72-132 upsample_feature).  lr_xyz int32 [n_lr,3], hr_xyz int32 [n_hr,3] in DEVICE memory;
 * hr2lr[i] = row of the low-resolution voxel hr_xyz[i] / stride (integer division per axis), where the row of an lr voxel is
 * its rank among the sorted (z,y,x) unique lr voxels -- the index into lr_xyz when that list is a sample's spatial locations --
 * or -1 when no lr voxel exists there. */
int scn_resolution_scatter(const int32_t *lr_xyz, int64_t n_lr, const int32_t *hr_xyz, int64_t n_hr, int stride,
                           int32_t *hr2lr, void *stream);

/* ---- SCN_BF16 operand copies ----------------------------------------------------------------------
 * The SCN_BF16 kernels read bf16 COPIES of the fp32 feature matrices.  By default every entry makes (and drops) the
 * copies it needs.  A caller that keeps them saves those passes: scn_bf16_operand registers `bf16` ([rows, c] bf16,
 * caller-owned) as the copy of the fp32 matrix at `fp32` for the NEXT convolution entry on this handle whose `in`
 * argument is that pointer.  ready = 1: already filled (e.g. by scn_bn_fwd's out_bf16, or by an earlier forward
 * call); ready = 0: the entry fills it.  Forward entries use it as the gathered operand, backward entries as the
 * `in` operand of the weight gradient.  scn_bf16_plan: bit 0 = the forward product of a [c_in -> c_out] layer reads a
 * bf16 `in`, bit 1 = its weight gradient does (0 unless precision is SCN_BF16 and the widths allow). */
int scn_bf16_operand(scn_meta *m, const float *fp32, void *bf16, int ready);
int scn_bf16_plan(int c_in, int c_out, int precision);

/* ---- SubmanifoldConvolution_updateOutput / _backward (sparseconvnet.h:50-61; drivers
 * CUDA/Convolution.cpp:104-210; kernels CUDA/Convolution.cu:447-534,695-753,1059-1152) ----------
 * out[N,Cout] = sum_k in[nbr_k(o)] * W[k];  *macs = sum_k n_k*Cin*Cout (the reference's return value) */
int scn_subm_fwd(scn_meta *m, const int64_t spatial_size[3], const float *in, const float *weight, const float *bias,
                 const float *residual, double *stats, float *out, int c_in, int c_out, int precision, void *stream,
                 double *macs);
/* stats (extension, may be NULL): [2][Cout] fp64, receives the column sums and sums of squares of the result
 * (accumulated in the kernel epilogue) for the BatchNorm that consumes it: pass it to scn_bn_fwd as stats_in and that
 * call skips its reduction pass.  Same availability as residual. */
/* residual (extension, may be NULL): [N,Cout] fp32 added to the result in the kernel epilogue -- the shortcut of a
 * residual block (networkArchitectures.py:225-240) without a separate add pass.  Only the tensor-core kernels take it:
 * scn_fuses_residual(c_in, c_out, precision) says whether this layer shape does. */
int scn_fuses_residual(int c_in, int c_out, int precision);
/* Inference: SubmanifoldConvolution followed by BatchNormalization (running statistics) + (leaky) ReLU in ONE kernel --
 * the fused BatchNorm+ReLU epilogue: out = leaky(bn_scale[c] * (conv + bias + residual) + bn_shift[c]); out_bf16 (may be
 * NULL) also receives the bf16 copy of out for the next SCN_BF16 convolution.  bn_scale / bn_shift [Cout] come from
 * scn_bn_eval_coeffs, which evaluates them exactly as scn_bn_fwd(train = 0) does, so the result is bit-identical to
 * the two separate calls.  Tensor-core shapes only (scn_fuses_residual). */
int scn_subm_fwd_bn(scn_meta *m, const int64_t spatial_size[3], const float *in, const float *weight, const float *bias,
                    const float *residual, const float *bn_scale, const float *bn_shift, float leakiness, float *out,
                    void *out_bf16, int c_in, int c_out, int precision, void *stream, double *macs);
int scn_bn_eval_coeffs(const float *running_mean, const float *running_var, const float *gamma, const float *beta,
                       int channels, float eps, float *scale, float *shift, void *stream);
/* d_in[N,Cin] (NULL: the input gradient is not needed and is skipped), d_weight[27,Cin,Cout], d_bias[Cout] (NULL when no bias) */
int scn_subm_bwd(scn_meta *m, const int64_t spatial_size[3], const float *in, const float *d_out, const float *weight,
                 float *d_in, float *d_weight, float *d_bias, int c_in, int c_out, int precision, void *stream);

/* ---- Convolution (size 2, stride 2): sparseconvnet.h:89-102, CUDA/Convolution.cpp:36-102 ------- */
int scn_conv_fwd(scn_meta *m, const int64_t in_size[3], const int64_t out_size[3], const float *in, const float *weight,
                 const float *bias, float *out, int c_in, int c_out, int precision, void *stream, double *macs);
int scn_conv_bwd(scn_meta *m, const int64_t in_size[3], const int64_t out_size[3], const float *in, const float *d_out,
                 const float *weight, float *d_in, float *d_weight, float *d_bias, int c_in, int c_out, int precision,
                 void *stream);

/* ---- Deconvolution (size 2, stride 2), reuses the down-convolution's rulebook:
 * sparseconvnet.h:103-116, CUDA/Deconvolution.cpp:19-84.  in_size is the COARSE scale. ----------- */
int scn_deconv_fwd(scn_meta *m, const int64_t in_size[3], const int64_t out_size[3], const float *in,
                   const float *weight, const float *bias, float *out, int c_in, int c_out, int precision, void *stream,
                   double *macs);
int scn_deconv_bwd(scn_meta *m, const int64_t in_size[3], const int64_t out_size[3], const float *in,
                   const float *d_out, const float *weight, float *d_in, float *d_weight, float *d_bias, int c_in,
                   int c_out, int precision, void *stream);

/* ---- BatchNormalization_updateOutput / _backward with the fused (leaky) ReLU
 * (sparseconvnet.h:21-33, CUDA/BatchNormalization.cpp:21-71, BatchNormalization.cu:14-199) --------
 * train: batch statistics, running = momentum*running + (1-momentum)*batch (unbiased var);
 * y = leaky(gamma*(x-mean)*invstd + beta); leakiness 0 = ReLU, 1 = identity.  gamma/beta may be NULL. */
/* out_bf16: optional second output, the bf16 copy of `out` ([n_rows, channels], channels % 4 == 0) for the SCN_BF16
 * convolution that consumes it (see scn_bf16_operand); NULL = none */
/* stats_in: optional [2][channels] fp64 column sums / sums of squares of `in` made by the kernel that produced it
 * (scn_subm_fwd's stats); used in training mode instead of a reduction pass over `in`; NULL = compute here */
int scn_bn_fwd(const float *in, float *out, void *out_bf16, const double *stats_in, float *save_mean, float *save_invstd,
               float *running_mean,
               float *running_var, const float *gamma, const float *beta, int64_t n_rows, int channels, float eps,
               float momentum, int train, float leakiness, void *stream);
/* d_out is NOT modified (the reference masks it in place, BatchNormalization.cu:151-153; no caller observes it).
 * `out` (the forward output) is accepted for symmetry with the reference entry but never read: the activation mask is
 * recomputed from `in`, gamma and beta exactly as the forward pass computed it, so `out` may be NULL. */
int scn_bn_bwd(const float *in, const float *out, const float *d_out, const float *save_mean, const float *save_invstd,
               const float *gamma, const float *beta, const float *d_in_add, float *d_in, float *d_gamma, float *d_beta,
               int64_t n_rows, int channels, float leakiness, void *stream);
/* d_in_add (extension, may be NULL): [n_rows, channels] fp32 added to d_in in the same pass -- the gradient that
 * reaches the same input through a residual shortcut, instead of a separate accumulation pass. */

/* ---- training-mode BatchNorm+ReLU fused with the convolution it feeds (the UNet pattern BN -> ReLU -> conv,
 * networkArchitectures.py:225-229; the reference fuses BN with its ReLU, BatchNormalization.cu:61-69) ----------------
 * Forward: scn_bn_fwd with out = NULL and out_bf16 set writes ONLY the bf16 operand the convolution gathers (the fp32
 * activation is never materialised); the convolution entry then receives that buffer as `in` together with
 * scn_bf16_operand(m, in, in, 1) -- with a ready copy registered, `in` is an identity key and is not dereferenced.
 * Backward: scn_bn_bwd_fusion registers the BatchNorm (its input x, saved statistics, affine parameters) for the NEXT
 * scn_subm_bwd / scn_conv_bwd / scn_deconv_bwd on the handle.  That entry's dgrad epilogue then recomputes the activation
 * mask from x, writes the MASKED gradient d' into d_in, and accumulates acc[0][c] = sum d', acc[1][c] = sum d'*x (fp64,
 * [2][channels], zeroed by the entry) -- the reduction pass of the BatchNorm backward disappears.  scn_bn_bwd_apply finishes:
 * d_in = (d' - mean(d') - (x - mean)*k) * invstd*gamma (+ d_in_add), d_gamma, d_beta.  scn_bn_bwd_fusable says whether a
 * [c_in -> c_out] layer's dgrad can do this in `precision` (tensor-core shapes only); otherwise use scn_bn_bwd. */
int scn_bn_bwd_fusion(scn_meta *m, const float *bn_in, const float *save_mean, const float *save_invstd, const float *gamma,
                      const float *beta, float leakiness, double *acc);
int scn_bn_bwd_fusable(int c_in, int c_out, int precision);
int scn_bn_bwd_apply(const float *in, const float *d_masked, const double *acc, const float *save_mean, const float *save_invstd,
                     const float *gamma, const float *d_in_add, int64_t ld_add, float *d_in, void *d_in_bf16, float *d_gamma,
                     float *d_beta, int64_t n_rows, int channels, void *stream);
/* ld_add: row stride of d_in_add in floats (0 = dense).  d_in_bf16 (optional, [n_rows, channels] bf16): a copy of d_in for the
 * backward products of the convolution that receives d_in as its d_out -- registered there with scn_grad_bf16(m, d_out, copy)
 * before the NEXT scn_subm_bwd / scn_conv_bwd / scn_deconv_bwd on the handle (one use, dense rows), it replaces that entry's own
 * cast pass over d_out.
 * Row-strided gradients: a JoinTable's backward hands its consumers column slices of one [N, 2c] gradient.  scn_grad_stride
 * registers the row stride (floats) of the d_out argument of the NEXT scn_subm_bwd / scn_conv_bwd / scn_deconv_bwd on the handle,
 * so the slice is read where it lies (by the bf16 cast every product of that entry works from) instead of being copied out
 * first.  One use; needs the SCN_BF16 path for all of the entry's products (the entry fails otherwise) and no bias gradient. */
int scn_grad_stride(scn_meta *m, int64_t ld);
int scn_grad_bf16(scn_meta *m, const float *d_out, const void *d_out_bf16);
/* Column statistics from the strided layers: scn_out_stats registers a [2][Cout] fp64 buffer that the NEXT scn_conv_fwd /
 * scn_deconv_fwd on the handle fills with the column sums / sums of squares of its result (as the `stats` argument of scn_subm_fwd
 * does), for the BatchNorm that consumes it.  One use; tensor-core shapes only (scn_fuses_residual). */
int scn_out_stats(scn_meta *m, double *stats);

#ifdef __cplusplus
}
#endif
#endif /* SCN_B200_H */

/*
 * libia_b200 -- C ABI of the B200-native IntrinsicAvatar render path.
 *
 * The reference (taconite/IntrinsicAvatar) has no plugin ABI: its seams are a Python model class
 * and five pybind11 extension modules.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, a negative IA_E* code otherwise; ia_last_error() gives text;
 *   - nothing throws, nothing synchronises the device unless stated ("syncs");
 *   - `d_` arguments are DEVICE pointers owned by the caller, `h_` arguments are HOST pointers
 *     (small parameter blocks), `stream` is a cudaStream_t passed as void*;
 *   - a context is bound to one device and is not thread-safe; use one context per host thread/stream.
 */
#ifndef IA_B200_H
#define IA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IA_OK 0
#define IA_EINVAL -1
#define IA_ECUDA -2
#define IA_ESTATE -3
#define IA_EOVERFLOW -4

#define IA_N_COUNTERS 32

typedef struct ia_ctx ia_ctx;

const char* ia_last_error(void);
int ia_version(void);
/* Storage of the per-frame blended bone transform (voxel_J), fixed at build time: 1 = 32-byte voxels (deformed voxel
 * centre fp32 + rotation fp16; Broyden roots agree with the reference's broyden_kernel,
 * models/deformers/fast_snarf/cuda/fuse_kernel/fuse_cuda_kernel_fast.cu:250-413, to ~1e-5), 0 = 48-byte fp32 voxels (the
 * reference's values and order of operations: roots within 2e-6, > 90 % bit-identical).                              */
int ia_voxel_format(void);

/* lifetime ---------------------------------------------------------------------------------- */
int ia_create(ia_ctx** out, int device);
int ia_destroy(ia_ctx* ctx);

/* ---- model state -------------------------------------------------------------------------- */
/* tiny-cuda-nn HashGrid tables + folded MLP weights (replaces the nn.Module parameters of
 * models/rf/geometry.py:107-122, models/rf/radiance.py:82-110, models/pbr/material.py:13-30,
 * models/rf/density.py:20-34).  Hash tables are referenced, not copied (caller keeps them alive).
 * h_level_* : 16 entries each.  MLP matrices are row-major [out][in] fp32 host arrays.          */
int ia_set_fields(ia_ctx* ctx, const float* d_geo_hash, const float* d_rad_hash, int64_t n_entries,
                  const float* h_level_scale, const int32_t* h_level_res, const int32_t* h_level_size,
                  const int32_t* h_level_offset,
                  const float* h_geo_w1, const float* h_geo_b1, const float* h_geo_w2, const float* h_geo_b2,
                  const float* h_rad_w1, const float* h_rad_b1, const float* h_rad_w2, const float* h_rad_b2,
                  const float* h_rad_w3, const float* h_rad_b3,
                  const float* h_mat_w1, const float* h_mat_b1, const float* h_mat_w2, const float* h_mat_b2,
                  const float* h_mat_w3, const float* h_mat_b3,
                  const float* h_mat_scale5, const float* h_mat_bias5,
                  const float* h_bbox6 /* canonical bbox min,max: prepare_bbox, geometry.py:61-68 */,
                  float beta, void* stream);

/* Skinning-weight voxel grid, reference layout [24, D, H, W] fp32 (ForwardDeformer.lbs_voxel_final,
 * models/deformers/fast_snarf/deformer_torch.py:195); repacked channels-last into the context.     */
int ia_set_lbs_voxels(ia_ctx* ctx, const float* d_lbs_voxel, int D, int H, int W,
                      const float* h_offset_kernel3, const float* h_scale_kernel3, void* stream);

/* Subject set-up on the device.
 * ia_smpl_lbs: SMPL linear blend skinning for one pose -- lbs() (models/deformers/smplx/lbs.py:152-248: blend_shapes,
 *   vertices2joints, batch_rodrigues, pose blend shapes, batch_rigid_transform :345-401, skinning) and the translation
 *   handling of SMPL.forward (models/deformers/smplx/body_models.py:342-358).  Device arrays fp32: v_template [V,3],
 *   shapedirs [V,3,NB], posedirs [207, V*3] (the layout body_models.py:156-159 stores), J_regressor [24,V], lbs_weights
 *   [V,24]; host: parents [24], betas [NB], pose [72] = global_orient + body_pose (axis-angle), transl [3].
 *   Out (device): vertices [V,3], joints [24,3], A [24,4,4] (relative joint transforms, transl in the last column).
 * ia_voxelize_lbs: the skinning-weight voxel grid of ForwardDeformer.switch_to_explicit + query_weights_smpl
 *   (models/deformers/fast_snarf/deformer_torch.py:139-197, 234-253; pytorch3d knn_points, lib/pytorch3d/cuda/knn.cu:27-312):
 *   K = 30 nearest vertices of every grid point, inverse-distance blend, 30 smoothing passes.  d_verts [V,3] canonical
 *   vertices, d_weights [V,24]; out d_lbs_voxel [24, res/4, res, res] (the layout ia_set_lbs_voxels takes) and the
 *   reference's offset_kernel / scale_kernel [3] on the host.  Both calls synchronise the stream (set-up, not per sample). */
int ia_smpl_lbs(ia_ctx* ctx, const float* d_v_template, const float* d_shapedirs, const float* d_posedirs,
                const float* d_J_regressor, const float* d_lbs_weights, const int32_t* h_parents, int V, int NB,
                const float* h_betas, const float* h_pose72, const float* h_transl3, float* d_vertices, float* d_joints,
                float* d_A, void* stream);
int ia_voxelize_lbs(ia_ctx* ctx, const float* d_verts, const float* d_weights, int V, int resolution, float* d_lbs_voxel,
                    float* h_offset_kernel3, float* h_scale_kernel3, void* stream);

/* Per-frame pose: tfs [24,4,4], w2s [4,4] (SNARFDeformer.prepare_deformer, snarf_deformer.py:98-106)
 * and runs the voxel precompute (precompute_kernel, .../cuda/precompute/precompute.cu:22-103).       */
int ia_set_pose(ia_ctx* ctx, const float* h_tfs, const float* h_w2s, void* stream);

/* Render constants (configs/config.yaml:43-80).  albedo_align_ratio may be NULL (= 1).               */
int ia_set_render_config(ia_ctx* ctx, const float* h_scene_aabb6, int num_samples_per_ray,
                         int num_samples_per_secondary_ray, float secondary_near, float secondary_far,
                         float occ_thre, const float* h_background3, const float* h_albedo_align_ratio3);

/* config.model.secondary_importance_sample / zero_crossing_search (configs/config.yaml:53-54; models/intrinsic_avatar.py:
 * 482-520).  Defaults (1, 1): the secondary march stops at the first +/- crossing of the SDF and places 4 fine samples
 * there (ray_resampling_sdf_fine, lib/nerfacc/cuda/csrc/cdf.cu:536-638).  (1, 0): the 4 fine samples follow the CDF of the
 * compositing weights of all coarse samples (ray_resampling_fine, cdf.cu:403-478).  (0, *): no resampling, the coarse
 * samples are rendered at their midpoints.                                                                              */
int ia_set_secondary_sampling(ia_ctx* ctx, int importance_sample, int zero_crossing_search);

/* Capacity of the primary-sample pool of ia_render, in samples (default: 64 per ray of the call).  A frame that needs
 * more renders the rays that did not fit as background and counts them in IA_CNT_OVERFLOW; the host grows the pool with
 * this call and renders again (engine.RenderEngine.render(check_overflow=True)).  The reference has no such limit: its
 * per-chunk tensors are sized by the sample count (models/intrinsic_avatar.py:1232-1262).                                */
int ia_reserve_samples(ia_ctx* ctx, int64_t n_samples);

/* Test-time occupancy grid (IntrinsicAvatarModel.prepare_test_occupancy_grid /
 * _compute_occupancy_grid, models/intrinsic_avatar.py:307-381; max_connected_component,
 * models/utils.py:152-163).  d_jitter: [res^3,3,3] uniforms in [0,1) (the reference's rand_like).
 * d_binaries_out (optional, may be NULL): res^3 bytes.                                              */
int ia_build_occupancy(ia_ctx* ctx, const float* h_aabb6, int res, const float* d_jitter,
                       uint8_t* d_binaries_out, void* stream);
/* Training-time occupancy update (OccGridEstimator._update, models/occ_grid/temporal_occ_grid.py:369-411, driven by
 * IntrinsicAvatarModel.update_step, models/intrinsic_avatar.py:232-264): ONE jittered point per cell
 * (d_jitter [res^3,3]), occs = max(occs * ema_decay, alpha) on the caller's EMA state d_occs [res^3] of the frame's
 * level (updated in place), then max-pool / threshold min(mean, occ_thre) / largest component like the test-time
 * build; the result becomes the context's occupancy grid.  d_binaries_out optional.                     */
int ia_update_occupancy_ema(ia_ctx* ctx, const float* h_aabb6, int res, const float* d_jitter, float* d_occs,
                            float ema_decay, float occ_thre, uint8_t* d_binaries_out, void* stream);
/* Install a caller-provided grid instead (res^3 bytes, cell = (x*res+y)*res+z).                      */
int ia_set_occupancy(ia_ctx* ctx, const float* h_aabb6, int res, const uint8_t* d_binaries, void* stream);

/* Environment light (EnvironmentLightTensor.update_pdf/sample/pdf/eval, lib/torch_pbr/light.py:259-446).
 * d_envmap [H,W,3] fp32 (referenced); d_u1/d_u2 [spp] uniforms replacing torch.rand in sample().
 * Must be called after ia_set_pose (directions are rotated into the SMPL-root frame).
 * Optional outputs (may be NULL): d_dirs_world_out [spp,3], d_em_out [spp,3], d_pdf_out [spp].        */
int ia_set_light(ia_ctx* ctx, const float* d_envmap, int H, int W, const float* d_u1, const float* d_u2,
                 int spp, float* d_dirs_world_out, float* d_em_out, float* d_pdf_out, void* stream);

/* render_mode = uniform_light: the light table is the n_rows x n_cols stratified sphere of
 * EnvironmentLightBase.sample_uniform_sphere_stratified in eval mode (lib/torch_pbr/light.py:161-217;
 * the reference uses 16 x 32 and asserts samples_per_pixel == 512, models/intrinsic_avatar.py:1391);
 * samples_per_pixel becomes n_rows * n_cols.  Optional outputs: d_dirs_world_out, d_em_out [spp,3].     */
int ia_set_light_uniform(ia_ctx* ctx, const float* d_envmap, int H, int W, int n_rows, int n_cols,
                         float* d_dirs_world_out, float* d_em_out, void* stream);

/* ---- the hot path --------------------------------------------------------------------------- */
/* Output buffers of one forward pass, all DEVICE pointers, row-major [n_rays, C].  Any may be NULL.  */
typedef struct ia_outputs {
    float* comp_rgb;        /* [n,3] */
    float* comp_normal;     /* [n,3] */
    float* opacity;         /* [n,1] */
    float* depth;           /* [n,1] */
    float* comp_albedo;     /* [n,3] */
    float* comp_roughness;  /* [n,1] */
    float* comp_metallic;   /* [n,1] */
    float* comp_rgb_phys;   /* [n,3] linear */
    float* comp_demod_phys; /* [n,3] linear */
    int32_t* num_samples;   /* [n]   shading samples per ray */
    /* composited "_full" buffers (models/intrinsic_avatar.py:1625-1645) */
    float* comp_rgb_full;        /* [n,3] sRGB, clamped */
    float* comp_rgb_phys_full;   /* [n,3] sRGB, clamped */
    float* comp_demod_phys_full; /* [n,3] sRGB, clamped */
    float* comp_albedo_full;     /* [n,3] */
    float* comp_roughness_full;  /* [n,1] */
    float* comp_metallic_full;   /* [n,1] */
    float* visibility;           /* [n,1] render_mode = uniform_light only (models/intrinsic_avatar.py:1427-1432, 1516-1517) */
} ia_outputs;

#define IA_RENDER_PRIMARY_ONLY 1 /* stop after the primary volume render (albedo_only / config 2) */
#define IA_RENDER_GI 2           /* global_illumination = true: add one indirect bounce           */
/* config.model.render_mode (models/intrinsic_avatar.py:1346-1440), bits 2-3 of `flags`:                */
#define IA_RENDER_MODE_SHIFT 2
#define IA_RENDER_MODE_MASK (3 << IA_RENDER_MODE_SHIFT)
#define IA_RENDER_LIGHT (0 << IA_RENDER_MODE_SHIFT)         /* pbr_light_forward :755-861 (default)         */
#define IA_RENDER_UNIFORM_LIGHT (1 << IA_RENDER_MODE_SHIFT) /* pbr_uniform_light_forward :654-753           */
#define IA_RENDER_MATS (2 << IA_RENDER_MODE_SHIFT)          /* pbr_mats_forward :863-948 (BSDF sampling)     */
#define IA_RENDER_MIS (3 << IA_RENDER_MODE_SHIFT)           /* pbr_mis_forward :547-652 (BSDF + light, MIS)  */
#define IA_RENDER_ADD_EMITTER 16 /* config.model.add_emitter: the envmap along the primary ray replaces the
                                    background colour in comp_rgb_phys / comp_demod_phys (:1319-1341, 1454-1490) */

/* IntrinsicAvatarModel.forward in eval mode, render_mode = light | uniform_light | mats | mis
 * (models/intrinsic_avatar.py:950-1666; compute_indirect_radiance :396-545; pbr_*_forward :547-948;
 * models/volrend.py:810-1020; models/pbr/utils.py:70-229).  d_rays [n,8] world-space o,d,near,far.
 * Randomness: the per-ray light permutation of light / uniform_light (reference: CPU rand + argsort,
 * :1355-1378) is the stateless keyed permutation of (seed, ray_index_base + ray); the uniforms of
 * MultiLobe.sample / emitter.sample in mats / mis (reference: torch.rand) are counter-based:
 * stream (seed, ray_index_base + ray, shading sample j, dim) -- see ia_rng_uniform in csrc/ia_pbr.cuh.
 * mats / mis read the envmap given to the last ia_set_light call: it must still be valid.  Does not sync. */
int ia_render(ia_ctx* ctx, const float* d_rays, int64_t n_rays, int64_t ray_index_base, int flags,
              uint32_t seed, const ia_outputs* out, void* stream);

/* Counters of the last ia_render / ia_query call (syncs the stream): see IA_CNT_* below.             */
int ia_get_counters(ia_ctx* ctx, uint64_t* h_counters, void* stream);
#define IA_CNT_HIT_RAYS 0
#define IA_CNT_SAMPLES 1        /* primary shading samples */
#define IA_CNT_QUERIES 2        /* posed-point queries (SDF only) */
#define IA_CNT_QUERIES_GRAD 3   /* posed-point queries with gradient/feature */
#define IA_CNT_BROYDEN_FETCH 4  /* voxel_J trilinear fetches */
#define IA_CNT_GEO_EVAL 5       /* canonical geometry evaluations (hash grid + MLP) */
#define IA_CNT_RAD_EVAL 6       /* radiance (+material) evaluations */
#define IA_CNT_SECONDARY_RAYS 7
#define IA_CNT_OVERFLOW 8       /* rays that exceeded the per-ray edge capacity / sample pool */
#define IA_CNT_SKIN_FETCH 9     /* 24-channel skinning-weight fetches */
#define IA_CNT_CHAINS_SKIPPED 10 /* Broyden chains whose initial point lies outside the voxel grid (exactly invalid) */
/* h_counters[IA_CNT_PRIMARY_BASE + i] = counter i as it stood when the primary stage of the last ia_render
 * had finished, so (total - primary) is the work of the shading kernel alone (roofline accounting).   */
#define IA_CNT_PRIMARY_BASE 16

/* Per-stage device timing (CUDA events recorded on the launching stream around each stage's kernels).
 * ia_set_timing(ctx, 1) enables recording; ia_get_timings syncs the stream and returns, for each
 * IA_STAGE_*, the elapsed ms of the most recent execution (-1 if the stage did not run), and the
 * total number of kernels launched by this context since creation.                                  */
#define IA_N_STAGES 8
#define IA_STAGE_PRECOMPUTE 0
#define IA_STAGE_OCCUPANCY 1
#define IA_STAGE_LIGHT 2
#define IA_STAGE_SETUP 3
#define IA_STAGE_PRIMARY 4
#define IA_STAGE_RESAMPLE 5
#define IA_STAGE_SHADE 6
#define IA_STAGE_COMPOSITE 7
int ia_set_timing(ia_ctx* ctx, int enable);
int ia_get_timings(ia_ctx* ctx, float* h_ms /* [IA_N_STAGES] */, uint64_t* h_launches, void* stream);

/* ---- op-level entry points (A/B tests against the reference's pybind modules) --------------- */
/* precompute.precompute(voxel_w, tfs, voxel_d, voxel_J, offset, scale)  precompute.cpp:6-13.
 * d_voxel_J_out: reference layout [12, D, H, W].                                                      */
int ia_op_precompute(ia_ctx* ctx, float* d_voxel_J_out, void* stream);

/* fuse_cuda.fuse_broyden(...) + filter.filter(x, mask)  fuse_cuda.cpp:14-28, filter.cpp:12-22.
 * d_xd [n,3] -> d_x [n,13,3], d_J_inv [n,13,3,3] (may be NULL), d_valid_raw [n,13] (before filter,
 * may be NULL), d_valid [n,13] (after filter).                                                        */
int ia_op_broyden(ia_ctx* ctx, const float* d_xd, int64_t n, float* d_x, float* d_J_inv,
                  uint8_t* d_valid_raw, uint8_t* d_valid, void* stream);

/* SNARFDeformer.deform + VolumeSDF.forward fused (snarf_deformer.py:187-261; geometry.py:124-172).
 * Outputs (any may be NULL): d_sdf [n], d_xc [n,3], d_valid [n]; with_grad: d_grad [n,3] (posed),
 * d_grad_cano [n,3], d_feature [n,13].                                                                */
int ia_op_query(ia_ctx* ctx, const float* d_xd, int64_t n, int with_grad, float* d_sdf, float* d_xc,
                uint8_t* d_valid, float* d_grad, float* d_grad_cano, float* d_feature, void* stream);

/* Training-mode building block: backward of VolumeSDF's network (models/rf/geometry.py:124-172:
 * hash-grid encoding + VanillaMLP 35 -> 64 -> 13, models/network_utils.py:58-79, 201-244; the reference differentiates it
 * with autograd through tiny-cuda-nn).  d_xc [n,3] canonical points, d_dout [n,13] upstream gradient of the 13 outputs
 * (channel 0 = sdf).  ADDS into: d_g_hash [2 * n_entries] (gradient of the geometry hash table, the layout of the table),
 * d_g_mlp [IA_GEO_END floats: W1^T [35][64] | b1 [64] | W2 [13][64] | b2 [16]] (gradient of the effective, weight-norm
 * folded weights; 3152 floats); writes d_g_x [n,3] (may be NULL): gradient with respect to the canonical position.   */
#define IA_GEO_GRAD_FLOATS 3152
int ia_op_geometry_backward(ia_ctx* ctx, const float* d_xc, const float* d_dout, int64_t n, float* d_g_hash, float* d_g_mlp,
                            float* d_g_x, void* stream);

/* Training-mode building block: backward of the implicit-differentiation correction of the Broyden roots,
 * ForwardDeformer.forward version 1 (models/deformers/fast_snarf/deformer_torch.py:57-76; forward_skinning / skinning_mask
 * :127-137, 213-227; query_weights :199-210) -- the reference builds  x_c = x_c* - J_inv (x_d(x_c*) - x_d(x_c*).detach())  and
 * lets autograd carry dL/dx_c to the bone transforms.  Inputs as ia_op_broyden returns them: d_xc [n,13,3] roots, d_valid [n,13]
 * (after filter), d_J_inv [n,13,3,3]; d_g_xc [n,13,3] upstream gradient (ia_op_geometry_backward's d_g_x of the valid roots).
 * ADDS into d_g_tfs [24][3][4]: the gradient with respect to rows 0..2 of the 24 bone transforms set by ia_set_pose.       */
int ia_op_deform_backward(ia_ctx* ctx, const float* d_xc, const uint8_t* d_valid, const float* d_J_inv, const float* d_g_xc,
                          int64_t n, float* d_g_tfs, void* stream);

/* Training-mode building block: backward of the two shading networks the render path evaluates with
 * ia_op_shade_fields -- VolumeRadiance.forward (models/rf/radiance.py:111-135: radiance hash grid + SH of the reflected view
 * direction + VanillaMLP 67 -> 64 -> 64 -> 3, sigmoid) and the material network (models/pbr/material.py:31-51: 48 -> 64 -> 64 -> 5,
 * sigmoid * scale (* albedo_align_ratio) + bias; models/network_utils.py:201-244, 360-428) -- which the reference differentiates
 * with autograd through tiny-cuda-nn.  Inputs as ia_op_shade_fields; d_drgb [n,3], d_dmat [n,5] upstream gradients of its outputs.
 * ADDS into d_g_rad_hash [2 * n_entries] (the radiance table's layout) and d_g_mlp [IA_SHADE_GRAD_FLOATS: radiance W1^T [67][64] |
 * b1 [64] | W2^T [64][64] | b2 [64] | W3 [3][64] | b3 [4] | material W1^T [48][64] | b1 | W2^T | b2 | W3 [5][64] | b3 [8]]
 * (effective, folded weights); writes (any may be NULL) d_g_x [n,3] (canonical position), d_g_feature [n,13] (feed it to
 * ia_op_query_backward as d_dout), d_g_normal [n,3] (world-space normal; the view direction is data).                       */
#define IA_SHADE_GRAD_FLOATS 16332
int ia_op_shade_fields_backward(ia_ctx* ctx, const float* d_xc, const float* d_feature, const float* d_view,
                                const float* d_normal, const float* d_drgb, const float* d_dmat, int64_t n,
                                float* d_g_rad_hash, float* d_g_mlp, float* d_g_x, float* d_g_feature, float* d_g_normal,
                                void* stream);

/* Training-mode building block: compositing along the primary rays and its backward -- what
 * rendering_with_normals_mats_sdf (models/volrend.py:336-364) builds from get_alpha (Laplace-CDF density, models/rf/density.py:17-34;
 * models/intrinsic_avatar.py:390-394), nerfacc 0.5.3 render_weight_from_alpha and accumulate_along_rays, and differentiates with
 * autograd.  d_packed_info [n_rays,2] = (first sample, count) like nerfacc; d_sdf / d_dists [n_samples]; d_values [n_samples, C]
 * (C <= IA_VOLREND_MAX_C: rgb, normal, materials ... concatenated by the caller); beta = LearnedLaplaceDensity.get_beta().
 * ia_op_volrend: d_weights [n_samples] (may be NULL), d_comp [n_rays, C], d_opacity [n_rays].
 * ia_op_volrend_backward: upstream d_dcomp [n_rays, C], d_dopacity [n_rays] (may be NULL), d_dweights [n_samples] (may be NULL:
 * the gradient on the weights themselves -- the physically based branch re-uses them for its shading samples,
 * models/pbr/utils.py:146-161) -> writes d_g_sdf [n_samples], d_g_values [n_samples, C] (may be NULL); ADDS into d_g_beta [1]. */
#define IA_VOLREND_MAX_C 16
int ia_op_volrend(ia_ctx* ctx, const int32_t* d_packed_info, const float* d_sdf, const float* d_dists, const float* d_values,
                  int n_channels, float beta, int64_t n_rays, float* d_weights, float* d_comp, float* d_opacity, void* stream);
int ia_op_volrend_backward(ia_ctx* ctx, const int32_t* d_packed_info, const float* d_sdf, const float* d_dists,
                           const float* d_values, int n_channels, float beta, const float* d_dcomp, const float* d_dopacity,
                           const float* d_dweights, int64_t n_rays, float* d_g_sdf, float* d_g_values, float* d_g_beta,
                           void* stream);

/* Training-mode forward / backward of the fused query: SNARFDeformer.deform with eval_mode=False
 * (models/deformers/snarf_deformer.py:170-261) = ForwardDeformer.forward version 1 (search + implicit-differentiation
 * correction, deformer_torch.py:34-76) -> VolumeSDF at every kept root -> min over the roots.
 * ia_op_query_train: the values ia_op_query(with_grad=1) returns (the correction is zero-valued) plus what the backward pass
 * reads: d_J_inv [n,3,3] = others['J_inv'] of the arg-min root (zeros without a root), d_best [n] = its init-bone slot 0..12.
 * d_grad / d_grad_cano / d_feature may be NULL.
 * ia_op_query_backward: for an upstream gradient d_dout [n,13] on the 13 network outputs at the arg-min root (channel 0 = sdf;
 * torch.min routes the gradient to that root only; a query without a root has the constant sdf 1e5 and no gradient):
 * ADDS into d_g_hash / d_g_mlp (layouts of ia_op_geometry_backward) and d_g_tfs [24][3][4] (ia_op_deform_backward);
 * writes d_g_x [n,3] = dL/dx_c (zeros for queries without a root).  The reference gets all of this from autograd.           */
int ia_op_query_train(ia_ctx* ctx, const float* d_xd, int64_t n, float* d_sdf, float* d_xc, uint8_t* d_valid, float* d_grad,
                      float* d_grad_cano, float* d_feature, float* d_J_inv, int32_t* d_best, void* stream);
int ia_op_query_backward(ia_ctx* ctx, const float* d_xc, const uint8_t* d_valid, const float* d_J_inv, const float* d_dout,
                         int64_t n, float* d_g_hash, float* d_g_mlp, float* d_g_tfs, float* d_g_x, void* stream);

/* Canonical SDF of n points, evaluated the way the wavefront integrator's geometry phase does: hash grid, then the
 * 35 -> 64 layer as warp-level tensor-core mma (TF32 inputs split in two, fp32 accumulate), softplus(beta = 100), sdf row
 * of the output layer.  Replaces VolumeSDF.forward without gradient (models/rf/geometry.py:124-146: encoding ->
 * VanillaMLP, models/network_utils.py:201-244) for A/B tests against the fp32 evaluation.  d_xc [n,3] -> d_sdf [n].   */
int ia_op_geometry(ia_ctx* ctx, const float* d_xc, int64_t n, float* d_sdf, void* stream);

/* radiance + material at canonical points (radiance.py:111-135, material.py:31-51):
 * d_xc [n,3], d_feature [n,13], d_view_world [n,3], d_normal_world [n,3] -> d_rgb [n,3], d_mat [n,5]. */
int ia_op_shade_fields(ia_ctx* ctx, const float* d_xc, const float* d_feature, const float* d_view_world,
                       const float* d_normal_world, int64_t n, float* d_rgb, float* d_mat, void* stream);

/* nerfacc traverse_grids as used by sampling_override (models/intrinsic_avatar.py:49-141), on the
 * context's occupancy grid.  Two calls: d_vals == NULL counts (d_n_edges, d_n_samples [n]); else writes
 * at the caller-computed exclusive prefix offsets d_edge_base / d_sample_base.                        */
int ia_op_traverse(ia_ctx* ctx, const float* d_rays_o, const float* d_rays_d, int64_t n, float near_plane,
                   float far_plane, float step, int32_t* d_n_edges, int32_t* d_n_samples,
                   const int32_t* d_edge_base, const int32_t* d_sample_base, float* d_vals,
                   uint8_t* d_is_left, uint8_t* d_is_right, float* d_t_starts, float* d_t_ends, void* stream);

/* lib.nerfacc ray_resampling (cdf.cu:151-215): packed_info [n_rays,2]; d_resample_packed_info
 * [n_rays,2] is an INPUT (exclusive prefix of (steps>0)*spp, spp) computed by the caller.             */
int ia_op_ray_resampling(ia_ctx* ctx, const int32_t* d_packed_info, const float* d_starts,
                         const float* d_ends, const float* d_weights, const float* d_sdfs, int64_t n_rays,
                         int spp, const int32_t* d_resample_packed_info, float* d_ts, float* d_offsets,
                         int64_t* d_indices, int32_t* d_fg_counts, int32_t* d_bg_counts,
                         int64_t* d_surface_idx, void* stream);
/* lib.nerfacc ray_resampling_merge (cdf.cu:336-401); outputs pre-zeroed by the caller.                */
int ia_op_ray_resampling_merge(ia_ctx* ctx, const int32_t* d_packed_info, const float* d_vals,
                               const uint8_t* d_is_left, const uint8_t* d_is_right, const float* d_weights,
                               int64_t n_rays, const int32_t* d_resample_packed_info, float* d_vals_out,
                               float* d_dists_out, uint8_t* d_is_left_out, uint8_t* d_is_right_out,
                               uint8_t* d_is_resample_out, uint8_t* d_is_fg_out, void* stream);
/* lib.nerfacc ray_resampling_fine (cdf_resampling_fine_kernel, cdf.cu:403-534): n fine intervals along the CDF of the
 * compositing weights; outputs pre-zeroed by the caller.                                                 */
int ia_op_ray_resampling_fine(ia_ctx* ctx, const int32_t* d_packed_info, const float* d_starts, const float* d_ends,
                              const float* d_weights, int64_t n_rays, const int32_t* d_resample_packed_info,
                              float* d_starts_out, float* d_ends_out, uint8_t* d_is_fg_out, void* stream);
/* lib.nerfacc ray_resampling_sdf_fine (cdf.cu:640-696); outputs pre-zeroed by the caller.             */
int ia_op_ray_resampling_sdf_fine(ia_ctx* ctx, const int32_t* d_packed_info, const float* d_starts,
                                  const float* d_ends, const float* d_alphas, const float* d_sdfs,
                                  int64_t n_rays, const int32_t* d_resample_packed_info, float* d_starts_out,
                                  float* d_ends_out, uint8_t* d_is_fg_out, void* stream);
/* lib.nerfacc unpack_info (pack.cu:84-107)                                                            */
int ia_op_unpack_info(ia_ctx* ctx, const int32_t* d_packed_info, int64_t n_rays, int64_t* d_ray_indices,
                      void* stream);
/* secondary-ray transmittance / indirect radiance (compute_indirect_radiance,
 * models/intrinsic_avatar.py:396-545): d_o, d_d [n,3] in the SMPL-root frame -> d_T [n], d_rgb [n,3]. */
int ia_op_secondary(ia_ctx* ctx, const float* d_o, const float* d_d, int64_t n, int gi, float* d_T,
                    float* d_rgb, void* stream);
/* MultiLobe.eval (lib/torch_pbr/bxdf.py:321-330): wi, n, wo [n,3], rough [n], albedo [n,3], metal [n]
 * -> diff [n], spec [n,3] (both include the cosine).                                                  */
int ia_op_brdf(ia_ctx* ctx, const float* d_wi, const float* d_n, const float* d_wo, const float* d_rough,
               const float* d_albedo, const float* d_metal, int64_t n, float* d_diff, float* d_spec,
               void* stream);
/* MultiLobe.sample with explicit uniforms d_sample [n,2] (bxdf.py:332-388) -> d_wo [n,3]; and
 * MultiLobe.pdf (bxdf.py:290-317) of d_wo_query [n,3] -> d_pdf [n].  Either output may be NULL.         */
int ia_op_bsdf_sample_pdf(ia_ctx* ctx, const float* d_wi, const float* d_n, const float* d_rough,
                          const float* d_albedo, const float* d_metal, const float* d_sample,
                          const float* d_wo_query, int64_t n, float* d_wo, float* d_pdf, void* stream);
/* Training-mode building block: the differentiable part of the physically based integrators.  The reference
 * traces the secondary rays under torch.no_grad() (models/intrinsic_avatar.py:673-706 pbr_uniform_light_forward -- the training
 * default, configs/config.yaml:46 --, :575-640, :763-800, :880-896), so light direction d_wo, inverse pdf, and
 * Li = em_li * transmittance + indirect radiance enter as per-sample arrays; what autograd walks is MultiLobe.eval
 * (lib/torch_pbr/bxdf.py:111-146, 217-265, 321-330), Lo_diff = Li diff inv_pdf, Lo_spec = Li spec inv_pdf,
 * Lo = (1 - metallic) albedo Lo_diff + Lo_spec (:736-751), with diff = spec = 0 where n.wo <= 1e-6 (:690).
 * All arrays are per shading sample: d_wi (towards the viewer), d_n (unit normal), d_wo, d_albedo, d_Li [n,3]; d_rough, d_metal,
 * d_inv_pdf [n].  ia_op_pbr_shade writes d_Lo and (unless NULL) d_Lo_diff, d_Lo_spec [n,3].  ia_op_pbr_shade_backward takes the
 * upstream gradients d_dLo and (may be NULL) d_dLo_diff, d_dLo_spec and writes d_g_n [n,3] (defined up to a component along n,
 * which the backward of the normalisation that produced n removes), d_g_rough [n], d_g_albedo [n,3], d_g_metal [n], d_g_Li [n,3]. */
int ia_op_pbr_shade(ia_ctx* ctx, const float* d_wi, const float* d_n, const float* d_wo, const float* d_rough,
                    const float* d_albedo, const float* d_metal, const float* d_Li, const float* d_inv_pdf, int64_t n,
                    float* d_Lo, float* d_Lo_diff, float* d_Lo_spec, void* stream);
int ia_op_pbr_shade_backward(ia_ctx* ctx, const float* d_wi, const float* d_n, const float* d_wo, const float* d_rough,
                             const float* d_albedo, const float* d_metal, const float* d_Li, const float* d_inv_pdf,
                             const float* d_dLo, const float* d_dLo_diff, const float* d_dLo_spec, int64_t n, float* d_g_n,
                             float* d_g_rough, float* d_g_albedo, float* d_g_metal, float* d_g_Li, void* stream);
/* Backward of EnvironmentLightTensor.eval (lib/torch_pbr/light.py:298-339, bilinear grid_sample, align_corners, border): ADDS
 * the gradient d_dem [n,3] on the emission of the WORLD directions d_dirs_world [n,3] into d_g_env [H,W,3], the texels of the
 * map of the last ia_set_light* call (the trainable light's softplus stays with the caller).                                  */
int ia_op_env_backward(ia_ctx* ctx, const float* d_dirs_world, const float* d_dem, int64_t n, float* d_g_env, void* stream);

/* EnvironmentLightTensor.sample / pdf / eval per direction, on the tables of the last ia_set_light*
 * call (light.py:259-446): d_u [n,2] -> d_dirs_world_out [n,3] (may be NULL); d_dirs_world [n,3] ->
 * d_pdf_out [n], d_em_out [n,3] (d_dirs_world NULL: the directions just sampled are used).             */
int ia_op_env(ia_ctx* ctx, const float* d_u, const float* d_dirs_world, int64_t n, float* d_dirs_world_out,
              float* d_pdf_out, float* d_em_out, void* stream);

/* ---- frame producer / consumer either side of the path ------------------------ */
/* AnimationDataset's per-frame ray tensor built on the device (datasets/animation.py:13-34 make_rays,
 * :29-33 transform_rays, :163-189 __getitem__; systems/intrinsic_avatar.py:100-109): d_rays [H*W,8] =
 * o, d, near, far.  h_Kinv9 = inv(K) row-major (float64, as numpy computes it); h_c2w12 = dataset camera
 * 3x4 (NULL = identity); h_ext12 = inv(w2c) of the frame 3x4 (NULL = identity).  Does not sync (the
 * matrices travel as a kernel argument).                                                             */
int ia_make_rays(ia_ctx* ctx, const double* h_Kinv9, const double* h_c2w12, const double* h_ext12, int H, int W,
                 float near_plane, float far_plane, float* d_rays, void* stream);
/* SaverMixin.get_rgb_image_ (utils/mixins.py:43-53): clip to [lo,hi] -> 0..255 -> uint8 (truncating),
 * optional RGB->BGR for cv2.imwrite.  d_img [n_pix, channels] fp32 -> d_out [n_pix, channels] uint8.     */
int ia_pack_rgb8(ia_ctx* ctx, const float* d_img, int64_t n_pix, int channels, float lo, float hi, int bgr,
                 uint8_t* d_out, void* stream);
/* One column of SaverMixin.get_image_grid_ / save_image_grid (utils/mixins.py:116-157) written into the uint8 grid
 * d_grid [H][grid_w][3] at pixel column x0.  kind 0 = 'rgb' (get_rgb_image_, :43-53; 1-2 channels zero-padded),
 * kind 1 = 'grayscale' (get_grayscale_image_, :87-101: nan_to_num, clip/scale, colour map d_lut [256][3] or NULL =
 * cmap None).  d_range (device, 2 floats) overrides lo/hi (data_range=None: min/max of the image).  The grid is in
 * the channel order of the file (RGB); bgr=1 writes BGR for a cv2.imwrite caller.                            */
int ia_pack_grid8(ia_ctx* ctx, const float* d_img, int H, int W, int channels, int kind, float lo, float hi,
                  const float* d_range, const uint8_t* d_lut, uint8_t* d_grid, int grid_w, int x0, int bgr, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IA_B200_H */

// Shared device-side types for the IntrinsicAvatar render path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace cg = cooperative_groups;

#define IA_N_BONES 24
#define IA_N_INIT 13
#define IA_N_LEVELS 16
#define IA_TEAM 16  // lanes that cooperate on one posed point: 13 Broyden inits / 16 hash levels
#define IA_CAP 256  // max edges (and samples) per primary ray
// voxel_J storage (DESIGN.md "Broyden fetch"):
//   IA_VOXEL32 = 0: 48-byte voxels, the blended 3x4 transform in fp32, 3 x LDG.128 per trilinear corner.  Bit-compatible with
//      the reference's broyden_kernel (same values, same order of operations).
//   IA_VOXEL32 = 1: 32-byte voxels = ONE aligned sector, one LDG.256 per corner: the deformed voxel centre
//      y_c = R_c c_c + t_c in fp32 (12 B) + the blended rotation R_c in fp16 (18 B) + 2 B pad.  The fetch evaluates
//      sum_c w_c (y_c + R_c (x - c_c)), so the fp16 error multiplies |x - c_c| <= one voxel (2 cm) instead of |x| ~ 1 m:
//      frame buffers within 4e-5 relative L2 of fp32 storage (scripts/voxel_precision_study.py; plain fp16 storage of the
//      3x4: 2e-2).  Roots agree with the reference kernel to ~1e-5 instead of 2e-6.  The L1 charges a scattered gather per
//      load instruction (scripts/gather_microbench.cu: 18.7 G fetches/s with 3 x LDG.128 per corner -- the rate the Broyden
//      phase of round 1 ran at -- vs 38.0 with one LDG.256).
#ifndef IA_VOXEL32
#define IA_VOXEL32 0
#endif
#if IA_VOXEL32
#define IA_VOXEL_F4 2  // float4 per voxel of voxel_J
#else
#define IA_VOXEL_F4 3
#endif

// Offsets (in floats) inside the packed MLP weight blob. "T" = stored input-major [in][64].
#define IA_GEO_W1T 0                          // [35][64]
#define IA_GEO_B1 (IA_GEO_W1T + 35 * 64)      // [64]
#define IA_GEO_W2 (IA_GEO_B1 + 64)            // [13][64]
#define IA_GEO_B2 (IA_GEO_W2 + 13 * 64)       // [16] (13 used)
#define IA_GEO_END (IA_GEO_B2 + 16)
#define IA_RAD_W1T IA_GEO_END                 // [67][64]
#define IA_RAD_B1 (IA_RAD_W1T + 67 * 64)
#define IA_RAD_W2T (IA_RAD_B1 + 64)           // [64][64]
#define IA_RAD_B2 (IA_RAD_W2T + 64 * 64)
#define IA_RAD_W3 (IA_RAD_B2 + 64)            // [3][64]
#define IA_RAD_B3 (IA_RAD_W3 + 3 * 64)        // [4]
#define IA_RAD_END (IA_RAD_B3 + 4)
#define IA_MAT_W1T IA_RAD_END                 // [48][64]
#define IA_MAT_B1 (IA_MAT_W1T + 48 * 64)
#define IA_MAT_W2T (IA_MAT_B1 + 64)           // [64][64]
#define IA_MAT_B2 (IA_MAT_W2T + 64 * 64)
#define IA_MAT_W3 (IA_MAT_B2 + 64)            // [5][64]
#define IA_MAT_B3 (IA_MAT_W3 + 5 * 64)        // [8]
#define IA_MLP_END (IA_MAT_B3 + 8)
// Behind the fp32 blob: mma.sync B fragments of the layers the tensor-core shading paths use (ia_mma.cuh), one float4
// {b0_hi, b0_lo, b1_hi, b1_lo} (TF32 split) per lane, [k-step][n-tile][32 lanes] per layer; built on the host by ia_set_fields.
#define IA_FRAG_FEAT 0                          // geometry 64 -> 13 (8 k-steps x 2 n-tiles), A from C fragments
#define IA_FRAG_BWD (IA_FRAG_FEAT + 8 * 2)      // d sdf / d input: hidden 64 -> 35 tile columns (8 x 5), A from C fragments
#define IA_FRAG_RAD1 (IA_FRAG_BWD + 8 * 5)      // radiance 67 -> 64 (9 x 8), A from the shading tile
#define IA_FRAG_RAD2 (IA_FRAG_RAD1 + 9 * 8)     // radiance 64 -> 64 (8 x 8), A from C fragments
#define IA_FRAG_RAD3 (IA_FRAG_RAD2 + 8 * 8)     // radiance 64 -> 3 (8 x 1)
#define IA_FRAG_MAT1 (IA_FRAG_RAD3 + 8)         // material 48 -> 64 (6 x 8), A from the shading tile
#define IA_FRAG_MAT2 (IA_FRAG_MAT1 + 6 * 8)     // material 64 -> 64
#define IA_FRAG_MAT3 (IA_FRAG_MAT2 + 8 * 8)     // material 64 -> 5 (8 x 1)
#define IA_FRAG_GEO1 (IA_FRAG_MAT3 + 8)         // geometry 35 -> 64 (5 x 8), A from the geometry / shading tile
#define IA_FRAG_END (IA_FRAG_GEO1 + 5 * 8)
#define IA_BLOB_FLOATS (IA_MLP_END + IA_FRAG_END * 128)

#define IA_SEC_ZERO_CROSSING 0
#define IA_SEC_IMPORTANCE 1
#define IA_SEC_PLAIN 2

// Per-frame constants, passed to every kernel by value (__grid_constant__).
struct IaFrame {
    // --- fast-SNARF
    float tfs[IA_N_BONES][12];  // rows 0..2 of each bone transform, row-major 3x4
    float w2s[12];              // world -> SMPL-root, row-major 3x4
    int init_bones[IA_N_INIT];
    float off[3], scl[3];       // reference offset_kernel / scale_kernel
    int D, H, W;                // LBS voxel grid (32,128,128)
    float vox_h[3], vox_b[3];   // canonical position of voxel (xi, yi, zi) = vox_b + (xi, yi, zi) * vox_h (x, y, z order)
    const float4* voxel_J;      // blended 3x4 per voxel, channels-last; layout per IA_VOXEL32
    const float4* lbs_w;        // [D*H*W][6] float4  : 24 skinning weights per voxel, channels-last
    // --- canonical fields
    const float2* geo_hash;
    const float2* rad_hash;
    float lvl_scale[IA_N_LEVELS];
    uint32_t lvl_res[IA_N_LEVELS], lvl_size[IA_N_LEVELS], lvl_off[IA_N_LEVELS];
    float center[3], scale[3];  // canonical bbox centre / extent (geometry.py:61-68)
    float beta;                 // Laplace density scale (already abs()+beta_min)
    const float* mlp;           // packed weight blob (IA_* offsets)
    float mat_scale[5], mat_bias[5];
    float albedo_ratio[3];
    // --- occupancy grid (test-time)
    const uint32_t* occ_bits;   // [res^3/32], cell = (x*res + y)*res + z
    int occ_res;
    float aabb[6];
    float step_primary;         // diag(scene_aabb) / num_samples_per_ray
    float sec_near, sec_far, sec_step;
    int sec_mode;               // secondary rays: 0 importance sampling + zero-crossing search (default), 1 importance sampling
                                // from the ray start (zero_crossing_search = false), 2 no importance sampling (IA_SEC_*)
    float background[3];
};

// One shading sample of a primary ray (written by the primary kernel, read by the PBR kernels).
struct IaSample {
    float ts, te, w, sdf;
    float n[3];
    float albedo[3];
    float rough, metal;
};
static_assert(sizeof(IaSample) == 48, "IaSample layout");
struct IaSampleAux {
    float rgb[3];
    float nw[3];   // world-space unit normal
    int slot;      // hit-ray slot of the sample
    int pad;
};
static_assert(sizeof(IaSampleAux) == 32, "IaSampleAux layout");

typedef cg::thread_block_tile<IA_TEAM> Team;

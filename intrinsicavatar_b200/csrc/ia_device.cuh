// Device functions of the fused posed-point query: Broyden x13 -> filter -> hash grid + MLP ->
// arg-min SDF, evaluated by a 16-lane team with no intermediate in HBM.
//
// Replaces (reference file:line):
//   broyden_kernel + grid_sampler_3d   models/deformers/fast_snarf/cuda/fuse_kernel/fuse_cuda_kernel_fast.cu:110-413
//   filter                              models/deformers/fast_snarf/cuda/filter/filter.cu:10-54
//   forward_skinning (fwd_tfs)          models/deformers/fast_snarf/deformer_torch.py:46-55,127-137,199-227
//   SNARFDeformer.deform                models/deformers/snarf_deformer.py:187-261
//   VolumeSDF.forward (+tcnn HashGrid, VanillaMLP, autograd d sdf/dx)  models/rf/geometry.py:124-172
//   VolumeRefDirRadiance.forward (+tcnn SH4)   models/rf/radiance.py:111-135
//   VolumeMaterial.forward (+LipshitzMLP)      models/pbr/material.py:31-51
#pragma once
#include <cuda_fp16.h>
#include "ia_types.cuh"
#include "ia_mma.cuh"

#define IA_FULL_TEAM 0xFFFFu

// ------------------------------------------------------------------------------------------------
// small math helpers
// torch.nn.Softplus(beta=100, threshold=20) (models/network_utils.py:201-244 via get_activation).
// IA_SOFTPLUS_FAST = 0: the reference's own expression log1p(exp(beta x)) / beta with libdevice expf / log1pf / IEEE
//   division (~40 instructions; 4 per lane and geometry evaluation = a third of the geometry phase).
// IA_SOFTPLUS_FAST = 1 (default): the algebraically identical max(x, 0) + log1p(exp(-|beta x|)) / beta on the
//   MUFU ex2 / lg2 units (~8 instructions).  The correction term is <= ln 2 / 100, so the absolute error of the
//   approximate exp / log (<= 2e-7) enters the result as <= 2e-9 -- below the one-ulp (~7e-9 at 0.1) rounding noise
//   the reference's expression has itself.
#ifndef IA_SOFTPLUS_FAST
#define IA_SOFTPLUS_FAST 1
#endif
__device__ __forceinline__ float ia_softplus100(float x) {
    float bx = 100.0f * x;
#if IA_SOFTPLUS_FAST
    if (bx > 20.0f) return x;
    return fmaxf(x, 0.f) + __logf(1.0f + __expf(-fabsf(bx))) * 0.01f;
#else
    return bx > 20.0f ? x : log1pf(expf(bx)) / 100.0f;
#endif
}
__device__ __forceinline__ float ia_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// Laplace-CDF density then alpha (models/rf/density.py:25-30; models/intrinsic_avatar.py:390-394)
__device__ __forceinline__ float ia_alpha(float sdf, float dist, float beta) {
    float sgn = (sdf > 0.0f) ? 1.0f : ((sdf < 0.0f) ? -1.0f : 0.0f);
    float sigma = (1.0f / beta) * (0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) / beta));
    return 1.0f - expf(-sigma * dist);
}

// ------------------------------------------------------------------------------------------------
// Trilinear fetch of the blended 3x4 bone transform, zero padding, align_corners=True
// (fuse_cuda_kernel_fast.cu:110-232).  Channels-last: one voxel = 3 x float4 = 48 contiguous bytes
// (the reference's channel-major layout costs 12 scattered 4-byte loads per corner).
__device__ __forceinline__ float ia_src_index(float coord, int size) {
    float x = ((coord + 1.f) / 2) * (size - 1);
    if (x > 2147483646.0f || x < -2147483648.0f || !isfinite(x)) x = -100.0f;
    return x;
}

// 256-bit read-only global load (sm_100: LDG.E.256): one 32-byte half voxel per instruction
__device__ __forceinline__ void ia_ld256(const float4* ptr, float r[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "l"(ptr));
}

// Corner set-up of a trilinear fetch.
struct IaCorners {
    int x0, y0, z0;
    float wx0, wx1, wy0, wy1, wz0, wz1;
};
__device__ __forceinline__ IaCorners ia_corners(const IaFrame& p, float gx, float gy, float gz) {
    IaCorners c;
    float ix = ia_src_index(gx, p.W), iy = ia_src_index(gy, p.H), iz = ia_src_index(gz, p.D);
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    c.x0 = (int)fx; c.y0 = (int)fy; c.z0 = (int)fz;
    c.wx1 = ix - fx; c.wy1 = iy - fy; c.wz1 = iz - fz;
    c.wx0 = (float)(c.x0 + 1) - ix; c.wy0 = (float)(c.y0 + 1) - iy; c.wz0 = (float)(c.z0 + 1) - iz;
    return c;
}
// True iff all 8 trilinear corners of the (normalised) point lie outside the voxel grid, i.e. the zero-padded
// fetch returns J == 0 exactly.  A Broyden chain whose INITIAL point is such a point can never become valid:
// Ji = 0 and g = -x_d, so the first step is u = 0, the fetch repeats with J = 0 and n = -x_d; then either
// |x_d|^2 < 1e-10 ("converged", but the point is outside [-1,1]^3 -> invalid), or > 1e-2 (diverged), or the
// rank-1 update divides 0 by 0 and every later residual is NaN (never converges).  Skipping such chains is
// therefore exact w.r.t. broyden_kernel (fuse_cuda_kernel_fast.cu:250-413).
__device__ __forceinline__ bool ia_all_corners_oob(const IaFrame& p, float gx, float gy, float gz) {
    int x0 = (int)floorf(ia_src_index(gx, p.W)), y0 = (int)floorf(ia_src_index(gy, p.H)),
        z0 = (int)floorf(ia_src_index(gz, p.D));
    return x0 < -1 || x0 > p.W - 1 || y0 < -1 || y0 > p.H - 1 || z0 < -1 || z0 > p.D - 1;
}

// The blended 3x4 transform at a point of the normalised voxel cube: J = [R | t], 8 corners in the reference order
// (tnw tne tsw tse bnw bne bsw bse: x fastest, then y, then z), zero padding (out-of-grid corners are skipped).
__device__ __forceinline__ void ia_fetch_J(const IaFrame& p, float gx, float gy, float gz, float J[12]) {
    const IaCorners cn = ia_corners(p, gx, gy, gz);
    const int W = p.W, H = p.H, D = p.D;
#if IA_VOXEL32 == 0
#pragma unroll
    for (int k = 0; k < 12; k++) J[k] = 0.0f;
#if IA_FETCH_BRANCHLESS
    // no per-corner branch: an out-of-grid corner reads voxel 0 with weight 0 (acc + 0 * v == acc exactly for finite v), so
    // that all 24 loads of a fetch can be scheduled ahead of the first use
    const bool xa = cn.x0 >= 0 && cn.x0 < W, xb = cn.x0 + 1 >= 0 && cn.x0 + 1 < W;
    const bool ya = cn.y0 >= 0 && cn.y0 < H, yb = cn.y0 + 1 >= 0 && cn.y0 + 1 < H;
    const bool za = cn.z0 >= 0 && cn.z0 < D, zb = cn.z0 + 1 >= 0 && cn.z0 + 1 < D;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const bool ok = ((c & 1) ? xb : xa) && ((c & 2) ? yb : ya) && ((c & 4) ? zb : za);
        int xi = cn.x0 + (c & 1), yi = cn.y0 + ((c >> 1) & 1), zi = cn.z0 + (c >> 2);
        float w = ((c & 1) ? cn.wx1 : cn.wx0) * ((c & 2) ? cn.wy1 : cn.wy0) * ((c & 4) ? cn.wz1 : cn.wz0);
        w = ok ? w : 0.f;
        const float4* v = p.voxel_J + (ok ? ((size_t)((zi * H + yi) * W + xi)) * 3 : 0);
        float4 a = __ldg(v), b = __ldg(v + 1), cc = __ldg(v + 2);
        J[0] = fmaf(a.x, w, J[0]); J[1] = fmaf(a.y, w, J[1]); J[2] = fmaf(a.z, w, J[2]); J[3] = fmaf(a.w, w, J[3]);
        J[4] = fmaf(b.x, w, J[4]); J[5] = fmaf(b.y, w, J[5]); J[6] = fmaf(b.z, w, J[6]); J[7] = fmaf(b.w, w, J[7]);
        J[8] = fmaf(cc.x, w, J[8]); J[9] = fmaf(cc.y, w, J[9]); J[10] = fmaf(cc.z, w, J[10]); J[11] = fmaf(cc.w, w, J[11]);
    }
#else
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int xi = cn.x0 + (c & 1), yi = cn.y0 + ((c >> 1) & 1), zi = cn.z0 + (c >> 2);
        float w = ((c & 1) ? cn.wx1 : cn.wx0) * ((c & 2) ? cn.wy1 : cn.wy0) * ((c & 4) ? cn.wz1 : cn.wz0);
        if (xi >= 0 && xi < W && yi >= 0 && yi < H && zi >= 0 && zi < D) {
            const float4* v = p.voxel_J + ((size_t)((zi * H + yi) * W + xi)) * 3;
            float4 a = __ldg(v), b = __ldg(v + 1), cc = __ldg(v + 2);
            J[0] = fmaf(a.x, w, J[0]); J[1] = fmaf(a.y, w, J[1]); J[2] = fmaf(a.z, w, J[2]); J[3] = fmaf(a.w, w, J[3]);
            J[4] = fmaf(b.x, w, J[4]); J[5] = fmaf(b.y, w, J[5]); J[6] = fmaf(b.z, w, J[6]); J[7] = fmaf(b.w, w, J[7]);
            J[8] = fmaf(cc.x, w, J[8]); J[9] = fmaf(cc.y, w, J[9]); J[10] = fmaf(cc.z, w, J[10]); J[11] = fmaf(cc.w, w, J[11]);
        }
    }
#endif
#else
    // 32-byte voxels: r[0..2] = y_c (fp32), r[3..7] = 10 halves R00 R01 | R02 R10 | R11 R12 | R20 R21 | R22 pad.
    //   sum_c w_c (y_c + R_c (x - c_c)),  c_c = c_000 + (bx hx, by hy, bz hz)
    //     = ybar + Rbar (x - c_000) - hx CX - hy CY - hz CZ,     CX = sum over corners with bx = 1 of w_c R_c[:, 0], ...
    // so the corner loop only accumulates (Rbar, ybar and the three column sums) and the translation
    // t_eff = ybar - Rbar c_000 - hx CX - hy CY - hz CZ is formed once; callers then evaluate Rbar x + t_eff as before
    // (the fp16 error of Rbar cancels between Rbar x and Rbar c_000 up to |x - c_000| <= one voxel).
    float Rb[9], yb[3], CX[3], CY[3], CZ[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Rb[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) { yb[k] = 0.f; CX[k] = 0.f; CY[k] = 0.f; CZ[k] = 0.f; }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int xi = cn.x0 + (c & 1), yi = cn.y0 + ((c >> 1) & 1), zi = cn.z0 + (c >> 2);
        float w = ((c & 1) ? cn.wx1 : cn.wx0) * ((c & 2) ? cn.wy1 : cn.wy0) * ((c & 4) ? cn.wz1 : cn.wz0);
        if (xi >= 0 && xi < W && yi >= 0 && yi < H && zi >= 0 && zi < D) {
            float r[8];
            ia_ld256(p.voxel_J + ((size_t)((zi * H + yi) * W + xi)) * 2, r);
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&r[3]));
            const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&r[4]));
            const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&r[5]));
            const float2 h3 = __half22float2(*reinterpret_cast<const __half2*>(&r[6]));
            const float R22 = __low2float(*reinterpret_cast<const __half2*>(&r[7]));
            const float wR[9] = {w * h0.x, w * h0.y, w * h1.x, w * h1.y, w * h2.x, w * h2.y, w * h3.x, w * h3.y, w * R22};
#pragma unroll
            for (int k = 0; k < 9; k++) Rb[k] += wR[k];
#pragma unroll
            for (int k = 0; k < 3; k++) yb[k] = fmaf(w, r[k], yb[k]);
            if (c & 1) { CX[0] += wR[0]; CX[1] += wR[3]; CX[2] += wR[6]; }
            if (c & 2) { CY[0] += wR[1]; CY[1] += wR[4]; CY[2] += wR[7]; }
            if (c & 4) { CZ[0] += wR[2]; CZ[1] += wR[5]; CZ[2] += wR[8]; }
        }
    }
    // canonical position of voxel (x0, y0, z0): the inverse of g = scl * (x + off) on the align_corners=True lattice
    const float c0 = fmaf((float)cn.x0, p.vox_h[0], p.vox_b[0]), c1 = fmaf((float)cn.y0, p.vox_h[1], p.vox_b[1]),
                c2 = fmaf((float)cn.z0, p.vox_h[2], p.vox_b[2]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        J[i * 4 + 0] = Rb[i * 3 + 0]; J[i * 4 + 1] = Rb[i * 3 + 1]; J[i * 4 + 2] = Rb[i * 3 + 2];
        float t = yb[i];
        t = fmaf(-Rb[i * 3 + 0], c0, t); t = fmaf(-Rb[i * 3 + 1], c1, t); t = fmaf(-Rb[i * 3 + 2], c2, t);
        t = fmaf(-p.vox_h[0], CX[i], t); t = fmaf(-p.vox_h[1], CY[i], t); t = fmaf(-p.vox_h[2], CZ[i], t);
        J[i * 4 + 3] = t;
    }
#endif
}

// One Broyden chain (one init bone) for one posed point.  fuse_cuda_kernel_fast.cu:250-413.
// Returns converged-and-inside flag; x is written only then (else zero, like the pre-zeroed output).
__device__ __forceinline__ bool ia_broyden_chain(const IaFrame& p, int bone, const float xd[3], float xo[3],
                                                 float Jinv_out[9], int* n_fetch) {
    const float* T = p.tfs[bone];
    float d0 = xd[0] - T[3], d1 = xd[1] - T[7], d2 = xd[2] - T[11];
    float x0 = d0 * T[0] + d1 * T[4] + d2 * T[8];
    float x1 = d0 * T[1] + d1 * T[5] + d2 * T[9];
    float x2 = d0 * T[2] + d1 * T[6] + d2 * T[10];
    float J[12];
    ia_fetch_J(p, p.scl[0] * (x0 + p.off[0]), p.scl[1] * (x1 + p.off[1]), p.scl[2] * (x2 + p.off[2]), J);
    int nf = 1;
    float Ji[9];
    Ji[0] = J[0]; Ji[3] = J[1]; Ji[6] = J[2];
    Ji[1] = J[4]; Ji[4] = J[5]; Ji[7] = J[6];
    Ji[2] = J[8]; Ji[5] = J[9]; Ji[8] = J[10];
    float g0 = J[0] * x0 + J[1] * x1 + J[2] * x2 + J[3] - xd[0];
    float g1 = J[4] * x0 + J[5] * x1 + J[6] * x2 + J[7] - xd[1];
    float g2 = J[8] * x0 + J[9] * x1 + J[10] * x2 + J[11] - xd[2];
    bool ok = false;
    xo[0] = xo[1] = xo[2] = 0.0f;
#pragma unroll 1
    for (int it = 0; it < 10; it++) {
        float u0 = -Ji[0] * g0 + -Ji[1] * g1 + -Ji[2] * g2;
        float u1 = -Ji[3] * g0 + -Ji[4] * g1 + -Ji[5] * g2;
        float u2 = -Ji[6] * g0 + -Ji[7] * g1 + -Ji[8] * g2;
        x0 += u0; x1 += u1; x2 += u2;
        float ix = p.scl[0] * (x0 + p.off[0]);
        float iy = p.scl[1] * (x1 + p.off[1]);
        float iz = p.scl[2] * (x2 + p.off[2]);
        ia_fetch_J(p, ix, iy, iz, J);
        nf++;
        float n0 = J[0] * x0 + J[1] * x1 + J[2] * x2 + J[3] - xd[0];
        float n1 = J[4] * x0 + J[5] * x1 + J[6] * x2 + J[7] - xd[1];
        float n2 = J[8] * x0 + J[9] * x1 + J[10] * x2 + J[11] - xd[2];
        float nrm = n0 * n0 + n1 * n1 + n2 * n2;
        if (nrm < 1e-5f * 1e-5f) {
            ok = ix >= -1 && ix <= 1 && iy >= -1 && iy <= 1 && iz >= -1 && iz <= 1;
            if (ok) {
                xo[0] = x0; xo[1] = x1; xo[2] = x2;
                if (Jinv_out) {
#pragma unroll
                    for (int k = 0; k < 9; k++) Jinv_out[k] = Ji[k];
                }
            }
            break;
        } else if (nrm > 1e-1f * 1e-1f) {
            break;
        }
        // rank-1 update of the inverse Jacobian (fuse_J_inv_update, :22-55)
        float dg0 = n0 - g0, dg1 = n1 - g1, dg2 = n2 - g2;
        float c0 = Ji[0] * u0 + Ji[3] * u1 + Ji[6] * u2;
        float c1 = Ji[1] * u0 + Ji[4] * u1 + Ji[7] * u2;
        float c2 = Ji[2] * u0 + Ji[5] * u1 + Ji[8] * u2;
        float s = c0 * dg0 + c1 * dg1 + c2 * dg2;
        float r0 = -Ji[0] * dg0 - Ji[1] * dg1 - Ji[2] * dg2;
        float r1 = -Ji[3] * dg0 - Ji[4] * dg1 - Ji[5] * dg2;
        float r2 = -Ji[6] * dg0 - Ji[7] * dg1 - Ji[8] * dg2;
        Ji[0] += c0 * (r0 + u0) / s; Ji[1] += c1 * (r0 + u0) / s; Ji[2] += c2 * (r0 + u0) / s;
        Ji[3] += c0 * (r1 + u1) / s; Ji[4] += c1 * (r1 + u1) / s; Ji[5] += c2 * (r1 + u1) / s;
        Ji[6] += c0 * (r2 + u2) / s; Ji[7] += c1 * (r2 + u2) / s; Ji[8] += c2 * (r2 + u2) / s;
        g0 = n0; g1 = n1; g2 = n2;
    }
    if (n_fetch) *n_fetch = nf;
    return ok;
}

// ------------------------------------------------------------------------------------------------
// tiny-cuda-nn HashGrid, one level, trilinear, optional d/dx.
// Level parameters of one hash-grid level; kernels that evaluate many points stage the 16 of them in
// shared memory (lane-indexed reads of the __grid_constant__ copy serialise in the constant cache).
struct IaLevel {
    float scale;
    uint32_t res, size, off;
};
__device__ __forceinline__ IaLevel ia_level(const IaFrame& p, int l) {
    IaLevel v = {p.lvl_scale[l], p.lvl_res[l], p.lvl_size[l], p.lvl_off[l]};
    return v;
}

// Hash-table entries of the fine (hashed) levels are touched once: keep them out of L1 so they do not
// evict the voxel_J lines the Broyden gathers re-use.  IA_HASH_NA: 0 = allocate, 1 = hashed levels only, 2 = all
#ifndef IA_HASH_NA
#define IA_HASH_NA 0
#endif
#ifndef IA_FETCH_BRANCHLESS
#define IA_FETCH_BRANCHLESS 0
#endif
__device__ __forceinline__ float2 ia_ldg_na(const float2* ptr) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(ptr));
    return v;
}

template <bool GRAD>
__device__ __forceinline__ void ia_hash_level(const float2* __restrict__ table, const IaLevel lv,
                                              const float xn[3], float& f0, float& f1, float dfdx[6]) {
    const float scale = lv.scale;
    const uint32_t res = lv.res, size = lv.size;
    const float2* tab = table + lv.off;
    // dense iff res^3 fits the level (levels 0-4 of the configured grid)
    const bool dense = (uint64_t)res * res * res <= (uint64_t)size;
    const bool pow2 = (size & (size - 1u)) == 0u;
    uint32_t g[3];
    float w[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float v = fmaf(scale, xn[d], 0.5f);
        float fl = floorf(v);
        g[d] = (uint32_t)(int)fl;
        w[d] = v - fl;
    }
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint32_t cx = g[0] + (c & 1), cy = g[1] + ((c >> 1) & 1), cz = g[2] + (c >> 2);
        uint32_t idx_d = cx + cy * res + cz * res * res;
        uint32_t idx_h = (cx * 1u) ^ (cy * 2654435761u) ^ (cz * 805459861u);
        uint32_t idx = dense ? idx_d : idx_h;
        if (pow2) idx &= size - 1u;            // hashed levels: size = 2^19
        else if (idx >= size) idx %= size;      // dense levels: only out-of-range corners wrap
#if IA_HASH_NA == 0
        v[c] = __ldg(tab + idx);
#elif IA_HASH_NA == 1
        v[c] = dense ? __ldg(tab + idx) : ia_ldg_na(tab + idx);
#else
        v[c] = ia_ldg_na(tab + idx);
#endif
    }
    f0 = 0.f; f1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float wt = ((c & 1) ? w[0] : 1.f - w[0]);
        wt *= ((c & 2) ? w[1] : 1.f - w[1]);
        wt *= ((c & 4) ? w[2] : 1.f - w[2]);
        f0 = fmaf(wt, v[c].x, f0);
        f1 = fmaf(wt, v[c].y, f1);
    }
    if (GRAD) {
        // d f / d xn_d = scale * sum over the 4 corner pairs along d of (w_other products) * (v_hi - v_lo)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int b1 = k & 1, b2 = k >> 1;
                float wt = (b1 ? w[d1] : 1.f - w[d1]) * (b2 ? w[d2] : 1.f - w[d2]);
                int lo = (b1 << d1) | (b2 << d2);
                int hi = lo | (1 << d);
                a0 = fmaf(wt, v[hi].x - v[lo].x, a0);
                a1 = fmaf(wt, v[hi].y - v[lo].y, a1);
            }
            dfdx[d * 2 + 0] = scale * a0;
            dfdx[d * 2 + 1] = scale * a1;
        }
    }
}

__device__ __forceinline__ float ia_team_sum(const Team& t, float v) {
#pragma unroll
    for (int o = IA_TEAM / 2; o > 0; o >>= 1) v += t.shfl_xor(v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Canonical geometry network for ONE canonical point, evaluated by the 16 lanes of a team:
// lane l owns hash level l and hidden units 4l..4l+3.  `w` points to the geometry weights
// (IA_GEO_* layout, normally staged in shared memory).  Everything returned is team-uniform.
//   GRAD = false : returns sdf only
//   GRAD = true  : also feature[13] (= raw network output incl. channel 0) and d sdf / d x (metric)
template <bool GRAD>
__device__ __forceinline__ float ia_team_geometry(const Team& team, const IaFrame& p, const float* __restrict__ w,
                                                  const float xc[3], float feat[13], float grad[3],
                                                  const IaLevel* __restrict__ lv_s = nullptr) {
    const int lane = team.thread_rank();
    float xn[3];
#pragma unroll
    for (int d = 0; d < 3; d++) xn[d] = (xc[d] - p.center[d]) / p.scale[d] + 0.5f;
    float f0, f1, dfdx[6];
    ia_hash_level<GRAD>(p.geo_hash, lv_s ? lv_s[lane] : ia_level(p, lane), xn, f0, f1, dfdx);
    // layer 1: pre[k] for hidden units 4*lane + k
    const float4* W1 = reinterpret_cast<const float4*>(w + IA_GEO_W1T) + lane;  // row stride 16 float4
    float4 acc = reinterpret_cast<const float4*>(w + IA_GEO_B1)[lane];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float xin = xn[d] * 2.0f - 1.0f;
        float4 ww = W1[d * 16];
        acc.x = fmaf(ww.x, xin, acc.x); acc.y = fmaf(ww.y, xin, acc.y);
        acc.z = fmaf(ww.z, xin, acc.z); acc.w = fmaf(ww.w, xin, acc.w);
    }
    // (the with-gradient variant keeps its loops rolled: it runs once per shading sample inside kernels whose
    //  straight-line code would otherwise overflow the instruction caches -- profiles/r1_k_primary_fused_summary.md)
    constexpr int kUnrollL1 = GRAD ? 1 : 16;
#pragma unroll kUnrollL1
    for (int l = 0; l < IA_N_LEVELS; l++) {
        float a = team.shfl(f0, l), b = team.shfl(f1, l);
        float4 wa = W1[(3 + 2 * l) * 16], wb = W1[(4 + 2 * l) * 16];
        acc.x = fmaf(wa.x, a, acc.x); acc.y = fmaf(wa.y, a, acc.y);
        acc.z = fmaf(wa.z, a, acc.z); acc.w = fmaf(wa.w, a, acc.w);
        acc.x = fmaf(wb.x, b, acc.x); acc.y = fmaf(wb.y, b, acc.y);
        acc.z = fmaf(wb.z, b, acc.z); acc.w = fmaf(wb.w, b, acc.w);
    }
    float h[4] = {ia_softplus100(acc.x), ia_softplus100(acc.y), ia_softplus100(acc.z), ia_softplus100(acc.w)};
    const float4* W2 = reinterpret_cast<const float4*>(w + IA_GEO_W2) + lane;    // row stride 16 float4
    float4 w20 = W2[0];
    float sdf = ia_team_sum(team, w20.x * h[0] + w20.y * h[1] + w20.z * h[2] + w20.w * h[3]) + w[IA_GEO_B2];
    if (GRAD) {
        feat[0] = sdf;
#pragma unroll
        for (int o = 1; o < 13; o++) {
            float4 ww = W2[o * 16];
            feat[o] = ia_team_sum(team, ww.x * h[0] + ww.y * h[1] + ww.z * h[2] + ww.w * h[3]) + w[IA_GEO_B2 + o];
        }
        // backward of sdf: delta_h = W2[0][h] * softplus'(pre_h),  softplus_beta' = sigmoid(beta * pre)
        float dl[4] = {w20.x * ia_sigmoid(100.f * acc.x), w20.y * ia_sigmoid(100.f * acc.y),
                       w20.z * ia_sigmoid(100.f * acc.z), w20.w * ia_sigmoid(100.f * acc.w)};
        float gxyz[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float4 ww = W1[d * 16];
            gxyz[d] = ia_team_sum(team, ww.x * dl[0] + ww.y * dl[1] + ww.z * dl[2] + ww.w * dl[3]);
        }
        // gradient w.r.t. this lane's two hash features
        float gf0 = 0.f, gf1 = 0.f;
#pragma unroll 1
        for (int l = 0; l < IA_N_LEVELS; l++) {
            float4 wa = W1[(3 + 2 * l) * 16], wb = W1[(4 + 2 * l) * 16];
            float sa = ia_team_sum(team, wa.x * dl[0] + wa.y * dl[1] + wa.z * dl[2] + wa.w * dl[3]);
            float sb = ia_team_sum(team, wb.x * dl[0] + wb.y * dl[1] + wb.z * dl[2] + wb.w * dl[3]);
            if (l == lane) { gf0 = sa; gf1 = sb; }
        }
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float part = gf0 * dfdx[d * 2 + 0] + gf1 * dfdx[d * 2 + 1];
            grad[d] = (gxyz[d] * 2.0f + ia_team_sum(team, part)) / p.scale[d];
        }
    }
    return sdf;
}

// ------------------------------------------------------------------------------------------------
// Blended forward LBS rotation  sum_j w_j(x_c) R_j  (deformer_torch.py:127-137,199-227):
// 24-channel trilinear fetch with BORDER padding, align_corners=True.  Team layout: lane = corner
// (lane & 7) x channel half (lane >> 3).
__device__ __forceinline__ void ia_team_fwd_rotation(const Team& team, const IaFrame& p, const float xc[3],
                                                     float R[9]) {
    const int lane = team.thread_rank();
    const int W = p.W, H = p.H, D = p.D;
    float ix = ((p.scl[0] * (xc[0] + p.off[0]) + 1.f) / 2) * (W - 1);
    float iy = ((p.scl[1] * (xc[1] + p.off[1]) + 1.f) / 2) * (H - 1);
    float iz = ((p.scl[2] * (xc[2] + p.off[2]) + 1.f) / 2) * (D - 1);
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    iz = fminf(fmaxf(iz, 0.f), (float)(D - 1));
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    int c = lane & 7, half = lane >> 3;
    int xi = (int)fx + (c & 1), yi = (int)fy + ((c >> 1) & 1), zi = (int)fz + (c >> 2);
    float wt = ((c & 1) ? ix - fx : 1.f - (ix - fx)) * ((c & 2) ? iy - fy : 1.f - (iy - fy)) *
               ((c & 4) ? iz - fz : 1.f - (iz - fz));
    bool inb = xi < W && yi < H && zi < D;  // after clamping, an out-of-range corner has weight 0
    float r[9];
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = 0.f;
    if (inb) {
        const float4* v = p.lbs_w + ((size_t)((zi * H + yi) * W + xi)) * 6 + half * 3;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            float4 a = __ldg(v + q);
            float ws[4] = {a.x * wt, a.y * wt, a.z * wt, a.w * wt};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float* T = p.tfs[half * 12 + q * 4 + e];
                r[0] = fmaf(ws[e], T[0], r[0]); r[1] = fmaf(ws[e], T[1], r[1]); r[2] = fmaf(ws[e], T[2], r[2]);
                r[3] = fmaf(ws[e], T[4], r[3]); r[4] = fmaf(ws[e], T[5], r[4]); r[5] = fmaf(ws[e], T[6], r[5]);
                r[6] = fmaf(ws[e], T[8], r[6]); r[7] = fmaf(ws[e], T[9], r[7]); r[8] = fmaf(ws[e], T[10], r[8]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = ia_team_sum(team, r[k]);
}

// ------------------------------------------------------------------------------------------------
// Result of the fused deform + geometry query (team-uniform).
struct IaQuery {
    float sdf;       // min over candidates, 1e5 if none converged
    float xc[3];     // canonical correspondence of the arg-min candidate (0 if none)
    bool valid;
    int n_valid;     // candidates after de-duplication
    // with-gradient variant:
    float grad[3];       // d sdf / d x_posed  (blended forward rotation applied)
    float grad_cano[3];
    float feat[13];
    int n_fetch;     // Broyden voxel fetches of this lane (counters)
};

template <bool GRAD>
__device__ __forceinline__ void ia_team_query(const Team& team, const IaFrame& p, const float* __restrict__ wgeo,
                                              const float xd[3], IaQuery& q) {
    const int lane = team.thread_rank();
    float x[3] = {0.f, 0.f, 0.f};
    bool ok = false;
    int nf = 0;
    if (lane < IA_N_INIT) ok = ia_broyden_chain(p, p.init_bones[lane], xd, x, nullptr, &nf);
    q.n_fetch = nf;
    // filter.cu: candidate i is dropped if a LATER valid candidate lies within 1e-4
    bool keep = ok;
#pragma unroll
    for (int j = 1; j < IA_N_INIT; j++) {
        float xj0 = team.shfl(x[0], j), xj1 = team.shfl(x[1], j), xj2 = team.shfl(x[2], j);
        bool vj = team.shfl((int)ok, j) != 0;
        float e0 = x[0] - xj0, e1 = x[1] - xj1, e2 = x[2] - xj2;
        float dist = e0 * e0 + e1 * e1 + e2 * e2;
        if (vj && j > lane && dist < 0.0001f * 0.0001f) keep = false;
    }
    unsigned mask = team.ballot(keep);
    q.n_valid = __popc(mask);
    q.valid = mask != 0;
    q.sdf = 1e5f;
    int best = 0;
    unsigned m = mask;
    while (m) {
        int c = __ffs(m) - 1;
        m &= m - 1;
        float xc[3] = {team.shfl(x[0], c), team.shfl(x[1], c), team.shfl(x[2], c)};
        float s = ia_team_geometry<false>(team, p, wgeo, xc, nullptr, nullptr);
        if (s < q.sdf) { q.sdf = s; best = c; }
    }
    // arg-min candidate; candidate 0's (zeroed) slot when nothing converged (torch.min over all-1e5)
    bool bvalid = (mask >> best) & 1u;
    float bx0 = team.shfl(x[0], best), bx1 = team.shfl(x[1], best), bx2 = team.shfl(x[2], best);
    q.xc[0] = bvalid ? bx0 : 0.f; q.xc[1] = bvalid ? bx1 : 0.f; q.xc[2] = bvalid ? bx2 : 0.f;
    if (GRAD) {
        q.grad[0] = q.grad_cano[0] = 0.f; q.grad[1] = q.grad_cano[1] = 0.f; q.grad[2] = q.grad_cano[2] = 1.f;
#pragma unroll
        for (int o = 0; o < 13; o++) q.feat[o] = 0.f;
        if (q.valid) {
            ia_team_geometry<true>(team, p, wgeo, q.xc, q.feat, q.grad_cano);
            float R[9];
            ia_team_fwd_rotation(team, p, q.xc, R);
            q.grad[0] = R[0] * q.grad_cano[0] + R[1] * q.grad_cano[1] + R[2] * q.grad_cano[2];
            q.grad[1] = R[3] * q.grad_cano[0] + R[4] * q.grad_cano[1] + R[5] * q.grad_cano[2];
            q.grad[2] = R[6] * q.grad_cano[0] + R[7] * q.grad_cano[1] + R[8] * q.grad_cano[2];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tcnn SphericalHarmonics degree 4 of unit direction (x,y,z)
__device__ __forceinline__ void ia_sh4(float x, float y, float z, float o[16]) {
    float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// acc[0..3] += W[row][4*lane..] * v  for an input-major weight matrix with 64 outputs
__device__ __forceinline__ void ia_axpy4(float4& acc, const float4* __restrict__ Wrow_lane, float v) {
    float4 ww = *Wrow_lane;
    acc.x = fmaf(ww.x, v, acc.x); acc.y = fmaf(ww.y, v, acc.y);
    acc.z = fmaf(ww.z, v, acc.z); acc.w = fmaf(ww.w, v, acc.w);
}

// hidden layer 64 -> 64 (ReLU inputs already applied), input-major weights; team layout as above
__device__ __forceinline__ float4 ia_team_dense64(const Team& team, const float* __restrict__ WT,
                                                  const float* __restrict__ B, const float h[4]) {
    const int lane = team.thread_rank();
    const float4* W = reinterpret_cast<const float4*>(WT) + lane;
    float4 acc = reinterpret_cast<const float4*>(B)[lane];
#pragma unroll 1
    for (int s = 0; s < IA_TEAM; s++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float v = team.shfl(h[k], s);
            ia_axpy4(acc, W + (s * 4 + k) * 16, v);
        }
    }
    return acc;
}

// Radiance (+ optional material) at one canonical point.  wmlp = full blob (IA_RAD_*, IA_MAT_*).
// view_w / normal_w are WORLD-space unit vectors; feat = geometry feature[13].
template <bool MATERIAL>
__device__ __forceinline__ void ia_team_radiance(const Team& team, const IaFrame& p, const float* __restrict__ wmlp,
                                                 const float xc[3], const float feat[13], const float view_w[3],
                                                 const float normal_w[3], float rgb[3], float mat[5]) {
    const int lane = team.thread_rank();
    float xn[3];
#pragma unroll
    for (int d = 0; d < 3; d++) xn[d] = (xc[d] - p.center[d]) / p.scale[d] + 0.5f;
    float f0, f1;
    ia_hash_level<false>(p.rad_hash, ia_level(p, lane), xn, f0, f1, nullptr);
    // reflect(-view, n) (models/utils.py:115), then the (d+1)/2 -> 2x-1 round trip of the encoding
    float v[3] = {-view_w[0], -view_w[1], -view_w[2]};
    float dn = v[0] * normal_w[0] + v[1] * normal_w[1] + v[2] * normal_w[2];
    float sh[16];
    {
        float r[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float refl = 2.f * dn * normal_w[d] - v[d];
            r[d] = ((refl + 1.f) / 2.f) * 2.f - 1.f;
        }
        ia_sh4(r[0], r[1], r[2], sh);
    }
    const float4* W1 = reinterpret_cast<const float4*>(wmlp + IA_RAD_W1T) + lane;
    const float4* M1 = reinterpret_cast<const float4*>(wmlp + IA_MAT_W1T) + lane;
    float4 a = reinterpret_cast<const float4*>(wmlp + IA_RAD_B1)[lane];
    float4 m = MATERIAL ? reinterpret_cast<const float4*>(wmlp + IA_MAT_B1)[lane] : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float xin = xn[d] * 2.0f - 1.0f;
        ia_axpy4(a, W1 + d * 16, xin);
        if (MATERIAL) ia_axpy4(m, M1 + d * 16, xin);
    }
#pragma unroll 1
    for (int l = 0; l < IA_N_LEVELS; l++) {
        float e0 = team.shfl(f0, l), e1 = team.shfl(f1, l);
        ia_axpy4(a, W1 + (3 + 2 * l) * 16, e0);
        ia_axpy4(a, W1 + (4 + 2 * l) * 16, e1);
        if (MATERIAL) {
            ia_axpy4(m, M1 + (3 + 2 * l) * 16, e0);
            ia_axpy4(m, M1 + (4 + 2 * l) * 16, e1);
        }
    }
    // lane o keeps feat[o] / sh[o] so the rolled loops can broadcast them with a shuffle (no register indexing)
    float my_feat = 0.f, my_sh = 0.f;
#pragma unroll
    for (int o = 0; o < 13; o++) my_feat = lane == o ? feat[o] : my_feat;
#pragma unroll
    for (int o = 0; o < 16; o++) my_sh = lane == o ? sh[o] : my_sh;
#pragma unroll 1
    for (int o = 0; o < 13; o++) {
        const float fv = team.shfl(my_feat, o);
        ia_axpy4(a, W1 + (35 + o) * 16, fv);
        if (MATERIAL) ia_axpy4(m, M1 + (35 + o) * 16, fv);
    }
#pragma unroll 1
    for (int o = 0; o < 16; o++) ia_axpy4(a, W1 + (48 + o) * 16, team.shfl(my_sh, o));
#pragma unroll
    for (int d = 0; d < 3; d++) ia_axpy4(a, W1 + (64 + d) * 16, normal_w[d]);
    float h1[4] = {fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)};
    float4 a2 = ia_team_dense64(team, wmlp + IA_RAD_W2T, wmlp + IA_RAD_B2, h1);
    float h2[4] = {fmaxf(a2.x, 0.f), fmaxf(a2.y, 0.f), fmaxf(a2.z, 0.f), fmaxf(a2.w, 0.f)};
    const float4* W3 = reinterpret_cast<const float4*>(wmlp + IA_RAD_W3) + lane;
#pragma unroll
    for (int o = 0; o < 3; o++) {
        float4 ww = W3[o * 16];
        float s = ia_team_sum(team, ww.x * h2[0] + ww.y * h2[1] + ww.z * h2[2] + ww.w * h2[3]) + wmlp[IA_RAD_B3 + o];
        rgb[o] = ia_sigmoid(s);
    }
    if (MATERIAL) {
        float g1[4] = {fmaxf(m.x, 0.f), fmaxf(m.y, 0.f), fmaxf(m.z, 0.f), fmaxf(m.w, 0.f)};
        float4 m2 = ia_team_dense64(team, wmlp + IA_MAT_W2T, wmlp + IA_MAT_B2, g1);
        float g2[4] = {fmaxf(m2.x, 0.f), fmaxf(m2.y, 0.f), fmaxf(m2.z, 0.f), fmaxf(m2.w, 0.f)};
        const float4* M3 = reinterpret_cast<const float4*>(wmlp + IA_MAT_W3) + lane;
#pragma unroll
        for (int o = 0; o < 5; o++) {
            float4 ww = M3[o * 16];
            float s = ia_team_sum(team, ww.x * g2[0] + ww.y * g2[1] + ww.z * g2[2] + ww.w * g2[3]) + wmlp[IA_MAT_B3 + o];
            mat[o] = ia_sigmoid(s) * p.mat_scale[o] + p.mat_bias[o];
        }
#pragma unroll
        for (int o = 0; o < 3; o++) mat[o] *= p.albedo_ratio[o];
    }
}

// ------------------------------------------------------------------------------------------------
// Lazy nerfacc-0.5.3-style grid traversal (single level, cone_angle 0): yields the samples of one
// ray in order, one at a time, so that callers can stop early.  Restates oracle/serial_ops.c
// `traverse_grid` (call sites models/intrinsic_avatar.py:84-93, 458-481).
struct IaMarcher {
    float t_last, t_trav, this_tmax, dt;
    float tdist[3], delta[3];
    int cur[3], step[3], over[3];
    bool continuous, done, cell_loaded, occ;

    __device__ __forceinline__ void init(const IaFrame& p, const float o[3], const float d[3], float near_plane,
                                         float far_plane, float step_size) {
        done = true;
        dt = step_size;
        float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        float tmin, tmax, tmin_t, tmax_t;
        const float* aabb = p.aabb;
        if (inv[0] >= 0) { tmin = (aabb[0] - o[0]) * inv[0]; tmax = (aabb[3] - o[0]) * inv[0]; }
        else             { tmin = (aabb[3] - o[0]) * inv[0]; tmax = (aabb[0] - o[0]) * inv[0]; }
        if (inv[1] >= 0) { tmin_t = (aabb[1] - o[1]) * inv[1]; tmax_t = (aabb[4] - o[1]) * inv[1]; }
        else             { tmin_t = (aabb[4] - o[1]) * inv[1]; tmax_t = (aabb[1] - o[1]) * inv[1]; }
        if (tmin > tmax_t || tmin_t > tmax) return;
        if (tmin_t > tmin) tmin = tmin_t;
        if (tmax_t < tmax) tmax = tmax_t;
        if (inv[2] >= 0) { tmin_t = (aabb[2] - o[2]) * inv[2]; tmax_t = (aabb[5] - o[2]) * inv[2]; }
        else             { tmin_t = (aabb[5] - o[2]) * inv[2]; tmax_t = (aabb[2] - o[2]) * inv[2]; }
        if (tmin > tmax_t || tmin_t > tmax) return;
        if (tmin_t > tmin) tmin = tmin_t;
        if (tmax_t < tmax) tmax = tmax_t;
        if (tmax <= 0) return;
        float this_tmin = fmaxf(tmin, near_plane);
        this_tmax = fminf(tmax, far_plane);
        if (!(this_tmin < this_tmax)) return;
        const float eps = 1e-6f;
        const int res = p.occ_res;
        t_last = near_plane;
        continuous = false;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float voxel = (aabb[3 + k] - aabb[k]) / (float)res;
            float rs = o[k] + d[k] * (this_tmin + eps);
            float re = o[k] + d[k] * (this_tmax - eps);
            int ci = (int)(((rs - aabb[k]) / (aabb[3 + k] - aabb[k])) * (float)res);
            int fi = (int)(((re - aabb[k]) / (aabb[3 + k] - aabb[k])) * (float)res);
            cur[k] = min(max(ci, 0), res - 1);
            fi = min(max(fi, 0), res - 1);
            int idelta = d[k] > 0 ? 1 : 0;
            float tm = ((aabb[k] + ((float)(cur[k] + idelta) * voxel) - rs) * inv[k]) + this_tmin;
            tdist[k] = (d[k] == 0.0f) ? this_tmax : tm;
            float sf = (d[k] == 0.0f) ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
            step[k] = (int)sf;
            delta[k] = (d[k] == 0.0f) ? this_tmax : voxel * inv[k] * sf;
            over[k] = fi + step[k];
        }
        done = false;
        cell_loaded = false;
    }

    // next sample [ts, te]; cont = shares its left edge with the previous sample
    __device__ __forceinline__ bool next(const uint32_t* __restrict__ occ_bits, int res, float& ts, float& te,
                                         bool& cont) {
        while (!done) {
            if (!cell_loaded) {
                t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
                int cell = (cur[0] * res + cur[1]) * res + cur[2];
                occ = (occ_bits[cell >> 5] >> (cell & 31)) & 1u;
                cell_loaded = true;
            }
            if (occ) {
                if (t_last + dt * 0.5f < t_trav) {
                    ts = t_last;
                    te = t_last + dt;
                    cont = continuous;
                    continuous = true;
                    t_last = te;
                    return true;
                }
            } else {
                while (t_last + dt * 0.5f < t_trav) t_last += dt;
                continuous = false;
            }
            // single_traversal
            bool ok = true;
            if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) {
                cur[0] += step[0]; tdist[0] += delta[0];
                if (cur[0] == over[0]) ok = false;
            } else if (tdist[1] < tdist[2]) {
                cur[1] += step[1]; tdist[1] += delta[1];
                if (cur[1] == over[1]) ok = false;
            } else {
                cur[2] += step[2]; tdist[2] += delta[2];
                if (cur[2] == over[2]) ok = false;
            }
            if (!ok || t_trav >= this_tmax) done = true;
            cell_loaded = false;
        }
        return false;
    }
};

__device__ __forceinline__ void ia_normalize(const float v[3], float o[3], float eps) {
    float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float d = fmaxf(n, eps);
    o[0] = v[0] / d; o[1] = v[1] / d; o[2] = v[2] / d;
}
// d @ R  (SMPL -> world for directions; snarf_deformer.py:153-158) incl. normalisation eps 1e-6
__device__ __forceinline__ void ia_dir_s2w(const IaFrame& p, const float d[3], float o[3]) {
    float v[3];
#pragma unroll
    for (int j = 0; j < 3; j++) v[j] = d[0] * p.w2s[0 * 4 + j] + d[1] * p.w2s[1 * 4 + j] + d[2] * p.w2s[2 * 4 + j];
    ia_normalize(v, o, 1e-6f);
}
// d @ R^T (world -> SMPL for directions; snarf_deformer.py:146-151)
__device__ __forceinline__ void ia_dir_w2s(const IaFrame& p, const float d[3], float o[3]) {
    float v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) v[i] = d[0] * p.w2s[i * 4 + 0] + d[1] * p.w2s[i * 4 + 1] + d[2] * p.w2s[i * 4 + 2];
    ia_normalize(v, o, 1e-6f);
}

// Training-mode building block (first step): backward of the canonical geometry network
//   out[13] = W2 softplus_100(W1 [2 xn - 1, hashgrid(xn)] + b1) + b2          (models/rf/geometry.py:124-172: VolumeSDF =
//   ProgressiveBandHashGrid + VanillaMLP, models/network_utils.py:58-79, 156-176, 201-244)
// with respect to the hash-grid entries, the (effective, weight-norm-folded) MLP weights and the input position, for an
// upstream gradient d_out[13] per point.  The reference gets these from autograd through tiny-cuda-nn's grid backward and
// torch.nn.Linear; here one 16-lane team per point (lane = hash level = hidden units 4 lane .. 4 lane + 3) recomputes the
// forward and scatters: weight gradients accumulate in shared memory per CTA and are flushed once, hash-grid gradients go
// to global memory with atomics (8 corners x 2 features per level), like tcnn's kernel_grid_backward.
// Not on the render path; checked against torch autograd on the oracle's network (tests/test_gpu_ops.py).
#pragma once

// corner indices and trilinear weights of one hash-grid level (the addressing of ia_hash_level)
__device__ __forceinline__ void ia_hash_corners(const IaLevel lv, const float xn[3], uint32_t idx[8], float wt[8], float w[3]) {
    const uint32_t res = lv.res, size = lv.size;
    const bool dense = (uint64_t)res * res * res <= (uint64_t)size;
    const bool pow2 = (size & (size - 1u)) == 0u;
    uint32_t g[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float v = fmaf(lv.scale, xn[d], 0.5f);
        float fl = floorf(v);
        g[d] = (uint32_t)(int)fl;
        w[d] = v - fl;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint32_t cx = g[0] + (c & 1), cy = g[1] + ((c >> 1) & 1), cz = g[2] + (c >> 2);
        uint32_t i = dense ? cx + cy * res + cz * res * res : (cx * 1u) ^ (cy * 2654435761u) ^ (cz * 805459861u);
        if (pow2) i &= size - 1u;
        else if (i >= size) i %= size;
        idx[c] = lv.off + i;
        float t = ((c & 1) ? w[0] : 1.f - w[0]);
        t *= ((c & 2) ? w[1] : 1.f - w[1]);
        t *= ((c & 4) ? w[2] : 1.f - w[2]);
        wt[c] = t;
    }
}

__global__ void __launch_bounds__(256) k_geometry_backward(const __grid_constant__ IaFrame p, const float* __restrict__ xc,
                                                           const float* __restrict__ d_out, long long n,
                                                           float* __restrict__ g_hash, float* __restrict__ g_mlp,
                                                           float* __restrict__ g_x, const uint8_t* __restrict__ valid) {
    extern __shared__ __align__(16) float smem[];
    float* w = smem;                    // geometry weights (IA_GEO_* layout)
    float* gw = smem + IA_GEO_END;      // their gradients, same layout
    ia_stage(w, p.mlp, IA_GEO_END);
    for (int i = threadIdx.x; i < IA_GEO_END; i += blockDim.x) gw[i] = 0.f;
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    const long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    const float2* tab = p.geo_hash;
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n; i += teams) {
        if (valid && !valid[i]) {      // a query without a root: no gradient (its sdf is the constant 1e5)
            if (g_x && lane < 3) g_x[i * 3 + lane] = 0.f;
            continue;
        }
        // ---- forward (as ia_team_geometry)
        float xn[3];
#pragma unroll
        for (int d = 0; d < 3; d++) xn[d] = (xc[i * 3 + d] - p.center[d]) / p.scale[d] + 0.5f;
        const IaLevel lv = ia_level(p, lane);
        uint32_t idx[8];
        float wt[8], wl[3];
        ia_hash_corners(lv, xn, idx, wt, wl);
        float2 v[8];
        float f0 = 0.f, f1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; c++) { v[c] = __ldg(tab + idx[c]); f0 = fmaf(wt[c], v[c].x, f0); f1 = fmaf(wt[c], v[c].y, f1); }
        const float4* W1 = reinterpret_cast<const float4*>(w + IA_GEO_W1T) + lane;
        float4 acc4 = reinterpret_cast<const float4*>(w + IA_GEO_B1)[lane];
        float acc[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
        float inp[35];
#pragma unroll
        for (int d = 0; d < 3; d++) inp[d] = xn[d] * 2.0f - 1.0f;
#pragma unroll
        for (int l = 0; l < IA_N_LEVELS; l++) { inp[3 + 2 * l] = team.shfl(f0, l); inp[4 + 2 * l] = team.shfl(f1, l); }
#pragma unroll
        for (int k = 0; k < 35; k++) {
            const float4 ww = W1[k * 16];
            acc[0] = fmaf(ww.x, inp[k], acc[0]); acc[1] = fmaf(ww.y, inp[k], acc[1]);
            acc[2] = fmaf(ww.z, inp[k], acc[2]); acc[3] = fmaf(ww.w, inp[k], acc[3]);
        }
        float h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = ia_softplus100(acc[k]);
        // ---- backward
        float dout[13];
#pragma unroll
        for (int o = 0; o < 13; o++) dout[o] = d_out[i * 13 + o];
        float dpre[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float dh = 0.f;
#pragma unroll
            for (int o = 0; o < 13; o++) {
                dh = fmaf(w[IA_GEO_W2 + o * 64 + 4 * lane + k], dout[o], dh);
                atomicAdd(&gw[IA_GEO_W2 + o * 64 + 4 * lane + k], dout[o] * h[k]);
            }
            // d softplus_beta(x) / dx = sigmoid(beta x); above the threshold (beta x > 20) softplus is the identity
            const float bx = 100.0f * acc[k];
            dpre[k] = dh * (bx > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-bx)));
            atomicAdd(&gw[IA_GEO_B1 + 4 * lane + k], dpre[k]);
        }
        if (lane == 0)
#pragma unroll
            for (int o = 0; o < 13; o++) atomicAdd(&gw[IA_GEO_B2 + o], dout[o]);
        float dinp_mine0 = 0.f, dinp_mine1 = 0.f, dxin[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 35; k++) {
            const float4 ww = W1[k * 16];
            atomicAdd(&gw[IA_GEO_W1T + k * 64 + 4 * lane + 0], dpre[0] * inp[k]);
            atomicAdd(&gw[IA_GEO_W1T + k * 64 + 4 * lane + 1], dpre[1] * inp[k]);
            atomicAdd(&gw[IA_GEO_W1T + k * 64 + 4 * lane + 2], dpre[2] * inp[k]);
            atomicAdd(&gw[IA_GEO_W1T + k * 64 + 4 * lane + 3], dpre[3] * inp[k]);
            const float s = ia_team_sum(team, ww.x * dpre[0] + ww.y * dpre[1] + ww.z * dpre[2] + ww.w * dpre[3]);
            if (k < 3) dxin[k] = s;
            else if (((k - 3) >> 1) == lane) { if ((k - 3) & 1) dinp_mine1 = s; else dinp_mine0 = s; }
        }
        // hash-grid entries of this lane's level
#pragma unroll
        for (int c = 0; c < 8; c++) {
            atomicAdd(&g_hash[(size_t)idx[c] * 2 + 0], wt[c] * dinp_mine0);
            atomicAdd(&g_hash[(size_t)idx[c] * 2 + 1], wt[c] * dinp_mine1);
        }
        if (g_x) {
            // d out / d xn through the trilinear weights of this level, then the xyz inputs (2 xn - 1), then xn = (x - c) / s + 1/2
            float gx[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int b1 = k & 1, b2 = k >> 1;
                    const float t = (b1 ? wl[d1] : 1.f - wl[d1]) * (b2 ? wl[d2] : 1.f - wl[d2]);
                    const int lo = (b1 << d1) | (b2 << d2), hi = lo | (1 << d);
                    a0 = fmaf(t, v[hi].x - v[lo].x, a0);
                    a1 = fmaf(t, v[hi].y - v[lo].y, a1);
                }
                gx[d] = ia_team_sum(team, lv.scale * (a0 * dinp_mine0 + a1 * dinp_mine1));
            }
            if (lane == 0)
#pragma unroll
                for (int d = 0; d < 3; d++) g_x[i * 3 + d] = (gx[d] + 2.0f * dxin[d]) / p.scale[d];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IA_GEO_END; i += blockDim.x)
        if (gw[i] != 0.f) atomicAdd(&g_mlp[i], gw[i]);
}

// ------------------------------------------------------------------------------------------------
// Training-mode building block: backward of the implicit-differentiation correction of the Broyden roots
// (ForwardDeformer.forward, version 1, models/deformers/fast_snarf/deformer_torch.py:57-76):
//   x_c = x_c* - J_inv (x_d(x_c*) - stopgrad(x_d(x_c*))),   x_d(x_c*) = sum_j w_j(x_c*) T_j [x_c*, 1]     (skinning_mask, :213-227)
// The roots x_c* and the skinning weights w_j(x_c*) (trilinear fetch of the fixed weight voxels, border padding, :199-210) are
// constants of the graph; the value of x_c is the root, and its only gradient path leads to the bone transforms:
//   u = -J_inv^T g_xc,      dL/dT_j[r][k] += w_j(x_c*) u[r] [x_c*, 1][k]        (r < 3, k < 4)
// One 16-lane team per (point, init bone) root: lane = corner (lane & 7) x channel half (lane >> 3), like ia_team_fwd_rotation;
// the 24 x 12 sums accumulate in shared memory per CTA and are flushed once.
__global__ void __launch_bounds__(256) k_deform_backward(const __grid_constant__ IaFrame p, const float* __restrict__ xc,
                                                         const uint8_t* __restrict__ valid, const float* __restrict__ J_inv,
                                                         const float* __restrict__ g_xc, long long n_roots,
                                                         float* __restrict__ g_tfs) {
    __shared__ float gt[IA_N_BONES * 12];
    for (int i = threadIdx.x; i < IA_N_BONES * 12; i += blockDim.x) gt[i] = 0.f;
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    const long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    const int W = p.W, H = p.H, D = p.D;
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n_roots; i += teams) {
        if (!valid[i]) continue;
        const float x[4] = {xc[i * 3], xc[i * 3 + 1], xc[i * 3 + 2], 1.0f};
        const float* Ji = J_inv + i * 9;
        const float g0 = g_xc[i * 3], g1 = g_xc[i * 3 + 1], g2 = g_xc[i * 3 + 2];
        float u[3];
#pragma unroll
        for (int r = 0; r < 3; r++) u[r] = -(Ji[r] * g0 + Ji[3 + r] * g1 + Ji[6 + r] * g2);
        // skinning weights at the root: this lane's corner, its 12 channels
        float ix = ((p.scl[0] * (x[0] + p.off[0]) + 1.f) / 2) * (W - 1);
        float iy = ((p.scl[1] * (x[1] + p.off[1]) + 1.f) / 2) * (H - 1);
        float iz = ((p.scl[2] * (x[2] + p.off[2]) + 1.f) / 2) * (D - 1);
        ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
        iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
        iz = fminf(fmaxf(iz, 0.f), (float)(D - 1));
        const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
        const int c = lane & 7, half = lane >> 3;
        const int xi = (int)fx + (c & 1), yi = (int)fy + ((c >> 1) & 1), zi = (int)fz + (c >> 2);
        const float wt = ((c & 1) ? ix - fx : 1.f - (ix - fx)) * ((c & 2) ? iy - fy : 1.f - (iy - fy)) *
                         ((c & 4) ? iz - fz : 1.f - (iz - fz));
        float w[12];
#pragma unroll
        for (int k = 0; k < 12; k++) w[k] = 0.f;
        if (xi < W && yi < H && zi < D) {   // after clamping, an out-of-range corner has weight 0
            const float4* v = p.lbs_w + ((size_t)((zi * H + yi) * W + xi)) * 6 + half * 3;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const float4 a = __ldg(v + q);
                w[q * 4] = a.x * wt; w[q * 4 + 1] = a.y * wt; w[q * 4 + 2] = a.z * wt; w[q * 4 + 3] = a.w * wt;
            }
        }
        // sum over the 8 corners (lanes of one half)
#pragma unroll
        for (int k = 0; k < 12; k++) {
            w[k] += team.shfl_xor(w[k], 1);
            w[k] += team.shfl_xor(w[k], 2);
            w[k] += team.shfl_xor(w[k], 4);
        }
        if (c == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) {
                if (w[k] == 0.f) continue;
                float* dst = gt + (half * 12 + k) * 12;
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int e = 0; e < 4; e++) atomicAdd(dst + r * 4 + e, w[k] * u[r] * x[e]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IA_N_BONES * 12; i += blockDim.x)
        if (gt[i] != 0.f) atomicAdd(&g_tfs[i], gt[i]);
}

// ------------------------------------------------------------------------------------------------
// Training-mode forward of the fused query (SNARFDeformer.deform, models/deformers/snarf_deformer.py:187-261, with
// ForwardDeformer.forward in training mode): the VALUES are those of the render path's query (the implicit-differentiation
// correction is zero-valued), and the kernel additionally keeps what the backward pass needs -- the arg-min candidate and the
// inverse Jacobian its Broyden chain ended with (others['J_inv'], deformer_torch.py:66).
__global__ void __launch_bounds__(256) k_query_train(const __grid_constant__ IaFrame p, const float* __restrict__ xd, long long n,
                                                     float* __restrict__ sdf, float* __restrict__ xc_out,
                                                     uint8_t* __restrict__ valid, float* __restrict__ grad,
                                                     float* __restrict__ grad_cano, float* __restrict__ feat,
                                                     float* __restrict__ J_inv, int* __restrict__ best_out) {
    extern __shared__ __align__(16) float smem[];
    ia_stage(smem, p.mlp, IA_GEO_END);
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    const long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n; i += teams) {
        const float pt[3] = {xd[i * 3 + 0], xd[i * 3 + 1], xd[i * 3 + 2]};
        float x[3] = {0.f, 0.f, 0.f}, Ji[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool ok = false;
        if (lane < IA_N_INIT) ok = ia_broyden_chain(p, p.init_bones[lane], pt, x, Ji, nullptr);
        bool keep = ok;     // filter.cu: candidate i is dropped if a LATER valid candidate lies within 1e-4
#pragma unroll
        for (int j = 1; j < IA_N_INIT; j++) {
            const float xj0 = team.shfl(x[0], j), xj1 = team.shfl(x[1], j), xj2 = team.shfl(x[2], j);
            const bool vj = team.shfl((int)ok, j) != 0;
            const float e0 = x[0] - xj0, e1 = x[1] - xj1, e2 = x[2] - xj2;
            if (vj && j > lane && e0 * e0 + e1 * e1 + e2 * e2 < 0.0001f * 0.0001f) keep = false;
        }
        const unsigned mask = team.ballot(keep);
        float s_min = 1e5f;
        int best = 0;
        unsigned m = mask;
        while (m) {
            const int c = __ffs(m) - 1;
            m &= m - 1;
            const float xc[3] = {team.shfl(x[0], c), team.shfl(x[1], c), team.shfl(x[2], c)};
            const float s = ia_team_geometry<false>(team, p, smem, xc, nullptr, nullptr);
            if (s < s_min) { s_min = s; best = c; }
        }
        const bool bvalid = (mask >> best) & 1u;
        float bx[3], bJ[9];
#pragma unroll
        for (int d = 0; d < 3; d++) { bx[d] = team.shfl(x[d], best); if (!bvalid) bx[d] = 0.f; }
#pragma unroll
        for (int k = 0; k < 9; k++) { bJ[k] = team.shfl(Ji[k], best); if (!bvalid) bJ[k] = 0.f; }
        float g[3] = {0.f, 0.f, 1.f}, gc[3] = {0.f, 0.f, 1.f}, f[13];
#pragma unroll
        for (int o = 0; o < 13; o++) f[o] = 0.f;
        if (bvalid) {
            ia_team_geometry<true>(team, p, smem, bx, f, gc);
            float R[9];
            ia_team_fwd_rotation(team, p, bx, R);
#pragma unroll
            for (int d = 0; d < 3; d++) g[d] = R[d * 3] * gc[0] + R[d * 3 + 1] * gc[1] + R[d * 3 + 2] * gc[2];
        }
        if (lane == 0) {
            sdf[i] = s_min;
            valid[i] = bvalid;
            best_out[i] = best;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                xc_out[i * 3 + d] = bx[d];
                if (grad) grad[i * 3 + d] = g[d];
                if (grad_cano) grad_cano[i * 3 + d] = gc[d];
            }
#pragma unroll
            for (int k = 0; k < 9; k++) J_inv[i * 9 + k] = bJ[k];
            if (feat)
#pragma unroll
                for (int o = 0; o < 13; o++) feat[i * 13 + o] = f[o];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Training-mode building block: backward of the two shading networks at a canonical point
//   rgb  = sigmoid(MLP_67-64-64-3 ([2 xn - 1, hashgrid_rad(xn), feature, SH4(reflect(-view, n)), n]))      models/rf/radiance.py:111-135
//   mats = sigmoid(MLP_48-64-64-5 ([2 xn - 1, hashgrid_rad(xn), feature])) * scale (* albedo ratio) + bias  models/pbr/material.py:31-51
// (VanillaMLP / LipshitzMLP with ReLU, models/network_utils.py:201-244, 360-428; weights here are the effective, folded ones)
// with respect to the radiance hash table, both networks' weights, and the inputs: position, geometry feature, normal.
// The reference gets these from autograd.  Team layout of ia_team_radiance: lane = hash level = hidden units 4 lane .. + 3.
// Weight gradients: a CTA stages the activations of its 16 teams' points in shared memory (per network: relu(a1), dL/da1,
// relu(a2), dL/da2, dL/ds), then all 512 threads add the 32 outer products into accumulators they own for the whole launch
// (thread t: column k = t & 63 of rows q = t >> 6, q + 8, ... of every weight matrix -- 34 registers), so that there is no
// atomic on a weight until the one flush per CTA at the end.  The weights are read from a shared-memory copy.  (The first
// version added every product with a shared-memory atomic from 16 teams at once and read the weights through L1: 35.9 ms per
// 2^20 points against 2.8 ms for the forward.)
static_assert(IA_SHADE_GRAD_FLOATS == IA_MLP_END - IA_RAD_W1T, "ia_b200.h: IA_SHADE_GRAD_FLOATS out of sync with the blob layout");

#define IA_ST_H1 0
#define IA_ST_DA1 64
#define IA_ST_H2 128
#define IA_ST_DA2 192
#define IA_ST_DS 256
#define IA_ST_FLOATS 264                     // one network's staged activations of one point

// One three-layer network: recomputes the forward from the team's input vector `inp` (shared memory, IN floats), applies the
// upstream gradient `dsig` [OUT] on the sigmoid outputs, stages the activations the weight gradients need in `st` (shared,
// IA_ST_FLOATS) and writes dL/d inp to `dinp` (shared, IN floats).  Returns the sigmoid outputs in `out`.
template <int IN, int OUT>
__device__ __forceinline__ void ia_team_mlp3_backward(const Team& team, const float* __restrict__ blob, float* __restrict__ st,
                                                      int o_w1t, int o_b1, int o_w2t, int o_b2, int o_w3, int o_b3,
                                                      const float* __restrict__ inp, const float* __restrict__ dsig,
                                                      float* __restrict__ dinp, float* __restrict__ out) {
    const int lane = team.thread_rank();
    const float4* W1 = reinterpret_cast<const float4*>(blob + o_w1t) + lane;
    const float4* W2 = reinterpret_cast<const float4*>(blob + o_w2t) + lane;
    const float4* W3 = reinterpret_cast<const float4*>(blob + o_w3) + lane;
    // ---- forward
    float4 a1 = reinterpret_cast<const float4*>(blob + o_b1)[lane];
#pragma unroll 1
    for (int i = 0; i < IN; i++) ia_axpy4(a1, W1 + i * 16, inp[i]);
    const float h1[4] = {fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f), fmaxf(a1.z, 0.f), fmaxf(a1.w, 0.f)};
    const float4 a2 = ia_team_dense64(team, blob + o_w2t, blob + o_b2, h1);
    const float h2[4] = {fmaxf(a2.x, 0.f), fmaxf(a2.y, 0.f), fmaxf(a2.z, 0.f), fmaxf(a2.w, 0.f)};
    float ds[OUT];
#pragma unroll
    for (int o = 0; o < OUT; o++) {
        const float4 ww = W3[o * 16];
        const float s = ia_team_sum(team, ww.x * h2[0] + ww.y * h2[1] + ww.z * h2[2] + ww.w * h2[3]) + blob[o_b3 + o];
        const float sg = ia_sigmoid(s);
        out[o] = sg;
        ds[o] = dsig[o] * sg * (1.0f - sg);
    }
    // ---- output layer
    float dh2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < OUT; o++) {
        const float4 ww = W3[o * 16];
        dh2[0] = fmaf(ww.x, ds[o], dh2[0]); dh2[1] = fmaf(ww.y, ds[o], dh2[1]);
        dh2[2] = fmaf(ww.z, ds[o], dh2[2]); dh2[3] = fmaf(ww.w, ds[o], dh2[3]);
        if (lane == 0) st[IA_ST_DS + o] = ds[o];
    }
    const float da2[4] = {a2.x > 0.f ? dh2[0] : 0.f, a2.y > 0.f ? dh2[1] : 0.f, a2.z > 0.f ? dh2[2] : 0.f, a2.w > 0.f ? dh2[3] : 0.f};
    // ---- hidden layer 64 -> 64 (input-major W2T[j][k]): dL/dh1[j] = sum_k W2T[j][k] da2[k]
    float dh1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < IA_TEAM; s++) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const float4 ww = W2[(s * 4 + kk) * 16];
            const float t = ia_team_sum(team, ww.x * da2[0] + ww.y * da2[1] + ww.z * da2[2] + ww.w * da2[3]);
            if (lane == s) dh1[kk] = t;
        }
    }
    const float da1[4] = {a1.x > 0.f ? dh1[0] : 0.f, a1.y > 0.f ? dh1[1] : 0.f, a1.z > 0.f ? dh1[2] : 0.f, a1.w > 0.f ? dh1[3] : 0.f};
    reinterpret_cast<float4*>(st + IA_ST_H1)[lane] = make_float4(h1[0], h1[1], h1[2], h1[3]);
    reinterpret_cast<float4*>(st + IA_ST_DA1)[lane] = make_float4(da1[0], da1[1], da1[2], da1[3]);
    reinterpret_cast<float4*>(st + IA_ST_H2)[lane] = make_float4(h2[0], h2[1], h2[2], h2[3]);
    reinterpret_cast<float4*>(st + IA_ST_DA2)[lane] = make_float4(da2[0], da2[1], da2[2], da2[3]);
    // ---- first layer: dL/d inp[i] = sum_k W1T[i][k] da1[k]
#pragma unroll 4
    for (int i = 0; i < IN; i++) {
        const float4 ww = W1[i * 16];
        const float t = ia_team_sum(team, ww.x * da1[0] + ww.y * da1[1] + ww.z * da1[2] + ww.w * da1[3]);
        if (lane == 0) dinp[i] = t;
    }
}

#define IA_SHB_THREADS 512
#define IA_SHB_NQ (IA_SHB_THREADS / 64)
// The weight gradients of one network a thread owns: column k = threadIdx.x & 63 of rows q + 8 r (q = threadIdx.x >> 6) of
// W1T [IN][64], W2T [64][64], W3 [OUT][64], and one bias entry (q = 0: b1[k], q = 1: b2[k], q = 2 and k < OUT: b3[k]).
template <int IN, int OUT>
struct IaMlpGrad {
    static constexpr int NQ = IA_SHB_NQ, R1 = (IN + NQ - 1) / NQ, R2 = 64 / NQ, R3 = (OUT + NQ - 1) / NQ;
    float w1[R1], w2[R2], w3[R3], b;
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int r = 0; r < R1; r++) w1[r] = 0.f;
#pragma unroll
        for (int r = 0; r < R2; r++) w2[r] = 0.f;
#pragma unroll
        for (int r = 0; r < R3; r++) w3[r] = 0.f;
        b = 0.f;
    }
    // adds the outer products of one staged point: `inp` its input vector, `st` its staged activations
    __device__ __forceinline__ void add(const float* __restrict__ inp, const float* __restrict__ st) {
        const int k = threadIdx.x & 63, q = threadIdx.x >> 6;
        const float da1 = st[IA_ST_DA1 + k], da2 = st[IA_ST_DA2 + k], h2 = st[IA_ST_H2 + k];
#pragma unroll
        for (int r = 0; r < R1; r++)
            if (q + NQ * r < IN) w1[r] = fmaf(inp[q + NQ * r], da1, w1[r]);
#pragma unroll
        for (int r = 0; r < R2; r++) w2[r] = fmaf(st[IA_ST_H1 + q + NQ * r], da2, w2[r]);
#pragma unroll
        for (int r = 0; r < R3; r++)
            if (q + NQ * r < OUT) w3[r] = fmaf(st[IA_ST_DS + q + NQ * r], h2, w3[r]);
        b += q == 0 ? da1 : q == 1 ? da2 : (q == 2 && k < OUT) ? st[IA_ST_DS + k] : 0.f;
    }
    // g: the gradient blob (indexed like the weight blob relative to IA_RAD_W1T)
    __device__ __forceinline__ void flush(float* __restrict__ g, int o_w1t, int o_b1, int o_w2t, int o_b2, int o_w3, int o_b3) const {
        const int k = threadIdx.x & 63, q = threadIdx.x >> 6;
#pragma unroll
        for (int r = 0; r < R1; r++)
            if (q + NQ * r < IN && w1[r] != 0.f) atomicAdd(g + (o_w1t - IA_RAD_W1T) + (q + NQ * r) * 64 + k, w1[r]);
#pragma unroll
        for (int r = 0; r < R2; r++)
            if (w2[r] != 0.f) atomicAdd(g + (o_w2t - IA_RAD_W1T) + (q + NQ * r) * 64 + k, w2[r]);
#pragma unroll
        for (int r = 0; r < R3; r++)
            if (q + NQ * r < OUT && w3[r] != 0.f) atomicAdd(g + (o_w3 - IA_RAD_W1T) + (q + NQ * r) * 64 + k, w3[r]);
        if (b != 0.f) {
            if (q == 0) atomicAdd(g + (o_b1 - IA_RAD_W1T) + k, b);
            else if (q == 1) atomicAdd(g + (o_b2 - IA_RAD_W1T) + k, b);
            else if (q == 2 && k < OUT) atomicAdd(g + (o_b3 - IA_RAD_W1T) + k, b);
        }
    }
};

// Jacobian-transpose product of ia_sh4: g[16] -> d/d(x, y, z)
__device__ __forceinline__ void ia_sh4_backward(float x, float y, float z, const float* __restrict__ g, float d[3]) {
    const float a = 0.48860251190291987f, b = 1.0925484305920792f, c = 0.94617469575755997f, e = 0.54627421529603959f,
                f = 0.59004358992664352f, q = 2.8906114426405538f, h = 0.45704579946446572f, k = 0.3731763325901154f,
                m = 1.4453057213202769f;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    d[0] = -a * g[3] + b * y * g[4] - b * z * g[7] + 2.f * e * x * g[8] - 6.f * f * x * y * g[9] + q * y * z * g[10] +
           h * (1.f - 5.f * z2) * g[13] + 2.f * m * x * z * g[14] + f * (-3.f * x2 + 3.f * y2) * g[15];
    d[1] = -a * g[1] + b * x * g[4] - b * z * g[5] - 2.f * e * y * g[8] + f * (-3.f * x2 + 3.f * y2) * g[9] + q * x * z * g[10] +
           h * (1.f - 5.f * z2) * g[11] - 2.f * m * y * z * g[14] + 6.f * f * x * y * g[15];
    d[2] = a * g[2] - b * y * g[5] + 2.f * c * z * g[6] - b * x * g[7] + q * x * y * g[10] - 10.f * h * y * z * g[11] +
           k * (15.f * z2 - 3.f) * g[12] - 10.f * h * x * z * g[13] + m * (x2 - y2) * g[14];
}

#define IA_SHB_TEAM_FLOATS (68 + 68 + 48 + 2 * IA_ST_FLOATS)    // per team: input vector, dL/d input of the two networks, staged activations
__global__ void __launch_bounds__(IA_SHB_THREADS) k_shade_fields_backward(const __grid_constant__ IaFrame p, const float* __restrict__ xc,
                                                               const float* __restrict__ feat, const float* __restrict__ view,
                                                               const float* __restrict__ nrm, const float* __restrict__ d_rgb,
                                                               const float* __restrict__ d_mat, long long n,
                                                               float* __restrict__ g_hash, float* __restrict__ g_mlp,
                                                               float* __restrict__ g_x, float* __restrict__ g_feat,
                                                               float* __restrict__ g_nrm) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TEAMS = IA_SHB_THREADS / IA_TEAM;
    float* wsm = smem + TEAMS * IA_SHB_TEAM_FLOATS;     // the two networks' weights, IA_SHADE_GRAD_FLOATS
    for (int i = threadIdx.x; i < IA_SHADE_GRAD_FLOATS / 4; i += IA_SHB_THREADS)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(p.mlp + IA_RAD_W1T) + i);
    const float* blob = wsm - IA_RAD_W1T;               // indexed with the blob's offsets
    __syncthreads();
    float* mine = smem + (threadIdx.x / IA_TEAM) * IA_SHB_TEAM_FLOATS;
    float* inp = mine;
    float* dr = mine + 68;
    float* dm = mine + 136;
    float* st_r = mine + 184;
    float* st_m = st_r + IA_ST_FLOATS;
    IaMlpGrad<67, 3> acc_r;
    IaMlpGrad<48, 5> acc_m;
    acc_r.zero(); acc_m.zero();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    for (long long base = (long long)blockIdx.x * TEAMS; base < n; base += (long long)gridDim.x * TEAMS) {
        const long long i = base + threadIdx.x / IA_TEAM;
        const bool active = i < n;                      // an idle team runs the same code on zeros: its products vanish
        // ---- inputs
        float xn[3];
#pragma unroll
        for (int d = 0; d < 3; d++) xn[d] = active ? (xc[i * 3 + d] - p.center[d]) / p.scale[d] + 0.5f : 0.5f;
        const IaLevel lv = ia_level(p, lane);
        uint32_t idx[8];
        float wt[8], wl[3];
        ia_hash_corners(lv, xn, idx, wt, wl);
        float2 v[8];
        float f0 = 0.f, f1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; c++) { v[c] = __ldg(p.rad_hash + idx[c]); f0 = fmaf(wt[c], v[c].x, f0); f1 = fmaf(wt[c], v[c].y, f1); }
        float vw[3] = {0.f, 0.f, 1.f}, nw[3] = {0.f, 0.f, 1.f};
        if (active) {
#pragma unroll
            for (int d = 0; d < 3; d++) { vw[d] = -view[i * 3 + d]; nw[d] = nrm[i * 3 + d]; }
        }
        const float dn = vw[0] * nw[0] + vw[1] * nw[1] + vw[2] * nw[2];
        float r[3];
#pragma unroll
        for (int d = 0; d < 3; d++) r[d] = ((2.f * dn * nw[d] - vw[d] + 1.f) / 2.f) * 2.f - 1.f;
        if (lane < 3) { inp[lane] = xn[lane] * 2.0f - 1.0f; inp[64 + lane] = nw[lane]; }
        inp[3 + 2 * lane] = f0; inp[4 + 2 * lane] = f1;
        if (lane < 13) inp[35 + lane] = active ? feat[i * 13 + lane] : 0.f;
        if (lane == 0) ia_sh4(r[0], r[1], r[2], inp + 48);
        team.sync();
        // ---- the two networks
        float dsig_r[3], dsig_m[5], out_r[3], out_m[5];
#pragma unroll
        for (int o = 0; o < 3; o++) dsig_r[o] = active ? d_rgb[i * 3 + o] : 0.f;
#pragma unroll
        for (int o = 0; o < 5; o++) dsig_m[o] = active ? d_mat[i * 5 + o] * p.mat_scale[o] * (o < 3 ? p.albedo_ratio[o] : 1.0f) : 0.f;
        ia_team_mlp3_backward<67, 3>(team, blob, st_r, IA_RAD_W1T, IA_RAD_B1, IA_RAD_W2T, IA_RAD_B2, IA_RAD_W3, IA_RAD_B3, inp,
                                     dsig_r, dr, out_r);
        ia_team_mlp3_backward<48, 5>(team, blob, st_m, IA_MAT_W1T, IA_MAT_B1, IA_MAT_W2T, IA_MAT_B2, IA_MAT_W3, IA_MAT_B3, inp,
                                     dsig_m, dm, out_m);
        __syncthreads();
        // ---- weight gradients: the CTA's 16 staged points into the accumulators each thread owns
#pragma unroll 1
        for (int t = 0; t < TEAMS; t++) {
            const float* T = smem + t * IA_SHB_TEAM_FLOATS;
            acc_r.add(T, T + 184);
            acc_m.add(T, T + 184 + IA_ST_FLOATS);
        }
        // ---- inputs: hash entries of this lane's level, position, feature, normal
        if (active) {
            const float de0 = dr[3 + 2 * lane] + dm[3 + 2 * lane], de1 = dr[4 + 2 * lane] + dm[4 + 2 * lane];
            // (samples without a root arrive with a zero upstream gradient, all at the same canonical point: their atomics
            // would queue on the same few table entries)
            if (de0 != 0.f || de1 != 0.f) {
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    atomicAdd(&g_hash[(size_t)idx[c] * 2 + 0], wt[c] * de0);
                    atomicAdd(&g_hash[(size_t)idx[c] * 2 + 1], wt[c] * de1);
                }
            }
            float gx[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int b1 = k & 1, b2 = k >> 1;
                    const float t = (b1 ? wl[d1] : 1.f - wl[d1]) * (b2 ? wl[d2] : 1.f - wl[d2]);
                    const int lo = (b1 << d1) | (b2 << d2), hi = lo | (1 << d);
                    a0 = fmaf(t, v[hi].x - v[lo].x, a0);
                    a1 = fmaf(t, v[hi].y - v[lo].y, a1);
                }
                gx[d] = ia_team_sum(team, lv.scale * (a0 * de0 + a1 * de1));
            }
            if (lane < 3) {
                if (g_x) g_x[i * 3 + lane] = ((lane == 0 ? gx[0] : lane == 1 ? gx[1] : gx[2]) + 2.0f * (dr[lane] + dm[lane])) / p.scale[lane];
            }
            if (g_feat && lane < 13) g_feat[i * 13 + lane] = dr[35 + lane] + dm[35 + lane];
            if (g_nrm && lane == 0) {
                float gr[3];
                ia_sh4_backward(r[0], r[1], r[2], dr + 48, gr);
                // refl = 2 (v . n) n - v:  d/dn_j = 2 v_j (g . n) + 2 (v . n) g_j
                const float gn = gr[0] * nw[0] + gr[1] * nw[1] + gr[2] * nw[2];
#pragma unroll
                for (int d = 0; d < 3; d++) g_nrm[i * 3 + d] = dr[64 + d] + 2.f * vw[d] * gn + 2.f * dn * gr[d];
            }
        }
        __syncthreads();                                // the staging is rewritten by the next round
    }
    acc_r.flush(g_mlp, IA_RAD_W1T, IA_RAD_B1, IA_RAD_W2T, IA_RAD_B2, IA_RAD_W3, IA_RAD_B3);
    acc_m.flush(g_mlp, IA_MAT_W1T, IA_MAT_B1, IA_MAT_W2T, IA_MAT_B2, IA_MAT_W3, IA_MAT_B3);
}

// ------------------------------------------------------------------------------------------------
// Training-mode building block: compositing along the primary rays and its backward.
//   alpha_i = 1 - exp(-sigma(sdf_i) dist_i),  sigma = Laplace-CDF density        models/rf/density.py:17-34; intrinsic_avatar.py:390-394
//   w_i = alpha_i prod_{j<i} (1 - alpha_j)                                       nerfacc 0.5.3 render_weight_from_alpha
//   comp[c] = sum_i w_i value_i[c],  opacity = sum_i w_i                         nerfacc accumulate_along_rays
// as `rendering_with_normals_mats_sdf` strings them together (models/volrend.py:336-364); the reference differentiates the
// chain with autograd (nerfacc's custom backward for the weights).  One thread per ray, samples in packed order.
//   dL/dalpha_i = dL/dw_i T_i - (sum_{j>i} dL/dw_j w_j) / (1 - alpha_i)
__device__ __forceinline__ float ia_sigma(float sdf, float beta) {
    const float sgn = (sdf > 0.0f) ? 1.0f : ((sdf < 0.0f) ? -1.0f : 0.0f);
    return (1.0f / beta) * (0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) / beta));
}

__global__ void __launch_bounds__(128) k_volrend(const int* __restrict__ packed_info, const float* __restrict__ sdf,
                                                 const float* __restrict__ dists, const float* __restrict__ values, int C,
                                                 float beta, long long n_rays, float* __restrict__ weights,
                                                 float* __restrict__ comp, float* __restrict__ opacity) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const int s = packed_info[r * 2], cnt = packed_info[r * 2 + 1];
    float acc[IA_VOLREND_MAX_C];
#pragma unroll
    for (int c = 0; c < IA_VOLREND_MAX_C; c++) acc[c] = 0.f;
    float T = 1.0f, op = 0.f;
    for (int i = s; i < s + cnt; i++) {
        const float a = ia_alpha(sdf[i], dists[i], beta);
        const float w = T * a;
        T *= (1.0f - a);
        if (weights) weights[i] = w;
        op += w;
#pragma unroll
        for (int c = 0; c < IA_VOLREND_MAX_C; c++)
            if (c < C) acc[c] = fmaf(w, values[(size_t)i * C + c], acc[c]);
    }
    opacity[r] = op;
#pragma unroll
    for (int c = 0; c < IA_VOLREND_MAX_C; c++)
        if (c < C) comp[r * C + c] = acc[c];
}

__global__ void __launch_bounds__(128) k_volrend_backward(const int* __restrict__ packed_info, const float* __restrict__ sdf,
                                                          const float* __restrict__ dists, const float* __restrict__ values,
                                                          int C, float beta, const float* __restrict__ d_comp,
                                                          const float* __restrict__ d_opacity,
                                                          const float* __restrict__ d_weights, long long n_rays,
                                                          float* __restrict__ g_sdf, float* __restrict__ g_values,
                                                          float* __restrict__ g_beta) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float gb = 0.f;
    if (r < n_rays) {
        const int s = packed_info[r * 2], cnt = packed_info[r * 2 + 1];
        float dc[IA_VOLREND_MAX_C];
#pragma unroll
        for (int c = 0; c < IA_VOLREND_MAX_C; c++) dc[c] = c < C ? d_comp[r * C + c] : 0.f;
        const float dop = d_opacity ? d_opacity[r] : 0.f;
        // forward: the transmittance in front of every sample, parked in g_sdf
        float T = 1.0f;
        for (int i = s; i < s + cnt; i++) {
            g_sdf[i] = T;
            T *= (1.0f - ia_alpha(sdf[i], dists[i], beta));
        }
        // reverse: suffix sum of dL/dw_j w_j
        float suffix = 0.f;
        for (int i = s + cnt - 1; i >= s; i--) {
            const float sd = sdf[i], dist = dists[i];
            const float sigma = ia_sigma(sd, beta);
            const float a = 1.0f - expf(-sigma * dist);
            const float Ti = g_sdf[i];
            const float w = Ti * a;
            float dw = dop + (d_weights ? d_weights[i] : 0.f);
#pragma unroll
            for (int c = 0; c < IA_VOLREND_MAX_C; c++)
                if (c < C) {
                    dw = fmaf(dc[c], values[(size_t)i * C + c], dw);
                    if (g_values) g_values[(size_t)i * C + c] = w * dc[c];
                }
            const float da = dw * Ti - suffix / fmaxf(1.0f - a, 1e-10f);
            suffix = fmaf(dw, w, suffix);
            // alpha = 1 - exp(-sigma dist);  sigma = (1/beta)(1/2 + 1/2 sgn(s) expm1(-|s| / beta))
            const float dsig = da * dist * (1.0f - a);
            const float e = expf(-fabsf(sd) / beta);
            g_sdf[i] = sd == 0.0f ? 0.0f : dsig * (-0.5f * e / (beta * beta));
            gb += dsig * (-sigma / beta + 0.5f * sd * e / (beta * beta * beta));
        }
    }
    for (int o = 16; o > 0; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
    if ((threadIdx.x & 31) == 0 && gb != 0.f) atomicAdd(g_beta, gb);
}

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
                                                           float* __restrict__ g_x) {
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

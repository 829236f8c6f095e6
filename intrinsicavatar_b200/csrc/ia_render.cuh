// Render pipeline kernels (included at the end of ia_kernels.cu).
//
//   k_primary_setup : world->SMPL ray transform, grid march count, hit list, default outputs
//   k_primary       : team-per-hit-ray: march -> 2x (SDF query + CDF merge resample) -> shading
//                     (deform + geometry w/ gradient + radiance + material) -> accumulate; emits
//                     the per-ray shading samples for the PBR stage
//   k_resample      : warp-per-hit-ray `ray_resampling` (spp shading samples, zero-crossing snap)
//   k_shade_wf      : wavefront secondary-ray integrator (ia_wavefront.cuh)
//   k_composite     : background composite + sRGB
#pragma once

#define IA_PRIMARY_THREADS 256
#define IA_SHADE_THREADS 256
#define IA_TILE 1024
#define IA_TILE_PIX 520  // >= IA_TILE / 2 + 2 pixel slots per tile (spp >= 2)

// work counter slots in ctx->d_work
#define IA_W_NHIT 0
#define IA_W_NSAMPLES 1
#define IA_W_PRIMARY_NEXT 2
#define IA_W_TILE_NEXT 3

#include "ia_wavefront.cuh"

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ia_ray_w2s(const IaFrame& p, const float* __restrict__ ray, float o[3], float d[3],
                                           float& far) {
    // transform_rays_w2s (models/deformers/snarf_deformer.py:128-144)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        o[i] = ray[0] * p.w2s[i * 4 + 0] + ray[1] * p.w2s[i * 4 + 1] + ray[2] * p.w2s[i * 4 + 2] + p.w2s[i * 4 + 3];
        d[i] = ray[3] * p.w2s[i * 4 + 0] + ray[4] * p.w2s[i * 4 + 1] + ray[5] * p.w2s[i * 4 + 2];
    }
    far = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]) + 1.0f;
}

// bg_rgb [n_rays][3]: what a background-assigned shading sample / an empty ray contributes to the physically
// based buffers: the background colour, or with add_emitter the envmap along the primary ray
// (emitter.eval(transform_dirs_s2w(rays_d)), models/intrinsic_avatar.py:1319-1341, 1454-1490).
// Hit list in RAY ORDER (deterministic): k_primary_setup flags the rays that enter an occupied cell and counts them per
// block, k_hit_scan turns the block counts into offsets, k_hit_compact writes (ray index, o, d, far) at
// offset[block] + rank within the block.  Neighbouring hit slots are then neighbouring rays of the caller's array
// (neighbouring pixels of an image row): the shading stage feeds them as bundles of parallel rays (ia_wavefront.cuh).
// (Round 1 appended with one atomic per warp: slots were chunks of <= 32 pixels from random places of the image.)
__global__ void __launch_bounds__(256) k_primary_setup(const __grid_constant__ IaFrame p, const float* __restrict__ rays, long long n_rays,
                                uint8_t* __restrict__ hit_flag, int* __restrict__ blk_cnt,
                                ia_outputs out, float* __restrict__ acc6, float* __restrict__ bg_rgb, const IaEnv E,
                                int add_emitter) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    float o[3], d[3], far = 0.f;
    if (r < n_rays) {
        ia_ray_w2s(p, rays + r * 8, o, d, far);
        IaMarcher m;
        m.init(p, o, d, 0.0f, 1e10f, p.step_primary);
        float ts, te;
        bool cont;
        hit = m.next(p.occ_bits, p.occ_res, ts, te, cont);
        hit_flag[r] = hit ? 1 : 0;
        // defaults of a ray without samples
        if (out.comp_rgb) { out.comp_rgb[r * 3] = 0.f; out.comp_rgb[r * 3 + 1] = 0.f; out.comp_rgb[r * 3 + 2] = 0.f; }
        if (out.comp_normal) { out.comp_normal[r * 3] = 0.f; out.comp_normal[r * 3 + 1] = 0.f; out.comp_normal[r * 3 + 2] = 0.f; }
        if (out.comp_albedo) { out.comp_albedo[r * 3] = 0.f; out.comp_albedo[r * 3 + 1] = 0.f; out.comp_albedo[r * 3 + 2] = 0.f; }
        if (out.opacity) out.opacity[r] = 0.f;
        if (out.depth) out.depth[r] = far;
        if (out.comp_roughness) out.comp_roughness[r] = 0.f;
        if (out.comp_metallic) out.comp_metallic[r] = 0.f;
        if (out.num_samples) out.num_samples[r] = 0;
        float bg[3] = {p.background[0], p.background[1], p.background[2]};
        if (add_emitter) {
            float dw[3];
            ia_dir_s2w(p, d, dw);
            ia_env_eval(E, dw, bg);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) bg_rgb[r * 3 + k] = bg[k];
#pragma unroll
        for (int k = 0; k < 6; k++) acc6[r * 6 + k] = hit ? 0.f : bg[k % 3];
    }
    const int n = __syncthreads_count(hit);
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = n;
}

// exclusive scan of the per-block hit counts (one CTA; in place) and the total -> work[IA_W_NHIT]
__global__ void __launch_bounds__(1024) k_hit_scan(int* __restrict__ blk_cnt, int n_blocks, int* __restrict__ work) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = b < n_blocks ? blk_cnt[b] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sum[lane] = winc - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int carry = carry_s;
        if (b < n_blocks) blk_cnt[b] = carry + warp_sum[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[warp] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) work[IA_W_NHIT] = carry_s;
}

__global__ void __launch_bounds__(256) k_hit_compact(const __grid_constant__ IaFrame p, const float* __restrict__ rays, long long n_rays,
                                                     const uint8_t* __restrict__ hit_flag, const int* __restrict__ blk_off,
                                                     int* __restrict__ hit_rays, float* __restrict__ hit_od) {
    __shared__ int warp_cnt[8];
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hit = r < n_rays && hit_flag[r];
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[warp] = __popc(b);
    __syncthreads();
    if (!hit) return;
    int slot = blk_off[blockIdx.x] + __popc(b & ((1u << lane) - 1));
    for (int w = 0; w < warp; w++) slot += warp_cnt[w];
    float o[3], d[3], far;
    ia_ray_w2s(p, rays + r * 8, o, d, far);
    hit_rays[slot] = (int)r;
    float* od = hit_od + (size_t)slot * 8;
    od[0] = o[0]; od[1] = o[1]; od[2] = o[2]; od[3] = d[0]; od[4] = d[1]; od[5] = d[2]; od[6] = far; od[7] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// Primary stage = three kernels (the first generation fused them in one 22k-instruction kernel that spent
// 87 % of its samples waiting for instruction fetch -- profiles/r1_k_primary_fused_summary.md):
//   k_prim_edges  : team per hit ray: march -> edges -> 2 x (SDF queries + CDF merge resample) -> final intervals
//   k_prim_shade  : team per SHADING SAMPLE (flat list): deform + geometry w/ gradient + radiance + material
//   k_prim_accum  : thread per hit ray: alpha -> weights -> 7 accumulations, weights into the sample records
struct IaPrimarySmem {
    float vals[2][IA_CAP];
    float aux[IA_CAP];
    uint8_t flags[2][IA_CAP];
};

// per-sample record written by k_prim_shade next to IaSample (which carries ts, te, sdf, n, albedo, rough, metal)

__global__ void __launch_bounds__(IA_PRIMARY_THREADS, 2)
k_prim_edges(const __grid_constant__ IaFrame p, const float* __restrict__ hit_od, int* __restrict__ hit_info,
             IaSample* __restrict__ samples, IaSampleAux* __restrict__ aux, long long sample_cap, int* __restrict__ work,
             unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) float smem[];
    float* wgeo = smem;
    IaPrimarySmem* tsm = reinterpret_cast<IaPrimarySmem*>(smem + IA_GEO_END);
    ia_stage(wgeo, p.mlp, IA_GEO_END);
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    IaPrimarySmem& S = tsm[threadIdx.x / IA_TEAM];
    const int n_hit = work[IA_W_NHIT];
    unsigned c_q = 0, c_fetch = 0, c_geo = 0, c_over = 0, c_samples = 0;
    const float step = p.step_primary;

    while (true) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&work[IA_W_PRIMARY_NEXT], 1);
        slot = team.shfl(slot, 0);
        if (slot >= n_hit) break;
        const float* od = hit_od + (size_t)slot * 8;
        const float o[3] = {od[0], od[1], od[2]}, d[3] = {od[3], od[4], od[5]};

        // ---- 1. grid march -> edge list (uniform across the team; lane 0 writes)
        int ne = 0;
        {
            IaMarcher m;
            m.init(p, o, d, 0.0f, 1e10f, step);
            float ts, te;
            bool cont;
            bool over = false;
            while (m.next(p.occ_bits, p.occ_res, ts, te, cont)) {
                if (ne + 2 > IA_CAP - 34) { over = true; break; }
                if (!cont) {
                    if (lane == 0) { S.vals[0][ne] = ts; S.flags[0][ne] = 1; S.vals[0][ne + 1] = te; S.flags[0][ne + 1] = 2; }
                    ne += 2;
                } else {
                    if (lane == 0) { S.vals[0][ne] = te; S.flags[0][ne - 1] |= 1; S.flags[0][ne] = 2; }
                    ne += 1;
                }
            }
            if (over) c_over++;
        }
        team.sync();
        int cur = 0;
        // ---- 2. two rounds of importance resampling (models/intrinsic_avatar.py:1185-1238); ONE query site:
        //         round 0 queries every edge position (coarse_alpha_fn), round 1 the interval midpoints (alpha_fn)
#pragma unroll 1
        for (int round = 0; round < 2; round++) {
            const float* vals = S.vals[cur];
            const uint8_t* fl = S.flags[cur];
            const int n_pts = round == 0 ? ne : ne - 1;
#pragma unroll 1
            for (int e = 0; e < n_pts; e++) {
                const bool need = round == 0 || (fl[e] & 1);
                float sd = 1e10f;
                if (need) {
                    const float t = round == 0 ? vals[e] : (vals[e] + vals[e + 1]) / 2.0f;
                    float x[3] = {o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t};
                    IaQuery q;
                    ia_team_query<false>(team, p, wgeo, x, q);
                    c_q++; c_fetch += q.n_fetch; c_geo += q.n_valid;
                    sd = q.sdf;
                }
                if (lane == 0) S.aux[e] = sd;
            }
            team.sync();
            if (lane == 0) {
                // round 0: alpha from min(sdf_L, sdf_R) with the fixed step; round 1: midpoint sdf, dist = t_R - t_L
                float T = 1.f;
                for (int e = 0; e < ne; e++) {
                    float a = 0.f;
                    if ((fl[e] & 1) && e + 1 < ne)
                        a = round == 0 ? ia_alpha(fminf(S.aux[e], S.aux[e + 1]), step, p.beta)
                                       : ia_alpha(S.aux[e], vals[e + 1] - vals[e], p.beta);
                    S.aux[e] = T * a;
                    T *= (1.f - a);
                }
            }
            team.sync();
            int n_out = 0;
            if (lane == 0) n_out = ia_merge_resample(vals, fl, S.aux, ne, 16, S.vals[cur ^ 1], S.flags[cur ^ 1], nullptr);
            n_out = team.shfl(n_out, 0);
            team.sync();
            ne = n_out;
            cur ^= 1;
        }
        // ---- 3. hand the final intervals to the shading kernel
        const float* vals = S.vals[cur];
        const uint8_t* fl = S.flags[cur];
        int n_iv = 0;
        for (int e = 0; e + 1 < ne; e++) n_iv += (fl[e] & 1);
        int base = 0;
        if (lane == 0) base = atomicAdd(&work[IA_W_NSAMPLES], n_iv);
        base = team.shfl(base, 0);
        const bool pool_ok = (long long)base + n_iv <= sample_cap;
        if (!pool_ok) {
            // The sample pool is full: this ray renders as background and IA_CNT_OVERFLOW tells the host (engine.render
            // grows the pool and renders the frame again).  Its part of the pool below the capacity is marked dead so
            // that k_prim_shade never reads an unwritten record.
            c_over++;
            for (long long i = (long long)base + lane; i < sample_cap && i < (long long)base + n_iv; i += IA_TEAM) aux[i].slot = -1;
        }
        if (pool_ok && lane == 0) {
            int k = 0;
            for (int e = 0; e + 1 < ne; e++) {
                if (!(fl[e] & 1)) continue;
                samples[(size_t)base + k].ts = vals[e];
                samples[(size_t)base + k].te = vals[e + 1];
                aux[(size_t)base + k].slot = slot;
                k++;
            }
        }
        if (lane == 0) {
            hit_info[slot * 2 + 0] = base;
            hit_info[slot * 2 + 1] = pool_ok ? n_iv : 0;
        }
        c_samples += pool_ok ? n_iv : 0;
        team.sync();
    }
    if (lane == 0) {
        if (c_q) atomicAdd(&counters[IA_CNT_QUERIES], c_q);
        if (c_geo) atomicAdd(&counters[IA_CNT_GEO_EVAL], c_geo);
        if (c_over) atomicAdd(&counters[IA_CNT_OVERFLOW], c_over);
        if (c_samples) atomicAdd(&counters[IA_CNT_SAMPLES], c_samples);
    }
    if (c_fetch) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], c_fetch);
}

// rendering_with_normals_mats_sdf's per-sample part (rgb_normal_mats_alpha_fn, models/intrinsic_avatar.py:1066-1156).
// A warp owns 16 consecutive shading samples: its two 16-lane teams run the fused query (Broyden + geometry with gradient +
// blended rotation) sample by sample and leave the inputs of the two colour networks as a row of the warp's shading tile
// (ia_mma.cuh); the radiance (67 -> 64 -> 64 -> 3) and material (48 -> 64 -> 64 -> 5) networks then run once per 16 samples
// on the tensor cores (3xTF32 mma.sync, pre-split B fragments from global memory through L1).  IA_PRIM_SHADE_MMA=0: the
// round-1 form (both networks per sample on the 16 lanes of the team, all weights in shared memory).
#ifndef IA_PRIM_SHADE_CTAS
#define IA_PRIM_SHADE_CTAS 2
#endif
#ifndef IA_PRIM_SHADE_MMA
#define IA_PRIM_SHADE_MMA 1
#endif
#if IA_PRIM_SHADE_MMA
#define IA_PRIM_SHADE_SMEM ((IA_GEO_END + (IA_PRIMARY_THREADS / 32) * IA_SHADE_TILE) * sizeof(float))
__global__ void __launch_bounds__(IA_PRIMARY_THREADS, IA_PRIM_SHADE_CTAS)
k_prim_shade(const __grid_constant__ IaFrame p, const float* __restrict__ hit_od, IaSample* __restrict__ samples,
             IaSampleAux* __restrict__ aux, long long sample_cap, const int* __restrict__ work,
             unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) float smem[];
    float* wgeo = smem;
    float* xs_all = smem + IA_GEO_END;
    ia_stage(wgeo, p.mlp, IA_GEO_END);
    for (int i = threadIdx.x; i < (IA_PRIMARY_THREADS / 32) * IA_SHADE_TILE; i += blockDim.x) xs_all[i] = 0.f;
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, l = lane & 15, half = lane >> 4, warp = threadIdx.x >> 5;
    float* xs = xs_all + warp * IA_SHADE_TILE;
    const float4* frags = reinterpret_cast<const float4*>(p.mlp + IA_MLP_END);
    const long long n = min((long long)work[IA_W_NSAMPLES], sample_cap);
    const long long n_batches = (n + 15) / 16;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    unsigned c_qg = 0, c_fetch = 0, c_geo = 0, c_rad = 0;
    for (long long bb = (long long)blockIdx.x * (blockDim.x >> 5); bb < n_batches; bb += warps) {
        const long long b = bb + warp;
        unsigned vmask = 0;
#pragma unroll 1
        for (int r0 = 0; r0 < 16; r0 += 2) {
            const int r = r0 + half;
            const long long i = b * 16 + r;
            const int slot = (b < n_batches && i < n) ? aux[i].slot : -1;
            if (slot < 0) continue;  // past the end, or the dead record of a ray that did not fit the pool (k_prim_edges)
            const float* od = hit_od + (size_t)slot * 8;
            const float o[3] = {od[0], od[1], od[2]}, d[3] = {od[3], od[4], od[5]};
            const float ts = samples[i].ts, te = samples[i].te;
            const float tm = (ts + te) / 2.0f;
            float x[3] = {o[0] + d[0] * tm, o[1] + d[1] * tm, o[2] + d[2] * tm};
            IaQuery q;
            ia_team_query<true>(team, p, wgeo, x, q);
            c_qg++; c_fetch += q.n_fetch; c_geo += q.n_valid + (q.valid ? 1 : 0);
            float view_w[3], nsm[3], nw[3];
            ia_dir_s2w(p, d, view_w);
            ia_normalize(q.grad, nsm, 1e-6f);
            ia_dir_s2w(p, q.grad, nw);
            if (q.valid) {
                // the row of the shading tile: radiance hash features, position, geometry feature, SH(reflected dir), normal
                float* row = xs + r * IA_SHADE_LD;
                float xn[3];
#pragma unroll
                for (int k = 0; k < 3; k++) xn[k] = (q.xc[k] - p.center[k]) / p.scale[k] + 0.5f;
                float f0, f1;
                ia_hash_level<false>(p.rad_hash, ia_level(p, l), xn, f0, f1, nullptr);
                const float v[3] = {-view_w[0], -view_w[1], -view_w[2]};
                const float dn = v[0] * nw[0] + v[1] * nw[1] + v[2] * nw[2];
                float rr[3], sh[16];
#pragma unroll
                for (int k = 0; k < 3; k++) rr[k] = ((2.f * dn * nw[k] - v[k] + 1.f) / 2.f) * 2.f - 1.f;
                ia_sh4(rr[0], rr[1], rr[2], sh);
                float my_sh = 0.f, my_f = 0.f, my_3 = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k++) my_sh = l == k ? sh[k] : my_sh;
#pragma unroll
                for (int k = 0; k < 13; k++) my_f = l == k ? q.feat[k] : my_f;
#pragma unroll
                for (int k = 0; k < 3; k++) my_3 = l == k ? xn[k] * 2.0f - 1.0f : (l == 3 + k ? nw[k] : my_3);
                *reinterpret_cast<float2*>(row + 2 * l) = make_float2(f0, f1);
                row[48 + l] = my_sh;
                if (l < 13) row[35 + l] = my_f;
                if (l < 3) row[32 + l] = my_3;
                else if (l < 6) row[64 + l - 3] = my_3;
                vmask |= 1u << r;
                c_rad++;
            }
            if (l == 0) {
                IaSample s;
                s.ts = ts; s.te = te; s.w = 0.f; s.sdf = q.sdf;
                s.n[0] = nsm[0]; s.n[1] = nsm[1]; s.n[2] = nsm[2];
                s.albedo[0] = 0.f; s.albedo[1] = 0.f; s.albedo[2] = 0.f;
                s.rough = 0.f; s.metal = 0.f;
                samples[i] = s;
                IaSampleAux a;
                a.rgb[0] = 0.f; a.rgb[1] = 0.f; a.rgb[2] = 0.f;
                a.nw[0] = nw[0]; a.nw[1] = nw[1]; a.nw[2] = nw[2];
                a.slot = slot; a.pad = 0;
                aux[i] = a;
            }
        }
        vmask |= __shfl_xor_sync(FULL, vmask, 16);
        // The warps of the CTA walk the weight fragments together (264 fragments = 135 KB per 16 samples, more than L1 keeps
        // next to the voxel gathers): one L2 read per CTA instead of one per warp.  Measured, primary stage at 512^2:
        // 44.0 ms (round-1 form) / 60.5 ms (this form without the barrier) / 41.1 ms (with it).
        __syncthreads();
        if (vmask) {
            float rgb[3], mat[5];
            ia_warp_mlp3<9, 3>(xs, IA_SHADE_LD, frags + IA_FRAG_RAD1 * 32, p.mlp + IA_RAD_B1, p.mlp + IA_RAD_B2, p.mlp + IA_RAD_B3, rgb);
            ia_warp_mlp3<6, 5>(xs, IA_SHADE_LD, frags + IA_FRAG_MAT1 * 32, p.mlp + IA_MAT_B1, p.mlp + IA_MAT_B2, p.mlp + IA_MAT_B3, mat);
            if (lane < 16 && ((vmask >> lane) & 1u)) {
                const long long i = b * 16 + lane;
#pragma unroll
                for (int k = 0; k < 3; k++) aux[i].rgb[k] = ia_sigmoid(rgb[k]);
#pragma unroll
                for (int k = 0; k < 5; k++) mat[k] = ia_sigmoid(mat[k]) * p.mat_scale[k] + p.mat_bias[k];
#pragma unroll
                for (int k = 0; k < 3; k++) samples[i].albedo[k] = mat[k] * p.albedo_ratio[k];
                samples[i].rough = mat[3];
                samples[i].metal = mat[4];
            }
        }
        __syncwarp();
    }
    if (l == 0) {
        if (c_qg) atomicAdd(&counters[IA_CNT_QUERIES_GRAD], c_qg);
        if (c_rad) atomicAdd(&counters[IA_CNT_RAD_EVAL], c_rad);
        if (c_geo) atomicAdd(&counters[IA_CNT_GEO_EVAL], c_geo);
        if (c_qg) atomicAdd(&counters[IA_CNT_SKIN_FETCH], c_qg);
    }
    if (c_fetch) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], c_fetch);
}
#else
#define IA_PRIM_SHADE_SMEM (IA_MLP_END * sizeof(float))
__global__ void __launch_bounds__(IA_PRIMARY_THREADS, IA_PRIM_SHADE_CTAS)
k_prim_shade(const __grid_constant__ IaFrame p, const float* __restrict__ hit_od, IaSample* __restrict__ samples,
             IaSampleAux* __restrict__ aux, long long sample_cap, const int* __restrict__ work,
             unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) float smem[];
    float* wmlp = smem;
    ia_stage(wmlp, p.mlp, IA_MLP_END);
    __syncthreads();
    Team team = cg::tiled_partition<IA_TEAM>(cg::this_thread_block());
    const int lane = team.thread_rank();
    const long long n = min((long long)work[IA_W_NSAMPLES], sample_cap);
    const long long teams = (long long)gridDim.x * (blockDim.x / IA_TEAM);
    unsigned c_qg = 0, c_fetch = 0, c_geo = 0, c_rad = 0;
    for (long long i = (long long)blockIdx.x * (blockDim.x / IA_TEAM) + threadIdx.x / IA_TEAM; i < n; i += teams) {
        const int slot = aux[i].slot;
        if (slot < 0) continue;  // dead record of a ray that did not fit the pool (k_prim_edges)
        const float* od = hit_od + (size_t)slot * 8;
        const float o[3] = {od[0], od[1], od[2]}, d[3] = {od[3], od[4], od[5]};
        const float ts = samples[i].ts, te = samples[i].te;
        const float tm = (ts + te) / 2.0f;
        float x[3] = {o[0] + d[0] * tm, o[1] + d[1] * tm, o[2] + d[2] * tm};
        IaQuery q;
        ia_team_query<true>(team, p, wmlp, x, q);
        c_qg++; c_fetch += q.n_fetch; c_geo += q.n_valid + (q.valid ? 1 : 0);
        float view_w[3], nsm[3], nw[3], rgb[3] = {0, 0, 0}, mat[5] = {0, 0, 0, 0, 0};
        ia_dir_s2w(p, d, view_w);
        ia_normalize(q.grad, nsm, 1e-6f);
        ia_dir_s2w(p, q.grad, nw);
        if (q.valid) {
            ia_team_radiance<true>(team, p, wmlp, q.xc, q.feat, view_w, nw, rgb, mat);
            c_rad++;
        }
        if (lane == 0) {
            IaSample s;
            s.ts = ts; s.te = te; s.w = 0.f; s.sdf = q.sdf;
            s.n[0] = nsm[0]; s.n[1] = nsm[1]; s.n[2] = nsm[2];
            s.albedo[0] = mat[0]; s.albedo[1] = mat[1]; s.albedo[2] = mat[2];
            s.rough = mat[3]; s.metal = mat[4];
            samples[i] = s;
            IaSampleAux a;
            a.rgb[0] = rgb[0]; a.rgb[1] = rgb[1]; a.rgb[2] = rgb[2];
            a.nw[0] = nw[0]; a.nw[1] = nw[1]; a.nw[2] = nw[2];
            a.slot = slot; a.pad = 0;
            aux[i] = a;
        }
    }
    if (lane == 0) {
        if (c_qg) atomicAdd(&counters[IA_CNT_QUERIES_GRAD], c_qg);
        if (c_rad) atomicAdd(&counters[IA_CNT_RAD_EVAL], c_rad);
        if (c_geo) atomicAdd(&counters[IA_CNT_GEO_EVAL], c_geo);
        if (c_qg) atomicAdd(&counters[IA_CNT_SKIN_FETCH], c_qg);
    }
    if (c_fetch) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], c_fetch);
}
#endif

// ------------------------------------------------------------------------------------------------
// k_prim_shade_wf: the same per-sample work on the WAVEFRONT machinery of the shading stage (ia_wavefront.cuh).  One
// persistent 512-thread CTA per SM takes tiles of WF_R = 1024 consecutive samples; their midpoints are the queries of one
// "round": Morton sort -> exact prune -> dense Broyden -> filter -> tensor-core geometry, then the arg-min root of every
// sample goes through the tensor-core shading batch (ia_warp_radiance16<PRIMARY>: feature + normal + rotation + radiance +
// material) in full 16-row batches.  The team kernel above spends 23 of its 27 ms in the per-sample query (13 chains in
// lock-step per 16 lanes, 283 M voxel fetches); here the chains are independent dense tasks.
#ifndef IA_PRIM_SHADE_WF
#define IA_PRIM_SHADE_WF 1
#endif
__global__ void __launch_bounds__(WF_THREADS, WF_CTAS_PER_SM)
k_prim_shade_wf(const __grid_constant__ IaFrame p, const float* __restrict__ hit_od, IaSample* __restrict__ samples,
                IaSampleAux* __restrict__ aux, long long sample_cap, const int* __restrict__ work,
                unsigned char* __restrict__ scratch, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char wf_smem[];
    WfShared& S = *reinterpret_cast<WfShared*>(wf_smem);
    wf_setup<WF_THREADS / 32>(p, S, scratch);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
    const long long n = min((long long)work[IA_W_NSAMPLES], sample_cap);
    const long long n_tiles = (n + WF_R - 1) / WF_R;
    unsigned c_live = 0;
    float nw_def[3];
    {
        const float up[3] = {0.f, 0.f, 1.f};      // invalid query: gradient (0, 0, 1) (snarf_deformer.py:192)
        ia_dir_s2w(p, up, nw_def);
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * WF_R;
        __syncthreads();
        if (tid == 0) { S.n_q = 0; S.n_gtask = 0; S.task_next = 0; S.n_btask = 0; S.n_gitask = 0; }
        __syncthreads();
        // ---- queries: the midpoint of every live sample of the tile
        for (int t = tid; t < WF_R; t += blockDim.x) {
            const long long i = base + t;
            const int slot = i < n ? aux[i].slot : -1;   // < 0: past the end / dead record of a ray that did not fit the pool
            S.qmask[t] = 0;
            if (slot >= 0) {
                const float* od = hit_od + (size_t)slot * 8;
                const float tm = (samples[i].ts + samples[i].te) / 2.0f;
#pragma unroll
                for (int k = 0; k < 3; k++) S.qx[k][t] = od[k] + od[3 + k] * tm;
                S.qlist[atomicAdd(&S.n_q, 1)] = (unsigned short)t;
                c_live++;
            }
        }
        __syncthreads();
        const int n_q = S.n_q;
        if (n_q == 0) continue;
#if WF_QSORT
        wf_sort_phase(p, S, n_q);
        __syncthreads();
#endif
        wf_prune_phase(p, S, n_q);
        __syncthreads();
        wf_broyden_phase(p, S);
        __syncthreads();
        wf_filter_phase(S, n_q);
        __syncthreads();
        wf_geometry_phase(p, S);
        __syncthreads();
        // ---- min SDF over the kept roots (snarf_deformer.py:242-259); the record of a sample without a root is final
        for (int t = tid; t < WF_R; t += blockDim.x) {
            const long long i = base + t;
            const int slot = i < n ? aux[i].slot : -1;
            if (slot < 0) continue;
            unsigned keep = S.qmask[t];
            float sdf = 1e5f;
            int best = -1;
            while (keep) {
                const int c = __ffs(keep) - 1;
                keep &= keep - 1;
                const float s = S.csdf[t * IA_N_INIT + c];
                if (s < sdf) { sdf = s; best = c; }
            }
            IaSample r = samples[i];
            r.w = 0.f; r.sdf = sdf;
            r.n[0] = 0.f; r.n[1] = 0.f; r.n[2] = 1.f;
            r.albedo[0] = 0.f; r.albedo[1] = 0.f; r.albedo[2] = 0.f; r.rough = 0.f; r.metal = 0.f;
            samples[i] = r;
            IaSampleAux a;
            a.rgb[0] = 0.f; a.rgb[1] = 0.f; a.rgb[2] = 0.f;
            a.nw[0] = nw_def[0]; a.nw[1] = nw_def[1]; a.nw[2] = nw_def[2];
            a.slot = slot; a.pad = 0;
            aux[i] = a;
            if (best >= 0) wf_push_gi(S, t, best, 0.f);
        }
        __syncthreads();
        // ---- shading batches
        {
            const int n_sh = S.n_gitask;
            const int per = (((n_sh + n_warps - 1) / n_warps) + 15) & ~15;    // whole 16-row batches per warp
            const int end = min(n_sh, (warp + 1) * per);
            float* xs = wf_xs(S) + warp * IA_SHADE_TILE;
            for (int b0 = warp * per; b0 < end; b0 += 16) {
                const int nb = min(16, end - b0);
                long long rec = 0;
                float x0 = 0.f, x1 = 0.f, x2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 1.f;
                if ((lane & 15) < nb) {
                    const int t = (int)(S.gitask[b0 + (lane & 15)].x & 0xffffu);
                    rec = base + t;
                    const float* cd = S.gixc + (b0 + (lane & 15)) * 3;
                    x0 = cd[0]; x1 = cd[1]; x2 = cd[2];
                    const float* od = hit_od + (size_t)aux[rec].slot * 8;
                    d0 = od[3]; d1 = od[4]; d2 = od[5];
                }
                float rgb[3], mat[5];
                ia_warp_radiance16<true>(p, S.lvl, wf_w(S), wf_w1f(p, S), xs, x0, x1, x2, d0, d1, d2, nb, rgb, mat, rec, samples, aux);
                if (lane < nb) {
#pragma unroll
                    for (int k = 0; k < 3; k++) { aux[rec].rgb[k] = rgb[k]; samples[rec].albedo[k] = mat[k]; }
                    samples[rec].rough = mat[3];
                    samples[rec].metal = mat[4];
                }
            }
            if (end > warp * per) wf_restore_geo_pads(S, warp);
            wf_count(S, WF_C_QG, (lane == 0 && end > warp * per) ? (unsigned)(end - warp * per) : 0u);
        }
    }
    wf_count(S, WF_C_Q, c_live);
    __syncthreads();
    if (tid == 0) {
        const unsigned long long q = S.cnt[WF_C_Q], f = S.cnt[WF_C_FETCH], g = S.cnt[WF_C_GEO], k = S.cnt[WF_C_SKIP], qg = S.cnt[WF_C_QG];
        if (q) atomicAdd(&counters[IA_CNT_QUERIES_GRAD], q);
        if (f) atomicAdd(&counters[IA_CNT_BROYDEN_FETCH], f);
        if (g + qg) atomicAdd(&counters[IA_CNT_GEO_EVAL], g + qg);
        if (k) atomicAdd(&counters[IA_CNT_CHAINS_SKIPPED], k);
        if (qg) { atomicAdd(&counters[IA_CNT_RAD_EVAL], qg); atomicAdd(&counters[IA_CNT_SKIN_FETCH], qg); }
    }
}

// weights (render_weight_from_alpha) and the 7 accumulations (models/volrend.py:952-1010), one thread per hit ray,
// samples in ray order (same summation order as the fused first generation)
__global__ void k_prim_accum(const __grid_constant__ IaFrame p, const int* __restrict__ hit_rays, float* __restrict__ hit_od,
                             const int* __restrict__ hit_info, IaSample* __restrict__ samples,
                             const IaSampleAux* __restrict__ aux, const int* __restrict__ work, ia_outputs out) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= work[IA_W_NHIT]) return;
    const int base = hit_info[slot * 2], n_iv = hit_info[slot * 2 + 1];
    float* od = hit_od + (size_t)slot * 8;
    const float far = od[6];
    float T = 1.f;
    float a_rgb[3] = {0, 0, 0}, a_n[3] = {0, 0, 0}, a_alb[3] = {0, 0, 0}, a_r = 0, a_m = 0, a_op = 0, a_dep = 0;
    for (int k = 0; k < n_iv; k++) {
        IaSample& s = samples[(size_t)base + k];
        const IaSampleAux& a = aux[(size_t)base + k];
        const float ts = s.ts, te = s.te, tm = (ts + te) / 2.0f;
        const float alpha = ia_alpha(s.sdf, te - ts, p.beta);
        const float w = T * alpha;
        T *= (1.f - alpha);
#pragma unroll
        for (int c3 = 0; c3 < 3; c3++) {
            a_rgb[c3] += w * a.rgb[c3];
            a_n[c3] += w * a.nw[c3];
            a_alb[c3] += w * s.albedo[c3];
        }
        a_r += w * s.rough; a_m += w * s.metal; a_op += w; a_dep += w * tm;
        s.w = w;
    }
    od[7] = a_op;
    const size_t r = (size_t)hit_rays[slot];
    if (out.comp_rgb) { out.comp_rgb[r * 3] = a_rgb[0]; out.comp_rgb[r * 3 + 1] = a_rgb[1]; out.comp_rgb[r * 3 + 2] = a_rgb[2]; }
    if (out.comp_normal) { out.comp_normal[r * 3] = a_n[0]; out.comp_normal[r * 3 + 1] = a_n[1]; out.comp_normal[r * 3 + 2] = a_n[2]; }
    if (out.comp_albedo) { out.comp_albedo[r * 3] = a_alb[0]; out.comp_albedo[r * 3 + 1] = a_alb[1]; out.comp_albedo[r * 3 + 2] = a_alb[2]; }
    if (out.opacity) out.opacity[r] = a_op;
    if (out.depth) out.depth[r] = a_dep + (1.0f - a_op) * far;
    if (out.comp_roughness) out.comp_roughness[r] = a_r;
    if (out.comp_metallic) out.comp_metallic[r] = a_m;
    if (out.num_samples) out.num_samples[r] = n_iv;
}

// ------------------------------------------------------------------------------------------------
// ray_resampling (cdf_resampling_kernel, cdf.cu:9-149) for one ray by one warp, parallel over the
// spp outputs.  The serial walk of the reference is equivalent to, per output j:
//   idx_j = first interval with u_j < cdf_next[idx]   (none -> background)
// and the zero-crossing snap only depends on i* = first crossing interval and j0 = first output at
// or past it whose interpolated SDF is negative.  Every float operation on a given
// output is the same as in the serial kernel.
// Generic strided input so it serves both the pipeline (IaSample AoS) and the op-level entry point.
struct IaResampleIn {
    const float* starts; const float* ends; const float* weights; const float* sdfs;
    int stride;  // in floats
};

__device__ __forceinline__ int ia_count_below(const float* __restrict__ u, int n, float c) {  // #{j : u_j < c}
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (u[mid] < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// cdf: per-warp shared scratch [IA_CAP].  Outputs may be NULL individually.
__device__ __forceinline__ void ia_warp_resample(const IaResampleIn& in, int steps, int spp, const float* __restrict__ u_table,
                                                 float* cdf, float transmittance, long long src_base,
                                                 float* __restrict__ o_t, float* __restrict__ o_off, long long* __restrict__ o_idx64,
                                                 int* __restrict__ o_src, float* __restrict__ o_w,
                                                 int* __restrict__ o_fg_counts, int* __restrict__ o_bg_count,
                                                 long long* __restrict__ o_surface, const bool perm = false,
                                                 const uint32_t perm_key = 0u, const long long o_stride = 1) {
    // o_t / o_src / o_w of output j go to position (perm ? ia_permute(j, spp, perm_key) : j) * o_stride: the light-major
    // layout of the shading stage (k_resample)
    const int lane = threadIdx.x & 31;
    const int st = in.stride;
    if (lane == 0) {
        float weights_sum = 0.0f;
        for (int j = 0; j < steps; j++) weights_sum += in.weights[(size_t)j * st];
        weights_sum += fmaxf(1.0f - weights_sum, 0.0f);
        float c = in.weights[0] / weights_sum;
        cdf[0] = c;
        for (int i = 1; i < steps; i++) { c += in.weights[(size_t)i * st] / weights_sum; cdf[i] = c; }
    }
    __syncwarp();
    // first crossing interval
    int istar = 0x7fffffff;
    for (int i = lane; i < steps - 1; i += 32) {
        float a = in.sdfs[(size_t)i * st], b = in.sdfs[(size_t)(i + 1) * st];
        if (a >= 0 && b < 0) istar = min(istar, i);
    }
    for (int o = 16; o > 0; o >>= 1) istar = min(istar, __shfl_xor_sync(0xffffffffu, istar, o));
    // pass 1: j0
    int j0 = 0x7fffffff;
    for (int j = lane; j < spp; j += 32) {
        float u = u_table[j];
        int idx = ia_count_below(cdf, steps, u);  // cdf non-decreasing: first idx with u < cdf[idx] == #{cdf <= u}
        // (ia_count_below counts cdf[i] < u; ties u == cdf[i] must NOT count as "u < cdf": fix up)
        while (idx < steps && !(u < cdf[idx])) idx++;
        if (idx < steps) {
            bool snap = idx > istar;
            if (idx == istar) {
                float s = in.starts[(size_t)idx * st], e = in.ends[(size_t)idx * st];
                float cp = idx ? cdf[idx - 1] : 0.0f, cn = cdf[idx];
                float scaling = (e - s) / (cn - cp);
                float offset = (u - cp) * scaling;
                float sp = in.sdfs[(size_t)idx * st], sn = in.sdfs[(size_t)(idx + 1) * st];
                float sdf_approx = sp + (sn - sp) * (offset / (e - s));
                snap = !(sdf_approx >= 0);
            }
            if (snap) j0 = min(j0, j);
        }
    }
    for (int o = 16; o > 0; o >>= 1) j0 = min(j0, __shfl_xor_sync(0xffffffffu, j0, o));
    // the t every snapped output repeats
    float t_snap = 0.f;
    if (j0 != 0x7fffffff) {
        int jj = j0 > 0 ? j0 - 1 : 0;
        float u = u_table[jj];
        int idx = ia_count_below(cdf, steps, u);
        while (idx < steps && !(u < cdf[idx])) idx++;
        float s = in.starts[(size_t)idx * st], e = in.ends[(size_t)idx * st];
        if (j0 > 0) {
            float cp = idx ? cdf[idx - 1] : 0.0f, cn = cdf[idx];
            float scaling = (e - s) / (cn - cp);
            float offset = (u - cp) * scaling;
            t_snap = offset + s;
        } else {
            t_snap = s;
        }
    }
    const int n_fg = ia_count_below(u_table, spp, cdf[steps - 1]);
    const int n_bg = spp - n_fg;
    if (o_bg_count && lane == 0) *o_bg_count = n_bg;
    if (o_surface && lane == 0) {
        // surface_idx is recorded when the walk advances past i*: that needs an output at or beyond
        // interval i*+1, or the walk ending there; the walk stops advancing once all spp are placed.
        // It advances past i* iff some output (or the end) lies beyond: last fg idx > i* or n_bg > 0.
        int last_idx = -1;
        if (n_fg > 0) {
            float u = u_table[n_fg - 1];
            last_idx = ia_count_below(cdf, steps, u);
            while (last_idx < steps && !(u < cdf[last_idx])) last_idx++;
        }
        bool passed = istar != 0x7fffffff && (last_idx > istar || n_bg > 0);
        *o_surface = passed ? (long long)istar + src_base : -1;
    }
    if (o_fg_counts) {
        for (int i = lane; i < steps; i += 32) {
            float cp = i ? cdf[i - 1] : 0.0f;
            o_fg_counts[i] = ia_count_below(u_table, spp, cdf[i]) - ia_count_below(u_table, spp, cp);
        }
    }
    const float end_last = in.ends[(size_t)(steps - 1) * st];
    // pass 2: outputs
    for (int j = lane; j < spp; j += 32) {
        float u = u_table[j];
        int idx = ia_count_below(cdf, steps, u);
        while (idx < steps && !(u < cdf[idx])) idx++;
        if (idx < steps) {
            float s = in.starts[(size_t)idx * st], e = in.ends[(size_t)idx * st];
            float cp = idx ? cdf[idx - 1] : 0.0f, cn = cdf[idx];
            float scaling = (e - s) / (cn - cp);
            float offset = (u - cp) * scaling;
            float t = offset + s;
            if (j >= j0) t = t_snap;
            const long long jo = (long long)(perm ? ia_permute((uint32_t)j, (uint32_t)spp, perm_key) : (uint32_t)j) * o_stride;
            if (o_t) o_t[jo] = t;
            if (o_off) o_off[j] = offset;
            if (o_idx64) o_idx64[j] = idx + src_base;
            if (o_src) o_src[jo] = (int)(idx + src_base);
            if (o_w) {
                int cnt = ia_count_below(u_table, spp, cn) - ia_count_below(u_table, spp, cp);
                o_w[jo] = in.weights[(size_t)idx * st] / (float)cnt;
            }
        } else {
            float offset = 10000.f;
            const long long jo = (long long)(perm ? ia_permute((uint32_t)j, (uint32_t)spp, perm_key) : (uint32_t)j) * o_stride;
            if (o_t) o_t[jo] = offset + end_last;
            if (o_off) o_off[j] = offset;
            if (o_idx64) o_idx64[j] = steps - 1 + src_base;
            if (o_src) o_src[jo] = -1;
            if (o_w) o_w[jo] = transmittance / (float)n_bg;
        }
    }
}

// Layout of the resampled streams (rs_t, rs_src, rs_w):
//   pixel-major  [slot][j]          mats / mis: sample j of hit ray `slot`
//   light-major  [kk][slot]         light-table modes with WF_LIGHT_MAJOR: kk = the light direction the keyed permutation
//                                   assigns to sample j of that ray (models/intrinsic_avatar.py:1355-1378).  The shading
//                                   stage then feeds bundles of PARALLEL rays from neighbouring pixels (ia_wavefront.cuh).
__global__ void __launch_bounds__(256) k_resample(const int* __restrict__ hit_info, const float* __restrict__ hit_od,
                                                  const IaSample* __restrict__ samples, const int* __restrict__ work,
                                                  int spp, const float* __restrict__ u_table, float* __restrict__ rs_t,
                                                  int* __restrict__ rs_src, float* __restrict__ rs_w, int light_major,
                                                  const int* __restrict__ hit_rays, long long ray_index_base, uint32_t seed) {
    __shared__ float cdf_s[8][IA_CAP];
    const int n_hit = work[IA_W_NHIT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int slot = blockIdx.x * 8 + warp; slot < n_hit; slot += gridDim.x * 8) {
        int base = hit_info[slot * 2], steps = hit_info[slot * 2 + 1];
        const size_t ob = light_major ? (size_t)slot : (size_t)slot * spp;
        const long long stride = light_major ? (long long)n_hit : 1;
        const uint32_t key = light_major ? ia_pixel_key(seed, (uint64_t)(ray_index_base + hit_rays[slot])) : 0u;
        if (steps == 0) {
            // no shading sample survived: all background with the ray's full transmittance
            for (int j = lane; j < spp; j += 32) {
                const size_t o = ob + (size_t)j * stride;   // (every position of the ray is written: the order does not matter)
                rs_t[o] = 0.f; rs_src[o] = -1; rs_w[o] = 1.0f / (float)spp;
            }
            continue;
        }
        const float* sp = reinterpret_cast<const float*>(samples + base);
        IaResampleIn in{sp + 0, sp + 1, sp + 2, sp + 3, (int)(sizeof(IaSample) / sizeof(float))};
        float trans = 1.0f - hit_od[(size_t)slot * 8 + 7];
        ia_warp_resample(in, steps, spp, u_table, cdf_s[warp], trans, (long long)base, rs_t + ob, nullptr, nullptr,
                         rs_src + ob, rs_w + ob, nullptr, nullptr, nullptr, light_major != 0, key, stride);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ia_srgb(float f) {
    float v = f <= 0.0031308f ? f * 12.92f : powf(fmaxf(f, 0.0031308f), 1.0f / 2.4f) * 1.055f - 0.055f;
    return fminf(fmaxf(v, 0.f), 1.f);
}

__global__ void k_composite(const __grid_constant__ IaFrame p, long long n, const float* __restrict__ acc6, ia_outputs out,
                            int primary_only, const float* __restrict__ vis, const float* __restrict__ bg_rgb) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (out.visibility) out.visibility[r] = vis ? vis[r] : 0.f;  // uniform_light only (models/intrinsic_avatar.py:1427-1432)
    float op = out.opacity ? out.opacity[r] : 0.f;
    float bgm = (p.background[0] + p.background[1] + p.background[2]) / 3.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        // primary-only: the background colour, or with add_emitter the envmap along the ray (acc6 may hold a hit ray's zero)
        float phys = primary_only ? bg_rgb[r * 3 + k] : acc6[r * 6 + k];
        float dem = primary_only ? bg_rgb[r * 3 + k] : acc6[r * 6 + 3 + k];
        if (out.comp_rgb_phys) out.comp_rgb_phys[r * 3 + k] = phys;
        if (out.comp_demod_phys) out.comp_demod_phys[r * 3 + k] = dem;
        if (out.comp_rgb_phys_full) out.comp_rgb_phys_full[r * 3 + k] = ia_srgb(phys);
        if (out.comp_demod_phys_full) out.comp_demod_phys_full[r * 3 + k] = ia_srgb(dem);
        if (out.comp_rgb_full && out.comp_rgb) out.comp_rgb_full[r * 3 + k] = ia_srgb(out.comp_rgb[r * 3 + k] + p.background[k] * (1.0f - op));
        if (out.comp_albedo_full && out.comp_albedo) out.comp_albedo_full[r * 3 + k] = out.comp_albedo[r * 3 + k] + 0.f * (1.0f - op);
    }
    if (out.comp_roughness_full && out.comp_roughness) out.comp_roughness_full[r] = out.comp_roughness[r] + bgm * (1.0f - op);
    if (out.comp_metallic_full && out.comp_metallic) out.comp_metallic_full[r] = out.comp_metallic[r] + bgm * (1.0f - op);
}

// ================================================================================================
static int ia_wf_scratch(ia_ctx* c) {
    if (!c->d_wf_scratch) return ia_realloc(&c->d_wf_scratch, (size_t)c->n_sm * WF_CTAS_PER_SM * WF_SCRATCH_BYTES);
    return IA_OK;
}

static int ia_ws_reserve(ia_ctx* c, int64_t n_rays, int spp, bool need_pbr) {
    if (n_rays > c->ws_rays) {
        if (ia_realloc(&c->d_hit_rays, (size_t)n_rays) || ia_realloc(&c->d_hit_od, (size_t)n_rays * 8) ||
            ia_realloc(&c->d_hit_info, (size_t)n_rays * 2) || ia_realloc(&c->d_acc, (size_t)n_rays * 6) ||
            ia_realloc(&c->d_vis, (size_t)n_rays) || ia_realloc(&c->d_bg, (size_t)n_rays * 3) ||
            ia_realloc(&c->d_hit_flag, (size_t)n_rays) || ia_realloc(&c->d_blk_cnt, (size_t)(n_rays + 255) / 256))
            return IA_ECUDA;
        c->ws_rays = n_rays;
    }
    // ~38 samples per HIT ray observed; 64 per ray covers full coverage.  Not a bound: a frame that needs more reports
    // IA_CNT_OVERFLOW and the host grows the pool (ia_reserve_samples) and renders again.
    if (int e = ia_reserve_samples(c, std::max<int64_t>(n_rays * 64, 1 << 16))) return e;
    if (need_pbr) {
        // worst case every ray hits; grown lazily to n_rays * spp (805 MB at 512^2 x 1024 for 25 % hits
        // would suffice, but the hit count is only known on the device)
        int64_t want = n_rays * (int64_t)spp;
        if (want > c->ws_resamples) {
            if (ia_realloc(&c->d_rs_t, (size_t)want) || ia_realloc(&c->d_rs_w, (size_t)want) ||
                ia_realloc(&c->d_rs_src, (size_t)want))
                return IA_ECUDA;
            c->ws_resamples = want;
        }
    }
    return IA_OK;
}

extern "C" int ia_reserve_samples(ia_ctx* c, int64_t n_samples) {
    IA_REQUIRE(c && n_samples >= 0, IA_EINVAL, "ia_reserve_samples: bad argument");
    IA_REQUIRE(n_samples < (1ll << 31), IA_EINVAL, "ia_reserve_samples: the pool is indexed with 32-bit integers");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    if (n_samples > c->ws_samples) {
        IA_CHECK_CUDA(cudaDeviceSynchronize());  // a render using the old pool may still be in flight
        if (ia_realloc(&c->d_samples, (size_t)n_samples) || ia_realloc(&c->d_samples_aux, (size_t)n_samples)) return IA_ECUDA;
        c->ws_samples = n_samples;
    }
    return IA_OK;
}

template <int MODE>
static int ia_launch_shade(ia_ctx* c, bool gi, int64_t ray_index_base, uint32_t seed, cudaStream_t st) {
    WfShadePolicy<MODE> pol;
    pol.p = nullptr; pol.hit_rays = c->d_hit_rays; pol.hit_od = c->d_hit_od; pol.samples = c->d_samples;
    pol.rs_t = c->d_rs_t; pol.rs_src = c->d_rs_src; pol.rs_w = c->d_rs_w; pol.work = c->d_work; pol.spp = c->spp;
    pol.ray_index_base = ray_index_base; pol.seed = seed; pol.light_dir_s = c->d_light_dir_s;
    pol.light_em = c->d_light_em; pol.light_pdf = c->d_light_pdf; pol.acc6 = c->d_acc; pol.n_total = 0; pol.gi = gi;
    pol.env = c->env; pol.vis = MODE == IA_MODE_UNIFORM_LIGHT ? c->d_vis : nullptr; pol.bg_rgb = c->d_bg;
    if (gi) {
        IA_CHECK_CUDA(cudaFuncSetAttribute(k_shade_wf<true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES(true)));
        k_shade_wf<true, MODE><<<c->n_sm * WF_CTAS_PER_SM, WF_THREADS, WF_SMEM_BYTES(true), st>>>(c->f, pol, c->d_wf_scratch, c->d_counters);
    } else {
        IA_CHECK_CUDA(cudaFuncSetAttribute(k_shade_wf<false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES(false)));
        k_shade_wf<false, MODE><<<c->n_sm * WF_CTAS_PER_SM, WF_THREADS, WF_SMEM_BYTES(false), st>>>(c->f, pol, c->d_wf_scratch, c->d_counters);
    }
    return IA_OK;
}

extern "C" int ia_render(ia_ctx* c, const float* d_rays, int64_t n_rays, int64_t ray_index_base, int flags, uint32_t seed,
                         const ia_outputs* out, void* stream) {
    IA_REQUIRE(c && d_rays && out, IA_EINVAL, "ia_render: NULL argument");
    IA_REQUIRE(c->have_fields && c->have_pose && c->have_occ && c->have_cfg, IA_ESTATE,
               "ia_render: fields / pose / occupancy / render config must be set");
    const bool primary_only = flags & IA_RENDER_PRIMARY_ONLY;
    const bool gi = flags & IA_RENDER_GI;
    const int mode = (flags & IA_RENDER_MODE_MASK) >> IA_RENDER_MODE_SHIFT;
    // add_emitter also applies to a primary-only (albedo_only) render: comp_rgb_phys is then the envmap along every ray
    // (models/intrinsic_avatar.py:1468-1478); it needs the light tables of ia_set_light*
    const bool add_emitter = (flags & IA_RENDER_ADD_EMITTER) && c->have_light;
    IA_REQUIRE(primary_only || c->have_light, IA_ESTATE, "ia_render: call ia_set_light first");
    IA_REQUIRE(primary_only || (mode == IA_MODE_UNIFORM_LIGHT) == c->light_uniform, IA_ESTATE,
               "ia_render: render_mode uniform_light needs ia_set_light_uniform, the other modes ia_set_light");
    IA_REQUIRE(c->f.occ_res * c->f.occ_res * c->f.occ_res / 8 <= 64 * 1024, IA_EINVAL, "ia_render: occupancy grid too large for shared memory");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    if (n_rays == 0) return IA_OK;
    IA_REQUIRE(n_rays < (1ll << 31), IA_EINVAL, "ia_render: too many rays");
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = ia_ws_reserve(c, n_rays, c->spp, !primary_only)) return e;
    IA_CHECK_CUDA(cudaMemsetAsync(c->d_work, 0, 8 * sizeof(int), st));
    IA_CHECK_CUDA(cudaMemsetAsync(c->d_counters, 0, IA_N_COUNTERS * sizeof(unsigned long long), st));
    IA_STAGE_BEGIN(c, IA_STAGE_SETUP, st);
    {
        const unsigned nb = (unsigned)((n_rays + 255) / 256);
        k_primary_setup<<<nb, 256, 0, st>>>(c->f, d_rays, n_rays, c->d_hit_flag, c->d_blk_cnt, *out, c->d_acc, c->d_bg, c->env,
                                            add_emitter ? 1 : 0);
        k_hit_scan<<<1, 1024, 0, st>>>(c->d_blk_cnt, (int)nb, c->d_work);
        k_hit_compact<<<nb, 256, 0, st>>>(c->f, d_rays, n_rays, c->d_hit_flag, c->d_blk_cnt, c->d_hit_rays, c->d_hit_od);
    }
    IA_STAGE_END(c, IA_STAGE_SETUP, st, 3);
    IA_LAUNCH_CHECK();
    {
        size_t sm1 = IA_GEO_END * sizeof(float) + (IA_PRIMARY_THREADS / IA_TEAM) * sizeof(IaPrimarySmem);
        size_t sm2 = IA_PRIM_SHADE_SMEM;
        IA_CHECK_CUDA(cudaFuncSetAttribute(k_prim_edges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
        IA_CHECK_CUDA(cudaFuncSetAttribute(k_prim_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        IA_CHECK_CUDA(cudaFuncSetAttribute(k_prim_shade_wf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES_SW(WF_THREADS / 32)));
        if (int e = ia_wf_scratch(c)) return e;
        IA_STAGE_BEGIN(c, IA_STAGE_PRIMARY, st);
        k_prim_edges<<<c->n_sm * 2, IA_PRIMARY_THREADS, sm1, st>>>(c->f, c->d_hit_od, c->d_hit_info, c->d_samples, c->d_samples_aux,
                                                                   c->ws_samples, c->d_work, c->d_counters);
#if IA_PRIM_SHADE_WF
        k_prim_shade_wf<<<c->n_sm * WF_CTAS_PER_SM, WF_THREADS, WF_SMEM_BYTES_SW(WF_THREADS / 32), st>>>(c->f, c->d_hit_od, c->d_samples, c->d_samples_aux,
                                                                                         c->ws_samples, c->d_work, c->d_wf_scratch, c->d_counters);
#else
        k_prim_shade<<<c->n_sm * IA_PRIM_SHADE_CTAS, IA_PRIMARY_THREADS, sm2, st>>>(c->f, c->d_hit_od, c->d_samples, c->d_samples_aux, c->ws_samples,
                                                                   c->d_work, c->d_counters);
#endif
        k_prim_accum<<<(unsigned)((n_rays + 127) / 128), 128, 0, st>>>(c->f, c->d_hit_rays, c->d_hit_od, c->d_hit_info, c->d_samples,
                                                                       c->d_samples_aux, c->d_work, *out);
        IA_STAGE_END(c, IA_STAGE_PRIMARY, st, 3);
        IA_LAUNCH_CHECK();
    }
    IA_CHECK_CUDA(cudaMemcpyAsync(c->d_counters + IA_CNT_PRIMARY_BASE, c->d_counters, IA_CNT_PRIMARY_BASE * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToDevice, st));
    if (!primary_only) {
        IA_STAGE_BEGIN(c, IA_STAGE_RESAMPLE, st);
        k_resample<<<c->n_sm * 4, 256, 0, st>>>(c->d_hit_info, c->d_hit_od, c->d_samples, c->d_work, c->spp, c->d_u_table,
                                                c->d_rs_t, c->d_rs_src, c->d_rs_w, (WF_LIGHT_MAJOR && mode <= IA_MODE_UNIFORM_LIGHT) ? 1 : 0,
                                                c->d_hit_rays, ray_index_base, seed);
        IA_STAGE_END(c, IA_STAGE_RESAMPLE, st, 1);
        IA_LAUNCH_CHECK();
        if (mode == IA_MODE_UNIFORM_LIGHT) IA_CHECK_CUDA(cudaMemsetAsync(c->d_vis, 0, (size_t)n_rays * sizeof(float), st));
        IA_STAGE_BEGIN(c, IA_STAGE_SHADE, st);
        {
            // wavefront integrator: one persistent CTA per SM
            IA_REQUIRE((long long)n_rays * c->spp < (1ll << 32), IA_EINVAL, "ia_render: n_rays * spp must be < 2^32");
            if (int e = ia_wf_scratch(c)) return e;
            int e = IA_OK;
            switch (mode) {
                case IA_MODE_LIGHT: e = ia_launch_shade<IA_MODE_LIGHT>(c, gi, ray_index_base, seed, st); break;
                case IA_MODE_UNIFORM_LIGHT: e = ia_launch_shade<IA_MODE_UNIFORM_LIGHT>(c, gi, ray_index_base, seed, st); break;
                case IA_MODE_MATS: e = ia_launch_shade<IA_MODE_MATS>(c, gi, ray_index_base, seed, st); break;
                default: e = ia_launch_shade<IA_MODE_MIS>(c, gi, ray_index_base, seed, st); break;
            }
            if (e) return e;
        }
        IA_STAGE_END(c, IA_STAGE_SHADE, st, 1);
        IA_LAUNCH_CHECK();
    }
    IA_STAGE_BEGIN(c, IA_STAGE_COMPOSITE, st);
    k_composite<<<(unsigned)((n_rays + 255) / 256), 256, 0, st>>>(c->f, n_rays, c->d_acc, *out, primary_only ? 1 : 0,
                                                                  (!primary_only && mode == IA_MODE_UNIFORM_LIGHT) ? c->d_vis : nullptr, c->d_bg);
    IA_STAGE_END(c, IA_STAGE_COMPOSITE, st, 1);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_get_counters(ia_ctx* c, uint64_t* h, void* stream) {
    IA_REQUIRE(c && h, IA_EINVAL, "ia_get_counters: NULL argument");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    unsigned long long tmp[IA_N_COUNTERS];
    int work[8];
    IA_CHECK_CUDA(cudaMemcpyAsync(tmp, c->d_counters, sizeof(tmp), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    IA_CHECK_CUDA(cudaMemcpyAsync(work, c->d_work, sizeof(work), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    IA_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < IA_N_COUNTERS; i++) h[i] = tmp[i];
    h[IA_CNT_HIT_RAYS] = (uint64_t)work[IA_W_NHIT];
    return IA_OK;
}

// ================================================================================================
// remaining op-level entry points
__global__ void k_op_resample(const int* __restrict__ packed, const float* __restrict__ starts, const float* __restrict__ ends,
                              const float* __restrict__ weights, const float* __restrict__ sdfs, long long n_rays, int spp,
                              const int* __restrict__ rpacked, const float* __restrict__ u_table, float* __restrict__ ts,
                              float* __restrict__ offs, long long* __restrict__ idx, int* __restrict__ fg, int* __restrict__ bg,
                              long long* __restrict__ surf) {
    __shared__ float cdf_s[8][IA_CAP];
    const int warp = threadIdx.x >> 5;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < n_rays; r += (long long)gridDim.x * 8) {
        int base = packed[r * 2], steps = packed[r * 2 + 1];
        if (steps == 0) continue;
        if (steps > IA_CAP) continue;  // guarded by the host
        int rb = rpacked[r * 2];
        IaResampleIn in{starts + base, ends + base, weights + base, sdfs + base, 1};
        ia_warp_resample(in, steps, spp, u_table, cdf_s[warp], 0.f, (long long)base, ts + rb, offs + rb, idx + rb, nullptr,
                         nullptr, fg + base, bg + r, surf + r);
        __syncwarp();
    }
}

__global__ void k_make_u_table(float* u, int spp) {
    if (threadIdx.x || blockIdx.x) return;
    float step = (1.0f - 1.0 / spp) / (spp - 1);
    float cu = 1.0 / (2 * spp);
    for (int j = 0; j < spp; j++) { u[j] = cu; cu += step; }
}

extern "C" int ia_op_ray_resampling(ia_ctx* c, const int32_t* d_packed, const float* d_starts, const float* d_ends,
                                    const float* d_weights, const float* d_sdfs, int64_t n_rays, int spp,
                                    const int32_t* d_rpacked, float* d_ts, float* d_offs, int64_t* d_idx, int32_t* d_fg,
                                    int32_t* d_bg, int64_t* d_surf, void* stream) {
    IA_REQUIRE(c && d_packed && d_starts && d_ends && d_weights && d_sdfs && d_rpacked && d_ts && d_offs && d_idx && d_fg &&
                   d_bg && d_surf, IA_EINVAL, "ia_op_ray_resampling: NULL argument");
    IA_REQUIRE(spp > 1, IA_EINVAL, "ia_op_ray_resampling: n_samples must be > 1");
    if (n_rays == 0) return IA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* u = nullptr;
    IA_CHECK_CUDA(cudaMallocAsync((void**)&u, spp * sizeof(float), st));
    k_make_u_table<<<1, 1, 0, st>>>(u, spp);
    IA_CHECK_CUDA(cudaMemsetAsync(d_surf, 0xff, n_rays * sizeof(int64_t), st));
    k_op_resample<<<(unsigned)std::min<int64_t>((n_rays + 7) / 8, 65535), 256, 0, st>>>(
        d_packed, d_starts, d_ends, d_weights, d_sdfs, n_rays, spp, d_rpacked, u, d_ts, d_offs, (long long*)d_idx, d_fg, d_bg,
        (long long*)d_surf);
    IA_LAUNCH_CHECK();
    IA_CHECK_CUDA(cudaFreeAsync(u, st));
    return IA_OK;
}

__global__ void k_op_merge(const int* __restrict__ packed, const float* __restrict__ vals, const uint8_t* __restrict__ il,
                           const uint8_t* __restrict__ ir, const float* __restrict__ weights, long long n_rays,
                           const int* __restrict__ rpacked, float* __restrict__ ovals, float* __restrict__ odists,
                           uint8_t* __restrict__ oil, uint8_t* __restrict__ oir, uint8_t* __restrict__ oisr,
                           uint8_t* __restrict__ oisfg) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    int base = packed[r * 2], steps = packed[r * 2 + 1];
    if (steps == 0 || steps > IA_CAP - 16) return;
    int rb = rpacked[r * 2];
    uint8_t fl[IA_CAP], ofl[IA_CAP];
    for (int i = 0; i < steps; i++) fl[i] = (il[base + i] ? 1 : 0) | (ir[base + i] ? 2 : 0);
    int n = ia_merge_resample(vals + base, fl, weights + base, steps, rpacked[r * 2 + 1] - steps, ovals + rb, ofl, odists + rb);
    for (int i = 0; i < n; i++) {
        oil[rb + i] = ofl[i] & 1; oir[rb + i] = (ofl[i] >> 1) & 1; oisr[rb + i] = (ofl[i] >> 2) & 1; oisfg[rb + i] = 1;
    }
}

extern "C" int ia_op_ray_resampling_merge(ia_ctx* c, const int32_t* d_packed, const float* d_vals, const uint8_t* d_il,
                                          const uint8_t* d_ir, const float* d_weights, int64_t n_rays, const int32_t* d_rpacked,
                                          float* d_ovals, float* d_odists, uint8_t* d_oil, uint8_t* d_oir, uint8_t* d_oisr,
                                          uint8_t* d_oisfg, void* stream) {
    IA_REQUIRE(c && d_packed && d_vals && d_il && d_ir && d_weights && d_rpacked && d_ovals && d_odists && d_oil && d_oir &&
                   d_oisr && d_oisfg, IA_EINVAL, "ia_op_ray_resampling_merge: NULL argument");
    if (n_rays == 0) return IA_OK;
    k_op_merge<<<(unsigned)((n_rays + 63) / 64), 64, 0, (cudaStream_t)stream>>>(d_packed, d_vals, d_il, d_ir, d_weights, n_rays,
                                                                               d_rpacked, d_ovals, d_odists, d_oil, d_oir, d_oisr,
                                                                               d_oisfg);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// cdf_resampling_sdf_fine_kernel (cdf.cu:536-638), thread per ray (op-level twin of the lazy tracer)
__global__ void k_op_sdf_fine(const int* __restrict__ packed, const float* __restrict__ starts_all, const float* __restrict__ ends_all,
                              const float* __restrict__ alphas_all, const float* __restrict__ sdfs_all, long long n_rays,
                              const int* __restrict__ rpacked, float* __restrict__ rs_all, float* __restrict__ re_all,
                              uint8_t* __restrict__ isfg_all) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const int base = packed[i * 2], steps = packed[i * 2 + 1];
    const int rb = rpacked[i * 2], rsteps = rpacked[i * 2 + 1];
    if (steps == 0) return;
    const float *starts = starts_all + base, *ends = ends_all + base, *alphas = alphas_all + base, *sdfs = sdfs_all + base;
    float *rs = rs_all + rb, *re = re_all + rb;
    uint8_t* isfg = isfg_all + rb;
    int idx = 0;
    float sdf_prev = sdfs[0];
    bool found = false;
    while (idx < steps) {
        idx += 1;
        if (idx >= steps) break;
        if (sdf_prev >= 0 && sdfs[idx] < 0 && !found) { idx -= 1; found = true; break; }
        sdf_prev = sdfs[idx];
    }
    if (!found) return;
    int num_bins = rsteps + 1;
    float cdf_step_size = (1.0f - 1.0 / num_bins) / rsteps;
    int j = 0;
    float trans = 1.0f;
    float weight = alphas[idx];
    trans *= (1.0f - alphas[idx]);
    float cdf_prev = 0.0f, cdf_next = weight;
    float cdf_u = 1.0 / (2 * num_bins);
    while (j < num_bins && idx < steps) {
        if (cdf_u < cdf_next) {
            float scaling = (ends[idx] - starts[idx]) / (cdf_next - cdf_prev);
            float t = (cdf_u - cdf_prev) * scaling + starts[idx];
            if (j < num_bins - 1) rs[j] = t;
            if (j > 0) { re[j - 1] = t; isfg[j - 1] = 1; }
            cdf_u += cdf_step_size;
            j += 1;
        } else {
            idx += 1;
            if (idx >= steps) break;
            weight = trans * alphas[idx];
            trans *= (1.0f - alphas[idx]);
            cdf_prev = cdf_next;
            cdf_next += weight;
        }
    }
}

extern "C" int ia_op_ray_resampling_sdf_fine(ia_ctx* c, const int32_t* d_packed, const float* d_starts, const float* d_ends,
                                             const float* d_alphas, const float* d_sdfs, int64_t n_rays, const int32_t* d_rpacked,
                                             float* d_rs, float* d_re, uint8_t* d_isfg, void* stream) {
    IA_REQUIRE(c && d_packed && d_starts && d_ends && d_alphas && d_sdfs && d_rpacked && d_rs && d_re && d_isfg, IA_EINVAL,
               "ia_op_ray_resampling_sdf_fine: NULL argument");
    if (n_rays == 0) return IA_OK;
    k_op_sdf_fine<<<(unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_packed, d_starts, d_ends, d_alphas, d_sdfs,
                                                                                     n_rays, d_rpacked, d_rs, d_re, d_isfg);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// cdf_resampling_fine_kernel (cdf.cu:403-478): thread per ray.  The CDF of the compositing weights of all samples,
// n + 1 stratified positions -> n consecutive intervals (op-level twin of the wavefront tracer with
// zero_crossing_search = false).
__global__ void k_op_fine(const int* __restrict__ packed, const float* __restrict__ starts_all, const float* __restrict__ ends_all,
                          const float* __restrict__ weights_all, long long n_rays, const int* __restrict__ rpacked,
                          float* __restrict__ rs_all, float* __restrict__ re_all, uint8_t* __restrict__ isfg_all) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const int base = packed[i * 2], steps = packed[i * 2 + 1];
    const int rb = rpacked[i * 2], rsteps = rpacked[i * 2 + 1];
    if (steps == 0) return;
    const float *starts = starts_all + base, *ends = ends_all + base, *weights = weights_all + base;
    float *rs = rs_all + rb, *re = re_all + rb;
    uint8_t* isfg = isfg_all + rb;
    float weights_sum = 0.0f;
    for (int j = 0; j < steps; j++) weights_sum += weights[j];
    weights_sum += fmaxf(1.0f - weights_sum, 0.0f);
    const int num_bins = rsteps + 1;
    const float cdf_step_size = (1.0f - 1.0 / num_bins) / rsteps;
    int idx = 0, j = 0;
    float cdf_prev = 0.0f, cdf_next = weights[idx] / weights_sum;
    float cdf_u = 1.0 / (2 * num_bins);
    while (j < num_bins && idx < steps) {
        if (cdf_u < cdf_next) {
            float scaling = (ends[idx] - starts[idx]) / (cdf_next - cdf_prev);
            float t = (cdf_u - cdf_prev) * scaling + starts[idx];
            if (j < num_bins - 1) rs[j] = t;
            if (j > 0) { re[j - 1] = t; isfg[j - 1] = 1; }
            cdf_u += cdf_step_size;
            j += 1;
        } else {
            idx += 1;
            if (idx >= steps) break;
            cdf_prev = cdf_next;
            cdf_next += weights[idx] / weights_sum;
        }
    }
}

extern "C" int ia_op_ray_resampling_fine(ia_ctx* c, const int32_t* d_packed, const float* d_starts, const float* d_ends,
                                         const float* d_weights, int64_t n_rays, const int32_t* d_rpacked, float* d_rs,
                                         float* d_re, uint8_t* d_isfg, void* stream) {
    IA_REQUIRE(c && d_packed && d_starts && d_ends && d_weights && d_rpacked && d_rs && d_re && d_isfg, IA_EINVAL,
               "ia_op_ray_resampling_fine: NULL argument");
    if (n_rays == 0) return IA_OK;
    k_op_fine<<<(unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_packed, d_starts, d_ends, d_weights, n_rays,
                                                                                 d_rpacked, d_rs, d_re, d_isfg);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

__global__ void k_op_unpack_info(const int* __restrict__ packed, long long n_rays, long long* __restrict__ ray_indices) {
    // one warp per ray, coalesced fill (unpack_info_kernel, pack.cu:7-28)
    long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rays) return;
    int base = packed[r * 2], steps = packed[r * 2 + 1];
    for (int j = threadIdx.x & 31; j < steps; j += 32) ray_indices[base + j] = r;
}

extern "C" int ia_op_unpack_info(ia_ctx* c, const int32_t* d_packed, int64_t n_rays, int64_t* d_ray_indices, void* stream) {
    IA_REQUIRE(c && d_packed, IA_EINVAL, "ia_op_unpack_info: NULL argument");
    if (n_rays == 0 || !d_ray_indices) return IA_OK;  // a zero-sample output buffer has no address
    k_op_unpack_info<<<(unsigned)((n_rays * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_packed, n_rays,
                                                                                           (long long*)d_ray_indices);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

extern "C" int ia_op_secondary(ia_ctx* c, const float* d_o, const float* d_d, int64_t n, int gi, float* d_T, float* d_rgb,
                               void* stream) {
    IA_REQUIRE(c && d_o && d_d && d_T, IA_EINVAL, "ia_op_secondary: NULL argument");
    IA_REQUIRE(c->have_fields && c->have_pose && c->have_occ && c->have_cfg, IA_ESTATE, "ia_op_secondary: state not set");
    if (n == 0) return IA_OK;
    {
        IA_REQUIRE(n < (1ll << 32), IA_EINVAL, "ia_op_secondary: too many rays");
        IA_CHECK_CUDA(cudaMemsetAsync(c->d_work, 0, 8 * sizeof(int), (cudaStream_t)stream));
        WfRaysPolicy pol;
        pol.ro = d_o; pol.rd = d_d; pol.n = n; pol.T_out = d_T; pol.rgb_out = d_rgb; pol.work = c->d_work;
        if (int e = ia_wf_scratch(c)) return e;
        int wf_blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + WF_FEED - 1) / WF_FEED, (int64_t)c->n_sm * WF_CTAS_PER_SM));
        if (gi) {
            IA_CHECK_CUDA(cudaFuncSetAttribute(k_rays_wf<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES(true)));
            k_rays_wf<true><<<wf_blocks, WF_THREADS, WF_SMEM_BYTES(true), (cudaStream_t)stream>>>(c->f, pol, c->d_wf_scratch, c->d_counters);
        } else {
            IA_CHECK_CUDA(cudaFuncSetAttribute(k_rays_wf<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES(false)));
            k_rays_wf<false><<<wf_blocks, WF_THREADS, WF_SMEM_BYTES(false), (cudaStream_t)stream>>>(c->f, pol, c->d_wf_scratch, c->d_counters);
        }
    }
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// op-level twin of the wavefront kernel's geometry phase: canonical SDF of n points on the tensor cores
// (ia_warp_geometry: hash grid -> 35 -> 64 layer as 3xTF32 mma -> softplus -> sdf row)
__global__ void __launch_bounds__(256) k_op_geometry(const __grid_constant__ IaFrame p, const float* __restrict__ xc, long long n,
                                                     float* __restrict__ sdf) {
    extern __shared__ __align__(16) float smem[];
    float* w = smem;
    float4* w1f = reinterpret_cast<float4*>(smem + IA_GEOC_END);
    float* xs_all = reinterpret_cast<float*>(w1f + IA_GEO_KSTEPS * 8 * 32);
    __shared__ IaLevel lvl[IA_N_LEVELS];
    ia_stage_geoc(w, p.mlp);
    ia_stage_bfrag(w1f, IA_GEO_KSTEPS, 8, [&](int k, int nn) { return ia_geo_w1(p.mlp, k, nn); });
    for (int i = threadIdx.x; i < (int)(blockDim.x >> 5) * 16 * IA_GEO_LD; i += blockDim.x) xs_all[i] = 0.f;
    if (threadIdx.x < IA_N_LEVELS) lvl[threadIdx.x] = ia_level(p, threadIdx.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = xs_all + warp * 16 * IA_GEO_LD;
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long b0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 16; b0 < n; b0 += n_warps * 16) {
        const int nb = (int)min((long long)16, n - b0);
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (lane < nb) { x0 = xc[(b0 + lane) * 3]; x1 = xc[(b0 + lane) * 3 + 1]; x2 = xc[(b0 + lane) * 3 + 2]; }
        const float s = ia_warp_geometry<16>(p, lvl, w, w1f, xs, x0, x1, x2, nb);
        if (lane < nb) sdf[b0 + lane] = s;
    }
}

extern "C" int ia_op_geometry(ia_ctx* c, const float* d_xc, int64_t n, float* d_sdf, void* stream) {
    IA_REQUIRE(c && d_xc && d_sdf, IA_EINVAL, "ia_op_geometry: NULL argument");
    IA_REQUIRE(c->have_fields, IA_ESTATE, "ia_op_geometry: call ia_set_fields first");
    if (n == 0) return IA_OK;
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const size_t sm = IA_GEOC_END * sizeof(float) + IA_GEO_KSTEPS * 8 * 32 * sizeof(float4) + 8 * 16 * IA_GEO_LD * sizeof(float);
    IA_CHECK_CUDA(cudaFuncSetAttribute(k_op_geometry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)c->n_sm * 2));
    k_op_geometry<<<blocks, 256, sm, (cudaStream_t)stream>>>(c->f, d_xc, n, d_sdf);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

__global__ void k_op_brdf(const float* __restrict__ wi, const float* __restrict__ nn, const float* __restrict__ wo,
                          const float* __restrict__ rough, const float* __restrict__ albedo, const float* __restrict__ metal,
                          long long n, float* __restrict__ diff, float* __restrict__ spec) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[3] = {wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]}, b[3] = {nn[i * 3], nn[i * 3 + 1], nn[i * 3 + 2]};
    float cc[3] = {wo[i * 3], wo[i * 3 + 1], wo[i * 3 + 2]}, al[3] = {albedo[i * 3], albedo[i * 3 + 1], albedo[i * 3 + 2]};
    float df, sp[3];
    ia_brdf_multilobe(a, b, cc, rough[i], al, metal[i], df, sp);
    diff[i] = df;
    spec[i * 3] = sp[0]; spec[i * 3 + 1] = sp[1]; spec[i * 3 + 2] = sp[2];
}

extern "C" int ia_op_brdf(ia_ctx* c, const float* d_wi, const float* d_n, const float* d_wo, const float* d_rough,
                          const float* d_albedo, const float* d_metal, int64_t n, float* d_diff, float* d_spec, void* stream) {
    IA_REQUIRE(c && d_wi && d_n && d_wo && d_rough && d_albedo && d_metal && d_diff && d_spec, IA_EINVAL, "ia_op_brdf: NULL");
    if (n == 0) return IA_OK;
    k_op_brdf<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_wi, d_n, d_wo, d_rough, d_albedo, d_metal, n, d_diff,
                                                                           d_spec);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

__global__ void k_op_bsdf(const float* __restrict__ wi, const float* __restrict__ nn, const float* __restrict__ rough,
                          const float* __restrict__ albedo, const float* __restrict__ metal, const float* __restrict__ sample,
                          const float* __restrict__ woq, long long n, float* __restrict__ wo, float* __restrict__ pdf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a[3] = {wi[i * 3], wi[i * 3 + 1], wi[i * 3 + 2]}, b[3] = {nn[i * 3], nn[i * 3 + 1], nn[i * 3 + 2]};
    const float al[3] = {albedo[i * 3], albedo[i * 3 + 1], albedo[i * 3 + 2]};
    float d[3] = {0.f, 0.f, 1.f};
    if (wo && sample) {
        ia_multilobe_sample(b, a, rough[i], al, metal[i], sample[i * 2], sample[i * 2 + 1], d);
        wo[i * 3] = d[0]; wo[i * 3 + 1] = d[1]; wo[i * 3 + 2] = d[2];
    }
    if (pdf) {
        if (woq) { d[0] = woq[i * 3]; d[1] = woq[i * 3 + 1]; d[2] = woq[i * 3 + 2]; }
        pdf[i] = ia_multilobe_pdf(a, b, d, rough[i], al, metal[i]);
    }
}

extern "C" int ia_op_bsdf_sample_pdf(ia_ctx* c, const float* d_wi, const float* d_n, const float* d_rough,
                                     const float* d_albedo, const float* d_metal, const float* d_sample,
                                     const float* d_wo_query, int64_t n, float* d_wo, float* d_pdf, void* stream) {
    IA_REQUIRE(c && d_wi && d_n && d_rough && d_albedo && d_metal, IA_EINVAL, "ia_op_bsdf_sample_pdf: NULL argument");
    IA_REQUIRE(!d_wo || d_sample, IA_EINVAL, "ia_op_bsdf_sample_pdf: d_wo needs d_sample");
    IA_REQUIRE(!d_pdf || d_wo_query || d_wo, IA_EINVAL, "ia_op_bsdf_sample_pdf: d_pdf needs d_wo_query or d_wo");
    if (n == 0) return IA_OK;
    k_op_bsdf<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_wi, d_n, d_rough, d_albedo, d_metal, d_sample,
                                                                           d_wo_query, n, d_wo, d_pdf);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

__global__ void k_op_env(const IaEnv E, const float* __restrict__ u, const float* __restrict__ dirs, long long n,
                         float* __restrict__ dirs_out, float* __restrict__ pdf, float* __restrict__ em) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d[3] = {0.f, 0.f, 1.f};
    if (u) {
        ia_env_sample(E, u[i * 2], u[i * 2 + 1], d);
        if (dirs_out) { dirs_out[i * 3] = d[0]; dirs_out[i * 3 + 1] = d[1]; dirs_out[i * 3 + 2] = d[2]; }
    }
    if (dirs) { d[0] = dirs[i * 3]; d[1] = dirs[i * 3 + 1]; d[2] = dirs[i * 3 + 2]; }
    if (pdf) pdf[i] = ia_env_pdf(E, d);
    if (em) {
        float e3[3];
        ia_env_eval(E, d, e3);
        em[i * 3] = e3[0]; em[i * 3 + 1] = e3[1]; em[i * 3 + 2] = e3[2];
    }
}

extern "C" int ia_op_env(ia_ctx* c, const float* d_u, const float* d_dirs_world, int64_t n, float* d_dirs_world_out,
                         float* d_pdf_out, float* d_em_out, void* stream) {
    IA_REQUIRE(c && (d_u || d_dirs_world), IA_EINVAL, "ia_op_env: need uniforms or directions");
    IA_REQUIRE(c->have_light, IA_ESTATE, "ia_op_env: call ia_set_light first");
    if (n == 0) return IA_OK;
    k_op_env<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c->env, d_u, d_dirs_world, n, d_dirs_world_out,
                                                                          d_pdf_out, d_em_out);
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// ================================================================================================
// Frame producer / consumer either side of the render path
//
// k_make_rays: AnimationDataset's rays (datasets/animation.py:13-34 make_rays, :29-33 transform_rays, :163-189
// __getitem__; concatenated to [n,8] by systems/intrinsic_avatar.py:100-109) generated on the device instead of
// being built with numpy per frame and copied over PCIe (8.4 MB per 512^2 frame).  The reference does this in
// float64 (K and the extrinsics are float64 arrays) and casts to float32 at the end; so does the kernel.
//   Kinv  : inv(K), row-major 3x3        c2w : the dataset-level camera (identity for AnimationDataset), 3x4
//   ext   : inv(w2c) of the frame, 3x4   (test split; identity otherwise)
struct IaCamera {
    double Kinv[9], c2w[12], ext[12];
};
__global__ void k_make_rays(const __grid_constant__ IaCamera cam, int H, int W, float near_, float far_, float* __restrict__ rays) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)H * W) return;
    const double* Kinv = cam.Kinv;
    const double* c2w = cam.c2w;
    const double* ext = cam.ext;
    const double x = (double)(float)(i % W), y = (double)(float)(i / W);
    // d_c = [x, y, 1] @ inv(K).T ; d_w = d_c @ c2w[:3,:3].T, normalised
    double dc[3], dw[3];
#pragma unroll
    for (int r = 0; r < 3; r++) dc[r] = x * Kinv[r * 3 + 0] + y * Kinv[r * 3 + 1] + 1.0 * Kinv[r * 3 + 2];
#pragma unroll
    for (int r = 0; r < 3; r++) dw[r] = dc[0] * c2w[r * 4 + 0] + dc[1] * c2w[r * 4 + 1] + dc[2] * c2w[r * 4 + 2];
    const double nrm = sqrt(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);
    // make_rays returns float32; transform_rays then works on those float32 arrays with a float32 c2w
    float o32[3], d32[3];
#pragma unroll
    for (int r = 0; r < 3; r++) { d32[r] = (float)(dw[r] / nrm); o32[r] = (float)c2w[r * 4 + 3]; }
    float* out = rays + i * 8;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const float e0 = (float)ext[r * 4 + 0], e1 = (float)ext[r * 4 + 1], e2 = (float)ext[r * 4 + 2], e3 = (float)ext[r * 4 + 3];
        out[r] = o32[0] * e0 + o32[1] * e1 + o32[2] * e2 + e3;
        out[3 + r] = d32[0] * e0 + d32[1] * e1 + d32[2] * e2;
    }
    out[6] = near_;
    out[7] = far_;
}

extern "C" int ia_make_rays(ia_ctx* c, const double* h_Kinv9, const double* h_c2w12, const double* h_ext12, int H, int W,
                            float near_plane, float far_plane, float* d_rays, void* stream) {
    IA_REQUIRE(c && h_Kinv9 && d_rays, IA_EINVAL, "ia_make_rays: NULL argument");
    IA_REQUIRE(H > 0 && W > 0, IA_EINVAL, "ia_make_rays: empty image");
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    static const double ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    IaCamera cam;  // passed by value: no copy, no sync
    memcpy(cam.Kinv, h_Kinv9, 9 * sizeof(double));
    memcpy(cam.c2w, h_c2w12 ? h_c2w12 : ident, 12 * sizeof(double));
    memcpy(cam.ext, h_ext12 ? h_ext12 : ident, 12 * sizeof(double));
    const long long n = (long long)H * W;
    k_make_rays<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cam, H, W, near_plane, far_plane, d_rays);
    c->n_launches += 1;
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// k_pack_rgb8: SaverMixin.get_rgb_image_ (utils/mixins.py:43-53) on the device: clip to [lo, hi], scale to 0..255,
// truncate to uint8 (numpy astype), optional RGB -> BGR swap (cv2.cvtColor before cv2.imwrite).  A frame then
// leaves the device as 1 byte per channel instead of 4.
__global__ void k_pack_rgb8(const float* __restrict__ img, long long n_pix, int C, float lo, float hi, int bgr,
                            uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix * C) return;
    const long long p = i / C;
    const int ch = (int)(i - p * C);
    float v = img[i];
    v = fminf(fmaxf(v, lo), hi);
    v = (v - lo) / (hi - lo) * 255.f;
    const int oc = (bgr && C >= 3 && ch < 3) ? 2 - ch : ch;
    out[p * C + oc] = (uint8_t)v;
}

extern "C" int ia_pack_rgb8(ia_ctx* c, const float* d_img, int64_t n_pix, int channels, float lo, float hi, int bgr,
                            uint8_t* d_out, void* stream) {
    IA_REQUIRE(c && d_img && d_out, IA_EINVAL, "ia_pack_rgb8: NULL argument");
    IA_REQUIRE(channels > 0 && hi > lo, IA_EINVAL, "ia_pack_rgb8: bad channels / data range");
    if (n_pix == 0) return IA_OK;
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const long long n = (long long)n_pix * channels;
    k_pack_rgb8<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_img, n_pix, channels, lo, hi, bgr, d_out);
    c->n_launches += 1;
    IA_LAUNCH_CHECK();
    return IA_OK;
}

// k_pack_grid8: one column of SaverMixin.get_image_grid_ (utils/mixins.py:116-144) written straight into the uint8 image
// grid [H][grid_w][3] at pixel column x0, so that a frame's whole PNG leaves the device as one 3-byte-per-pixel copy.
//   kind 0 'rgb'       get_rgb_image_ (:43-53): clip to [lo, hi], scale to 0..255, truncate; a 1- or 2-channel image is
//                      padded with zeros
//   kind 1 'grayscale' get_grayscale_image_ (:87-101): nan_to_num, clip / scale, truncate, then the colour map `lut`
//                      [256][3] (NULL = cmap None: the grey value three times)
// range: two floats on the DEVICE overriding lo / hi (data_range=None: the image's own min / max, computed without a
// host round trip).  The grid holds the channel order of the FILE (RGB); bgr = 1 swaps it for a cv2.imwrite caller.
__global__ void k_pack_grid8(const float* __restrict__ img, int H, int W, int C, int kind, float lo, float hi,
                             const float* __restrict__ range, const uint8_t* __restrict__ lut, uint8_t* __restrict__ grid,
                             int grid_w, int x0, int bgr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)H * W) return;
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    if (range) { lo = range[0]; hi = range[1]; }
    uint8_t px[3] = {0, 0, 0};
    if (kind == 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
            if (ch < C) {
                float v = fminf(fmaxf(img[i * C + ch], lo), hi);
                px[ch] = (uint8_t)((v - lo) / (hi - lo) * 255.f);
            }
    } else {
        float v = img[i];
        if (isnan(v)) v = 0.f;                                   // np.nan_to_num; +-inf end up clipped
        v = fminf(fmaxf(v, lo), hi);
        const uint8_t q = (uint8_t)((v - lo) / (hi - lo) * 255.f);
        if (lut) { px[0] = lut[q * 3]; px[1] = lut[q * 3 + 1]; px[2] = lut[q * 3 + 2]; }
        else px[0] = px[1] = px[2] = q;
    }
    uint8_t* o = grid + ((long long)y * grid_w + x0 + x) * 3;
    o[0] = bgr ? px[2] : px[0]; o[1] = px[1]; o[2] = bgr ? px[0] : px[2];
}

extern "C" int ia_pack_grid8(ia_ctx* c, const float* d_img, int H, int W, int channels, int kind, float lo, float hi,
                             const float* d_range, const uint8_t* d_lut, uint8_t* d_grid, int grid_w, int x0, int bgr,
                             void* stream) {
    IA_REQUIRE(c && d_img && d_grid, IA_EINVAL, "ia_pack_grid8: NULL argument");
    IA_REQUIRE(kind == 0 || kind == 1, IA_EINVAL, "ia_pack_grid8: kind must be 0 (rgb) or 1 (grayscale)");
    IA_REQUIRE(channels > 0 && channels <= 3 && (kind == 0 || channels == 1), IA_EINVAL, "ia_pack_grid8: bad channel count");
    IA_REQUIRE(d_range || hi > lo, IA_EINVAL, "ia_pack_grid8: empty data range");
    IA_REQUIRE(x0 >= 0 && x0 + W <= grid_w, IA_EINVAL, "ia_pack_grid8: column outside the grid");
    if ((long long)H * W == 0) return IA_OK;
    IA_CHECK_CUDA(cudaSetDevice(c->device));
    const long long n = (long long)H * W;
    k_pack_grid8<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_img, H, W, channels, kind, lo, hi, d_range, d_lut,
                                                                                d_grid, grid_w, x0, bgr);
    c->n_launches += 1;
    IA_LAUNCH_CHECK();
    return IA_OK;
}
